// Host-side objects behind the opaque handles of the C ABI (shared by prove.cu and setup.cu).
#pragma once
#include "common.cuh"
#include <memory>
#include <vector>

namespace za {

// ------------------------------------------------------------------ device-resident bases
struct Bases {
    Ctx* ctx;
    int group;            // 1 = G1, 2 = G2
    size_t n;
    bool has_infinity;
    DevBuf pts;           // Affine<Fq> or Affine<Fq2>, Montgomery form
    // optional fixed-base table: table[i*W + w] = 2^(c w) * pts[i]  (see bases_table_kernel)
    DevBuf table;
    int tab_c = 0, tab_W = 0;
    size_t tab_lo = 0, tab_n = 0;     // the table covers bases [tab_lo, tab_lo + tab_n)
};


// ------------------------------------------------------------------ proving key
struct Pk {
    Ctx* ctx;
    // verifying key, host side, Montgomery form
    G1Affine alpha_g1, beta_g1, delta_g1;
    G2Affine beta_g2, gamma_g2, delta_g2;
    std::vector<G1Affine> ic;
    std::unique_ptr<Bases> h, l, a, b_g1, b_g2;
    uint32_t rank0_weight = 1000;   // per mille: rank 0's share of the witness multiexps relative to the other ranks
    // Explicit point ranges of this device (za_pk_partition_ranges): which part of each of the five queries it adds up.
    // Order: H (over m - 1 coefficients), L (over the aux), A (over the A exponent list), B in G1, B in G2 (both over the B
    // exponent list).  Not set: the ranges follow (rank, world, rank0_weight) of the call.
    bool plan_set = false;
    size_t plan_lo[5] = {0, 0, 0, 0, 0}, plan_hi[5] = {0, 0, 0, 0, 0};
};
enum { Q_H = 0, Q_L = 1, Q_A = 2, Q_B1 = 3, Q_B2 = 4 };


// ------------------------------------------------------------------ circuit (R1CS on device)
// Everything that depends only on the constraint system is computed once at upload: the CSR matrices in
// Montgomery form, the three density maps of bellman's ProvingAssignment (they depend on which variables occur in
// A / B rows, never on the witness: density.inc(i) fires for every term) and the compacted index
// lists the density-filtered multiexps need (K10).
struct Circuit {
    Ctx* ctx;
    uint32_t ni, na, nc;
    DevBuf ptr[3], col[3], coeff[3];     // col = slot in the witness vector [inputs | aux]
    std::vector<uint32_t> h_ptr[3], h_col[3];   // host copies (the trusted setup needs the column view)
    std::vector<uint8_t> a_aux_density, b_in_density, b_aux_density;
    DevBuf a_aux_idx, b_in_idx, b_aux_idx;
    uint32_t a_aux_total = 0, b_in_total = 0, b_aux_total = 0;
    // positions in the witness vector [inputs | aux] of the exponents of the whole A query (all inputs, then
    // the aux with a_aux_density) and of the whole B query (inputs with b_input_density, then aux with b_aux_density)
    DevBuf a_cat_idx, b_cat_idx;
    std::vector<uint32_t> h_a_cat, h_b_cat;      // host copies of the two lists
    uint32_t a_cat_total = 0, b_cat_total = 0;
};


}  // namespace za

struct za_bases { std::unique_ptr<za::Bases> b; };
struct za_pk { std::unique_ptr<za::Pk> p; };
struct za_circuit { std::unique_ptr<za::Circuit> c; };

