/* za_b200 — C ABI of the B200-native Groth16 proving backend for za.
 *
 * This is the drop-in boundary.  The reference has no plugin API: its hot path sits
 * behind (inner) the Rust calls that /root/reference/prover/src/groth16/prover.rs makes into
 * bellman_ce/pairing_ce and (outer) the C ABI of binding/c.  A Rust toolchain is not available in the
 * build image, so the seam is expressed as plain C: handles, caller-owned buffers, int status codes and
 * za_last_error().  Each entry point names the reference interface it replaces.  INTEGRATION.md shows
 * the `extern "C"` block a maintainer adds to the prover crate to bind them.
 *
 * Encodings at the seam
 *   Fr scalar  : 32 bytes little-endian canonical integer (== FrRepr([u64;4]) memory image,
 *                what `into_repr()` yields at bellman's multiexp boundary)
 *   G1 affine  : 64 bytes  x || y, 32-byte LE canonical each; 64 zero bytes = infinity
 *   G2 affine  : 128 bytes x.c0 || x.c1 || y.c0 || y.c1;     128 zero bytes = infinity
 *   proof      : a (64) || b (128) || c (64)
 *   Parameters : the exact byte stream bellman's Parameters::write produces (big-endian
 *                uncompressed points, u32 BE counts) — /root/reference/prover/src/groth16/format.rs:250,285
 *
 * All functions return ZA_OK (0) or a negative ZA_ERR_* code; za_last_error() gives the text.
 * Nothing here falls back to a CPU implementation: without a CUDA device every compute entry
 * point fails with ZA_ERR_CUDA.
 */
#ifndef ZA_B200_H
#define ZA_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct za_ctx za_ctx;     /* one per GPU: stream, cached NTT domains, scratch */
typedef struct za_bases za_bases; /* a device-resident array of G1 or G2 affine bases (Montgomery form) */
typedef struct za_pk za_pk;       /* device-resident bellman Parameters (vk on host, queries on device) */
typedef struct za_circuit za_circuit; /* device-resident constraint system + its density maps */

enum {
    ZA_OK = 0,
    ZA_ERR_CUDA = -1,                 /* no device / CUDA failure */
    ZA_ERR_INVALID = -2,              /* bad argument */
    ZA_ERR_UNEXPECTED_IDENTITY = -3,  /* SynthesisError::UnexpectedIdentity */
    ZA_ERR_POLY_DEGREE_TOO_LARGE = -4,/* SynthesisError::PolynomialDegreeTooLarge (domain >= 2^28) */
    ZA_ERR_IO = -5,                   /* short / malformed Parameters stream, query too short */
    ZA_ERR_NOT_ON_CURVE = -6,
    ZA_ERR_NOT_IN_SUBGROUP = -7,
    ZA_ERR_BAD_ENCODING = -8,
    ZA_ERR_BUFFER_TOO_SMALL = -9,
    ZA_ERR_NOT_CANONICAL = -10,       /* a scalar >= r was passed */
    ZA_ERR_UNCONSTRAINED_VARIABLE = -11 /* SynthesisError::UnconstrainedVariable (setup) */
};

#define ZA_VAR_AUX 0x80000000u /* constraint term variable: bit 31 set = aux index, else input index */

/* What bellman's ConstraintSystem receives from CircomCircuit::synthesize
 * (/root/reference/prover/src/groth16/prover.rs:45-103): the enforce(A, B, C) rows as three CSR
 * matrices, terms in insertion order.  Input 0 is the constant `one`; C already carries the sign
 * flip of prover.rs:98.  The input-consistency rows bellman appends are added internally. */
typedef struct {
    uint32_t num_inputs;      /* including `one` */
    uint32_t num_aux;
    uint32_t num_constraints;
    const uint32_t *ptr[3];   /* [num_constraints + 1] term offsets for A, B, C */
    const uint32_t *var[3];   /* variable of each term */
    const uint8_t *coeff[3];  /* 32-byte LE canonical coefficient of each term */
} za_r1cs;

/* Optional intermediates of za_create_proof for element-wise parity checks; any pointer may be NULL. */
typedef struct {
    uint8_t *a_eval, *b_eval, *c_eval; /* (num_constraints + num_inputs) * 32 */
    uint8_t *h_coeffs;                 /* (m - 1) * 32 */
    uint8_t *msm_g1;                   /* 7 * 64: h, l, a_inputs, a_aux, b1_inputs, b1_aux, (unused) */
    uint8_t *msm_g2;                   /* 2 * 128: b2_inputs, b2_aux */
    uint8_t *a_aux_density, *b_input_density, *b_aux_density; /* one byte per variable */
} za_trace;

/* ---- context ------------------------------------------------------------------------------ */
const char *za_last_error(void);
int za_version(void);
int za_device_count(void);
int za_ctx_create(int device, za_ctx **out);
void za_ctx_destroy(za_ctx *ctx);
/* Issue all work of this context on `cuda_stream` (a cudaStream_t; NULL is the CUDA default stream).
 * A new context starts on its own non-blocking stream; za_ctx_use_own_stream goes back to it. */
int za_ctx_set_stream(za_ctx *ctx, void *cuda_stream);
int za_ctx_use_own_stream(za_ctx *ctx);
int za_ctx_synchronize(za_ctx *ctx);
/* Kernels launched through this context so far (bench.py reports the delta as gpu_launches). */
uint64_t za_ctx_launch_count(const za_ctx *ctx);

/* Per-kernel-class device timing (CUDA events on the context's stream), used by bench.py for the roofline
 * numbers.  za_ctx_profile_read synchronises, fills out[24] = ms[8] | work[8] | spans[8] and resets.
 * Classes: 0 msm_accumulate G1 (work = mixed additions), 1 msm_accumulate G2, 2 NTT transform (work = elements),
 * 3 MSM digit sort (work = scalars), 4 MSM bucket reduction (work = buckets), 5 constraint evaluation (work = rows),
 * 6 H pointwise (work = elements), 7 unused. */
int za_ctx_profile(za_ctx *ctx, int on);
int za_ctx_profile_read(za_ctx *ctx, double *out);

/* ---- EvaluationDomain --------------------------------------------------------------------
 * Replaces bellman_ce domain.rs EvaluationDomain::{fft, ifft, coset_fft, icoset_fft} as used by
 * create_proof (entered at prover.rs:173).  Natural order in and out, like bellman's.
 * mode: 0 fft, 1 ifft, 2 coset_fft, 3 icoset_fft. */
enum { ZA_NTT_FFT = 0, ZA_NTT_IFFT = 1, ZA_NTT_COSET_FFT = 2, ZA_NTT_ICOSET_FFT = 3 };
/* host buffer of 2^log_n canonical scalars, transformed in place (H2D + D2H inside) */
int za_ntt(za_ctx *ctx, uint8_t *data, int log_n, int mode);
/* device-resident: `d_data` holds `batch` vectors of 2^log_n Montgomery-form Fr (32 B each), in place */
int za_ntt_device(za_ctx *ctx, void *d_data, int log_n, int mode, int batch);
/* canonical <-> Montgomery on device, n elements in place (dir 0: to Montgomery, 1: to canonical) */
int za_fr_convert_device(za_ctx *ctx, void *d_data, size_t n, int dir);
/* create_proof's H block: a, b, c evaluations (len each, canonical, host) -> h coefficients
 * ((m-1) * 32 bytes, m = next power of two >= len).  checkpoints (optional, 8*m*32 bytes) receives
 * a.ifft, a.coset_fft, b.ifft, b.coset_fft, c.ifft, c.coset_fft, (a*b-c)/Z, icoset_fft — computed by
 * the unfused transforms so every intermediate vector can be compared with the reference's. */
int za_h_poly(za_ctx *ctx, const uint8_t *a, const uint8_t *b, const uint8_t *c, size_t len,
              uint8_t *h_out, uint8_t *checkpoints);
/* device-resident fused version: d_a/d_b/d_c hold m Montgomery Fr each (zero padded), result
 * (canonical form, m entries of which the first m-1 are the h scalars) is left in d_a. */
int za_h_poly_device(za_ctx *ctx, void *d_a, void *d_b, void *d_c, int log_m);

/* ---- multiexp ------------------------------------------------------------------------------
 * Replaces bellman_ce multiexp.rs `multiexp(pool, (bases, skip), density, exponents)`.
 * group: 1 = G1, 2 = G2. */
int za_bases_upload(za_ctx *ctx, int group, const uint8_t *bases, size_t n, za_bases **out);
void za_bases_free(za_bases *b);
size_t za_bases_len(const za_bases *b);
/* sum over i of scalars[i] * bases[offset + k(i)], where k(i) counts the set density entries before i
 * (density == NULL: FullDensity, k(i) = i).  scalars: n_exp canonical 32-byte values on the host.
 * out: affine point (64 or 128 bytes).  Errors like bellman: ZA_ERR_IO if the bases run out. */
int za_multiexp(za_ctx *ctx, const za_bases *bases, size_t offset, const uint8_t *scalars, size_t n_exp,
                const uint8_t *density, uint8_t *out);
/* device-resident scalars (canonical 32-byte LE, n of them, FullDensity) */
int za_multiexp_device(za_ctx *ctx, const za_bases *bases, size_t offset, const void *d_scalars, size_t n,
                       uint8_t *out);
/* Multi-GPU building block: like za_multiexp_device but returns the unnormalised partial
 * sum as XYZZ coordinates in canonical form (4 x 32 bytes G1, 4 x 64 bytes G2) so that per-GPU partial
 * results can be added on the host with za_point_sum. */
int za_multiexp_partial_device(za_ctx *ctx, const za_bases *bases, size_t offset, const void *d_scalars,
                               size_t n, uint8_t *out_xyzz);
/* host: add `count` XYZZ partial sums and normalise to affine */
int za_point_sum(int group, const uint8_t *xyzz, size_t count, uint8_t *out_affine);

/* ---- Groth16 -------------------------------------------------------------------------------
 * za_pk_load replaces bellman Parameters::read(reader, checked) (format.rs:285): parses the byte
 * stream, converts to Montgomery form on the GPU, rejects off-curve points, points at infinity in the
 * queries and (checked != 0) G2 points outside the r-torsion. */
int za_pk_load(za_ctx *ctx, const uint8_t *params, size_t len, int checked, za_pk **out);
void za_pk_free(za_pk *pk);
/* counts[6] = |ic|, |h|, |l|, |a|, |b_g1|, |b_g2| */
int za_pk_counts(const za_pk *pk, uint32_t *counts);
/* vk_out: alpha_g1 (64) beta_g1 (64) beta_g2 (128) gamma_g2 (128) delta_g1 (64) delta_g2 (128) ic (64 each),
 * interchange encoding — what JsonVerifyingKey::from_bellman (format.rs:143-160) reads from params.vk */
int za_pk_vk(const za_pk *pk, uint8_t *vk_out, size_t size);
/* The constraint system is what CircomCircuit::synthesize feeds bellman (prover.rs:45-103).  It is
 * uploaded once per circuit: the CSR matrices, and the density maps of bellman's ProvingAssignment,
 * which depend only on the constraint structure. */
int za_circuit_upload(za_ctx *ctx, const za_r1cs *cs, za_circuit **out);
void za_circuit_free(za_circuit *c);
/* Replaces bellman groth16::create_proof(circuit, params, r, s) (create_random_proof at prover.rs:173
 * is this with r, s drawn from the RNG).  inputs: num_inputs canonical scalars (inputs[0] = 1),
 * aux: num_aux.  proof_out: 256 bytes. */
int za_create_proof(za_ctx *ctx, const za_pk *pk, const za_circuit *circuit, const uint8_t *inputs, const uint8_t *aux,
                    const uint8_t *r, const uint8_t *s, uint8_t *proof_out, za_trace *trace);
/* Same with the witness [inputs | aux] (canonical, (num_inputs+num_aux)*32 bytes) already in device memory. */
int za_create_proof_device(za_ctx *ctx, const za_pk *pk, const za_circuit *circuit, const void *d_witness,
                           const uint8_t *r, const uint8_t *s, uint8_t *proof_out);
/* constraints.satisfies_with_signals (constraint.rs:29-67; called before proving at prover.rs:155), on the GPU:
 * *first_bad = -1 if every row satisfies A(w) * B(w) == C(w), else the smallest failing row. */
int za_circuit_satisfied(za_ctx *ctx, const za_circuit *circuit, const uint8_t *inputs, const uint8_t *aux,
                         int64_t *first_bad);
/* info[7] = num_inputs, num_aux, num_constraints, |a_aux_density|, |b_input_density|, |b_aux_density|, log2(m) */
int za_circuit_info(const za_circuit *circuit, uint32_t *info);

/* ---- create_proof in stages, for one process per GPU ----------------------------------
 * Stage 1, one GPU: witness -> a, b, c -> H coefficients.  d_h: m * 32 bytes of device memory; on
 *   return its first m-1 entries are the canonical h scalars.
 * Stage 2, every GPU (each holds the proving key): the eight multiexps restricted to this rank's
 *   contiguous point range of every query; d_h needs to be valid on [lo, hi) = share of m-1 only.
 *   partials_out: ZA_PARTIALS_BYTES (6 G1 + 2 G2 XYZZ sums, canonical coordinates).
 * Stage 3, host: add the `world` partial records and assemble the proof (a few group operations). */
#define ZA_PARTIALS_BYTES 1280
/* Optional, once per (pk, circuit, rank): rebuild the proving key's fixed-base tables for exactly the point range
 * this rank owns in every query, with the window size chosen for the share (smaller shares want fewer buckets). */
int za_pk_partition(za_ctx *ctx, za_pk *pk, const za_circuit *circuit, int rank, int world);
/* The same with rank 0 taking only rank0_weight_permille / 1000 of an ordinary rank's share of the witness
 * multiexps (L, A, B): rank 0 also runs the H-polynomial pipeline of stage 1 while the others already work. */
/* The point-range rule itself (host only): [lo, hi) of `count` items for `rank`; 1000 = equal shares. */
int za_share_weighted(uint64_t count, int rank, int world, uint32_t rank0_weight_permille, uint64_t *lo, uint64_t *hi);
int za_pk_partition_weighted(za_ctx *ctx, za_pk *pk, const za_circuit *circuit, int rank, int world,
                             uint32_t rank0_weight_permille);
/* Explicit point ranges instead of (rank, world): this device adds up [lo[q], hi[q]) of query q, q = 0 H (over the m - 1
 * coefficients), 1 L (over the aux), 2 A (over the A exponent list: all inputs, then the aux of a_aux_density), 3 B in G1,
 * 4 B in G2 (both over the B exponent list).  Builds the fixed-base tables of exactly these ranges; afterwards the rank /
 * world arguments of za_prove_msm_enqueue / _partials are ignored for this key.  A proof is the sum of the partial records of
 * any set of devices whose ranges tile every query.  lo = hi = NULL: back to the whole queries.
 * za_prover_plan: the ranges za_prover gives each of n_devices devices (lo_out / hi_out: 5 entries per device): the four
 * witness queries laid end to end in units of estimated time and cut into ONE contiguous piece per device (a device then
 * runs one or two long multiexps instead of a slice of all five), H split evenly, device 0 (which runs the H pipeline
 * first) a shorter piece. */
int za_pk_partition_ranges(za_ctx *ctx, za_pk *pk, const za_circuit *circuit, const uint64_t *lo, const uint64_t *hi);
int za_prover_plan(const za_circuit *circuit, int n_devices, uint64_t *lo_out, uint64_t *hi_out);
/* the same from the five query lengths (H, L, A, B, B) and the domain size alone: host only, no device needed */
int za_prover_plan_counts(const uint64_t *counts, uint64_t domain, int n_devices, uint64_t *lo_out, uint64_t *hi_out);
int za_prove_h_device(za_ctx *ctx, const za_circuit *circuit, const void *d_witness, void *d_h);
/* Destinations of the NEXT za_prove_h_device on this context: h[k] is stored at (char *)outs[j] + 32 k for the first j with
 * k < his[j] (the last part takes the rest) instead of d_h.  outs[j] may be memory of a peer device this context's device
 * has peer access to: the last pass of the last transform then writes every rank's slice into that rank's memory over
 * NVLink and no copy follows the H pipeline (what za_prover does).  n <= 16; needs a domain of 2^12 or more; consumed
 * (cleared) by that call.  n = 0 clears it. */
int za_ctx_set_h_scatter(za_ctx *ctx, int n, void *const *outs, const uint64_t *his);
/* Stage 2 in two asynchronous halves, so that the multiexps that do not depend on the H polynomial (ZA_MSM_WITNESS:
 * L, A, B-G1, B-G2) run on every GPU while rank 0 still computes it; ZA_MSM_H (needs d_h on the rank's share)
 * follows once the h scalars have arrived.  za_prove_msm_collect waits for all of them -> partials record. */
#define ZA_MSM_WITNESS 1
#define ZA_MSM_H 2
int za_prove_msm_enqueue(za_ctx *ctx, const za_pk *pk, const za_circuit *circuit, const void *d_witness,
                         const void *d_h, int rank, int world, int which);
int za_prove_msm_collect(za_ctx *ctx, uint8_t *partials_out);
int za_prove_msm_partials(za_ctx *ctx, const za_pk *pk, const za_circuit *circuit, const void *d_witness,
                          const void *d_h, int rank, int world, uint8_t *partials_out);
int za_prove_assemble(const za_pk *pk, const uint8_t *partials, int world, const uint8_t *r, const uint8_t *s,
                      uint8_t *proof_out);

/* ---- several GPUs of one box behind ONE call ----------------------------------------------
 * The reference's caller is a single process (create_random_proof at prover.rs:173).  A za_prover owns a context, the
 * proving key and the circuit on every listed device and one host thread per device; za_prover_create_proof is
 * create_proof with every multiexp cut by point range over the devices: device 0 evaluates the constraints and runs
 * the H-polynomial transforms, writes each peer's slice of the h scalars straight into that peer's memory over
 * NVLink (cudaMemcpyPeerAsync), every device uploads only the witness span its point ranges read, and the host adds
 * the per-device partial sums (8 points each) and assembles the proof.  No collective; nothing but CUDA runtime calls.
 * With one device it is za_create_proof. */
typedef struct za_prover za_prover;
int za_prover_create(const int *devices, int n_devices, za_prover **out);
void za_prover_destroy(za_prover *p);
int za_prover_device_count(const za_prover *p);
/* the context of device k (its launch counter and per-class timing: za_ctx_launch_count, za_ctx_profile) */
za_ctx *za_prover_ctx(za_prover *p, int k);
/* Parameters::read(params, checked) (format.rs:285) on every device, or a key of known multiples (za_pk_synthetic) */
int za_prover_load_pk(za_prover *p, const uint8_t *params, size_t len, int checked);
int za_prover_synthetic_pk(za_prover *p, const uint32_t *counts);
int za_prover_set_circuit(za_prover *p, const za_r1cs *cs);
int za_prover_vk(const za_prover *p, uint8_t *vk_out, size_t size);
int za_prover_pk_counts(const za_prover *p, uint32_t *counts);
/* create_proof(circuit, params, r, s).  inputs / aux: host buffers as for za_create_proof; inputs == NULL proves the
 * witness of the last za_prover_upload_witness again (witness resident on the devices: bench.py's `value`). */
int za_prover_upload_witness(za_prover *p, const uint8_t *inputs, const uint8_t *aux);
int za_prover_create_proof(za_prover *p, const uint8_t *inputs, const uint8_t *aux, const uint8_t *r, const uint8_t *s,
                           uint8_t *proof_out);
uint64_t za_prover_launch_count(const za_prover *p);
/* info[3] = witness bytes uploaded per proof over all devices, device 0's share weight (per mille), bytes read back */
int za_prover_info(const za_prover *p, uint64_t *info);

/* ---- synthetic inputs and measurement utilities ---------------------------------------
 * bases[i] = (first_multiple + i) * G: distinct points with known discrete logarithms, so a full-size
 * multiexp is checkable with one scalar multiplication.  Generated on the GPU. */
int za_bases_generate(za_ctx *ctx, int group, size_t n, uint64_t first_multiple, za_bases **out);
int za_bases_download(za_ctx *ctx, const za_bases *bases, size_t offset, size_t n, uint8_t *out);
/* Build the fixed-base table of a bases array (entry i*W + w = 2^(c w) * P_i; W = ceil(255 / c) copies of the
 * array): every later multiexp over it uses ONE bucket space for all windows and c up to 20 (13 windows instead
 * of 16).  Proving-key queries get their table at load.  Returns the window size c (> 0), 0 if no table was
 * built (fewer than 4096 points, a point at infinity, or a table over the memory bound: 64 GiB or 45 % of the free device memory), or a negative error. */
int za_bases_precompute(za_ctx *ctx, za_bases *bases);
/* A proving key of the given query sizes (counts[6] = |ic|, |h|, |l|, |a|, |b_g1|, |b_g2|) whose bases are
 * known multiples of the generators: query q entry i = ((q+1) * 2^32 + i + 1) * G with q = 0..3 for
 * h, l, a, b (b_g1 and b_g2 share multipliers); alpha, beta, gamma, delta = 3, 5, 7, 11; ic[i] = 13 + i.
 * Not a valid CRS — the work per proof is identical and every proof element has a closed form. */
int za_pk_synthetic(za_ctx *ctx, const uint32_t *counts, za_pk **out);
/* Measured 32-bit integer multiply-add throughput of the device (dependency-free mad.lo.u32 on all SMs). */
int za_imad_peak(za_ctx *ctx, double *imads_per_second);

/* ---- verification (host side, no GPU needed) ------------------------------------------------------------
 * Replaces bellman prepare_verifying_key + verify_proof (prover.rs:191-200, helper.rs:153-158).
 * vk: the za_pk_vk layout (576 + 64 * n_ic bytes).  *valid = 1 / 0.  A wrong number of public inputs is
 * ZA_ERR_INVALID (bellman: SynthesisError::MalformedVerifyingKey). */
int za_verify_proof(const uint8_t *vk, size_t n_ic, const uint8_t *proof, const uint8_t *public_inputs,
                    size_t n_public, int *valid);
/* JsonVerifyingKey (format.rs:130-167): alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, gamma_g2, ic, input_names */
int za_vk_to_json(const uint8_t *vk, size_t n_ic, const char *const *input_names, size_t n_names, char *buf,
                  size_t size);
/* generate_solidity (prover/src/groth16/ethereum.rs:216-261): the verifier contract as text, by substitution of the
 * eight placeholders <%vk_a%> <%vk_b%> <%vk_gamma%> <%vk_delta%> <%vk_inputs_length%> <%vk_inputs%>
 * <%vk_gammaABC_length%> <%vk_gammaABC_pts%> — G1 as "x,y", G2 as "[x.c1,x.c0],[y.c1,y.c0]" (ethereum.rs:227-238),
 * coordinates in ff_ce's Repr Display ("0x" + 64 hex digits), the input names in Rust's {:?} list form.
 * contract_template: the text to fill (the reference's CONTRACT_TEMPLATE, ethereum.rs:8-214, gives the reference's
 * exact contract); NULL = the library's own template (an independent Groth16 verifier contract with the same
 * verifyTx(a, b, c, input) / Verified interface).  *needed (optional) receives the size the text needs. */
int za_vk_to_solidity(const uint8_t *vk, size_t n_ic, const char *const *input_names, size_t n_names,
                      const char *contract_template, char *buf, size_t size, size_t *needed);
/* helper::verify (helper.rs:149-158): vk JSON + proof-with-inputs JSON -> *valid */
int za_verify_json(const char *vk_json, const char *proof_json, int *valid);

/* ---- proving.key container and circuit synthesis (host side) ---------------------------------------------
 * read_pk / write_pk of format.rs:223-293: AST blob (opaque), constraints (bincode QEQ), ignored signals,
 * then bellman's Parameters.  Constraints cross the ABI as three CSR matrices over SIGNAL ids, 32-byte LE
 * coefficients, za's convention a*b + c = 0.  Two passes: za_pkfile_scan sizes the buffers
 * (info[6] = num_constraints, num_ignore, nnz_a, nnz_b, nnz_c, max signal id + 1), za_pkfile_read fills them. */
int za_pkfile_scan(const uint8_t *file, size_t len, uint64_t *info, size_t *params_offset, size_t *ast_offset,
                   size_t *ast_len);
int za_pkfile_read(const uint8_t *file, size_t len, uint32_t *const *ptr, uint32_t *const *sig,
                   uint8_t *const *coeff, uint32_t *ignore);
int za_pkfile_write(const uint8_t *ast, size_t ast_len, uint32_t num_constraints, const uint32_t *const *ptr,
                    const uint32_t *const *sig, const uint8_t *const *coeff, const uint32_t *ignore,
                    uint32_t num_ignore, const uint8_t *params, size_t params_len, uint8_t *out, size_t size,
                    size_t *out_len);
/* CircomCircuit::synthesize (prover.rs:45-103) as data: signal -> bellman variable (input if is_public, else aux,
 * none if ignored; signal 0 = input 0 = one), per-term variables for za_r1cs, and C = -c (prover.rs:98). */
int za_synthesize(uint32_t n_signals, const uint8_t *is_public, const uint32_t *ignore, uint32_t num_ignore,
                  uint32_t num_constraints, const uint32_t *const *ptr, const uint32_t *const *sig,
                  const uint8_t *c_coeff, uint32_t *var_of_signal, uint32_t *const *out_var, uint8_t *out_c_coeff,
                  uint32_t *num_inputs, uint32_t *num_aux);

/* ---- trusted setup --------------------------------------------------------------------------------------
 * Replaces bellman generate_parameters(circuit, g1, g2, alpha, beta, gamma, delta, tau): what
 * generate_random_parameters (prover.rs:122) computes after drawing those seven values from the RNG.
 * Scalars: 32-byte LE canonical; g1 / g2: affine interchange encoding.  out receives the byte stream of
 * bellman's Parameters::write; size it with za_parameters_max_size (the A / B queries shrink when points at
 * infinity are filtered out; *out_len is the actual length). */
size_t za_parameters_max_size(const za_circuit *circuit);
int za_generate_parameters(za_ctx *ctx, const za_circuit *circuit, const uint8_t *alpha, const uint8_t *beta,
                           const uint8_t *gamma, const uint8_t *delta, const uint8_t *tau, const uint8_t *g1,
                           const uint8_t *g2, uint8_t *out, size_t size, size_t *out_len);

/* JsonProofAndInput (format.rs:80-128): compact JSON, "0x"+64 hex coordinates, decimal public inputs.
 * public_inputs: n canonical scalars. Returns ZA_ERR_BUFFER_TOO_SMALL if len >= size (binding/c lib.rs:23). */
int za_proof_to_json(const uint8_t *proof, const uint8_t *public_inputs, size_t n_public, char *buf, size_t size);

#ifdef __cplusplus
}
#endif
#endif
