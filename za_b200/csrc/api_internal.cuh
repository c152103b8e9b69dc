// Internal host entry points shared between the translation units of libza_b200.so.
#pragma once
#include "common.cuh"

namespace za {

// ntt.cu
Fr host_domain_omega(int log_n);
void launch_powers_public(Ctx* ctx, Fr* out, size_t count, const Fr& base, const Fr& first);
void fr_convert(Ctx* ctx, Fr* d, size_t n, int dir);
void ntt_mode(Ctx* ctx, Fr* buf, int log_n, int mode, int batch);
void h_poly_device(Ctx* ctx, Fr* a, Fr* b, Fr* c, int log_m, Fr* h_out = nullptr);
void h_poly_checkpointed(Ctx* ctx, Fr* a, Fr* b, Fr* c, int log_m, uint8_t* ck);

// msm.cu
template <class F> uint32_t bases_import(Ctx* ctx, void* d_pts, size_t n);
template <class F> XYZZ<F> msm_run(Ctx* ctx, const Affine<F>* d_bases, const uint32_t* d_scalars, size_t n, bool has_infinity);
// a further query sorted / accumulated / reduced together with the first one (fixed-base table of the same window layout)
struct MsmPart { const void* table; const uint32_t* scalars; size_t n; };
template <class F> void msm_enqueue(Ctx* ctx, int slot, const Affine<F>* d_bases, const uint32_t* d_scalars, size_t n, bool has_infinity, int share_sort,
                                    const Affine<F>* d_table, int tab_c, int tab_W, const MsmPart* more = nullptr, int n_more = 0);
template <class F> void bases_table_build(Ctx* ctx, const Affine<F>* d_pts, size_t n, int c, int W, Affine<F>* d_table);
template <class F> XYZZ<F> msm_finish(Ctx* ctx, int slot, XYZZ<F>* parts_out = nullptr);
void msm_abort(Ctx* ctx);
int msm_window_bits(size_t n);
template <class F> void bases_generate(Ctx* ctx, Affine<F>* d_out, size_t n, uint64_t first, const Affine<F>& G);
double imad_peak(Ctx* ctx);
template <class F> void xyzz_normalise(Ctx* ctx, const XYZZ<F>* d_in, Affine<F>* d_out, size_t n);
Fq host_g1_b();
Fq2 host_g2_b();

}  // namespace za

// the opaque context of the C ABI
struct za_ctx {
    za::Ctx c;
    cudaStream_t own = nullptr;
};
