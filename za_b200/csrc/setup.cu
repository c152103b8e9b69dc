// Trusted setup on the GPU: bellman_ce groth16/generator.rs `generate_parameters(circuit, g1, g2, alpha, beta,
// gamma, delta, tau)` — what `generate_random_parameters` at /root/reference/prover/src/groth16/prover.rs:122
// computes after drawing those seven values from the RNG ("next" row N2).
//
//   powers of tau -> h[i] = g1 * (tau^i * (tau^m - 1) / delta), i < m-1
//   ifft(powers)  -> Lagrange coefficients L_k(tau)
//   per variable v: at = sum coeff * L_row over column v of A (inputs also get L of their consistency row), bt, ct
//   a = g1*at, b_g1 = g1*bt, b_g2 = g2*bt, ic / l = g1 * ((beta*at + alpha*bt + ct) / gamma | delta)
//   points at infinity are filtered out of a, b_g1, b_g2; any l at infinity is UnconstrainedVariable
// The 4 n + m group elements are independent fixed-base scalar multiplications: an 8-bit windowed table of the base
// (32 x 255 affine points, L2 resident) and 32 mixed additions per scalar, then one shared-inversion normalisation.
// Output: the byte stream of bellman's Parameters::write (format.rs:250).
#include "common.cuh"
#include "api_internal.cuh"
#include "objects.cuh"
#include <string.h>

namespace za {

template <class T>
static __device__ __forceinline__ T ld16(const T* p) {
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) d[i] = s[i];
    return r;
}
template <class T>
static __device__ __forceinline__ void st16(T* p, const T& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) d[i] = s[i];
}

#define FB_WINDOWS 32
#define FB_ENTRIES 255
// tab[j*255 + d-1] = d * 2^(8j) * G
template <class F>
__global__ void fixed_table_kernel(XYZZ<F>* tab, const Affine<F> G) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= FB_WINDOWS) return;
    XYZZ<F> P = XYZZ<F>::from_affine(G);
    for (int k = 0; k < 8 * j; k++) P = xyzz_dbl<F>(P);
    XYZZ<F> acc = P;
    for (int d = 1; d <= FB_ENTRIES; d++) {
        st16(tab + j * FB_ENTRIES + d - 1, acc);
        xyzz_add<F>(acc, P);
    }
}
// out[i] = scalar[i] * G   (scalars in Montgomery form)
template <class F>
__global__ void __launch_bounds__(128) fixed_mul_kernel(const Affine<F>* __restrict__ tab, const Fr* __restrict__ scalars, size_t n, XYZZ<F>* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr k = fp_from_mont<FrParams>(ld16(scalars + i));
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int j = 0; j < FB_WINDOWS; j++) {
        uint32_t d = (k.v[j >> 2] >> (8 * (j & 3))) & 0xffu;
        if (d) {
            Affine<F> P = ld16(tab + j * FB_ENTRIES + d - 1);
            xyzz_madd<F>(acc, P.x, P.y, false);
        }
    }
    st16(out + i, acc);
}
// column view of one constraint matrix: out[v] = sum over the terms of column v of coeff * L[row]
__global__ void qap_column_kernel(const uint32_t* __restrict__ colptr, const uint32_t* __restrict__ term, const uint32_t* __restrict__ row,
                                  const Fr* __restrict__ coeff, const Fr* __restrict__ L, uint32_t nv, Fr* out) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    Fr acc = Fr::zero();
    for (uint32_t t = colptr[v]; t < colptr[v + 1]; t++) acc = acc + ld16(coeff + term[t]) * ld16(L + row[t]);
    st16(out + v, acc);
}
// at[i] += L[nc + i] for the inputs (the input-consistency rows bellman appends); then
// ext[v] = (beta*at + alpha*bt + ct) * (v < ni ? gamma^-1 : delta^-1)
__global__ void qap_ext_kernel(Fr* at, const Fr* bt, const Fr* ct, const Fr* L, uint32_t ni, uint32_t nv, uint32_t nc, const Fr alpha, const Fr beta,
                               const Fr gamma_inv, const Fr delta_inv, Fr* ext) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    Fr a = ld16(at + v);
    if (v < ni) { a = a + ld16(L + nc + v); st16(at + v, a); }
    Fr e = (beta * a + alpha * ld16(bt + v) + ld16(ct + v)) * (v < ni ? gamma_inv : delta_inv);
    st16(ext + v, e);
}
__global__ void fr_scale_const_kernel(const Fr* in, Fr* out, size_t n, const Fr k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st16(out + i, ld16(in + i) * k);
}
// affine Montgomery -> bellman uncompressed big-endian record (infinity: 0x40 then zeros)
template <int NCOORD>
__global__ void export_be_kernel(const Fq* __restrict__ pts, size_t n, uint8_t* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fq c[NCOORD];
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < NCOORD; k++) { c[k] = ld16(pts + i * NCOORD + k); for (int j = 0; j < 8; j++) any |= c[k].v[j]; }
    uint32_t* o = reinterpret_cast<uint32_t*>(out + i * NCOORD * 32);
    if (!any) { for (int j = 0; j < NCOORD * 8; j++) o[j] = 0; o[0] = 0x40u; return; }
#pragma unroll
    for (int k = 0; k < NCOORD; k++) {
        Fq v = fp_from_mont<FqParams>(c[k]);
        const int dk = NCOORD == 4 ? (k ^ 1) : k;          // G2: c1 before c0
#pragma unroll
        for (int j = 0; j < 8; j++) o[dk * 8 + j] = __byte_perm(v.v[7 - j], 0, 0x0123);
    }
}

static inline unsigned nblk(size_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

template <class F>
static void fixed_base_mul(Ctx* ctx, const Affine<F>& G, const Fr* d_scalars, size_t n, Affine<F>* d_out) {
    if (!n) return;
    DevBuf tabx((size_t)FB_WINDOWS * FB_ENTRIES * sizeof(XYZZ<F>)), tab((size_t)FB_WINDOWS * FB_ENTRIES * sizeof(Affine<F>));
    fixed_table_kernel<F><<<1, 32, 0, ctx->stream>>>(tabx.as<XYZZ<F>>(), G);
    xyzz_normalise<F>(ctx, tabx.as<XYZZ<F>>(), tab.as<Affine<F>>(), (size_t)FB_WINDOWS * FB_ENTRIES);
    DevBuf acc(n * sizeof(XYZZ<F>));
    fixed_mul_kernel<F><<<nblk(n, 128), 128, 0, ctx->stream>>>(tab.as<Affine<F>>(), d_scalars, n, acc.as<XYZZ<F>>());
    ctx->launches += 2;
    ZA_CUDA(cudaGetLastError());
    xyzz_normalise<F>(ctx, acc.as<XYZZ<F>>(), d_out, n);
    ZA_CUDA(cudaStreamSynchronize(ctx->stream));
}

static Fr fr_from_le(const uint8_t* p, const char* what) {
    Fr c; memcpy(c.v, p, 32);
    if (!fp_is_canonical<FrParams>(c.v)) throw ZaError(ZA_ERR_NOT_CANONICAL, std::string(what) + " is not a canonical Fr element");
    return fp_to_mont<FrParams>(c);
}
static Fq fq_from_le2(const uint8_t* p) { Fq c; memcpy(c.v, p, 32); return fp_to_mont<FqParams>(c); }
static void put_be32(std::vector<uint8_t>& o, uint32_t v) { o.push_back(v >> 24); o.push_back(v >> 16); o.push_back(v >> 8); o.push_back(v); }

template <class F, int NCOORD>
static void append_points(Ctx* ctx, std::vector<uint8_t>& out, const Affine<F>* d_pts, size_t n, bool with_count, bool filter_inf, bool* saw_inf) {
    std::vector<uint8_t> rec(n * NCOORD * 32 + 16);
    if (n) {
        DevBuf d(n * NCOORD * 32);
        export_be_kernel<NCOORD><<<nblk(n, 128), 128, 0, ctx->stream>>>(reinterpret_cast<const Fq*>(d_pts), n, d.as<uint8_t>());
        ctx->launches++;
        ZA_CUDA(cudaGetLastError());
        ZA_CUDA(cudaMemcpyAsync(rec.data(), d.p, n * NCOORD * 32, cudaMemcpyDeviceToHost, ctx->stream));
        ZA_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    size_t kept = 0;
    for (size_t i = 0; i < n; i++) {
        bool inf = (rec[i * NCOORD * 32] & 0x40) != 0;
        if (inf && saw_inf) *saw_inf = true;
        if (inf && filter_inf) continue;
        if (kept != i) memmove(rec.data() + kept * NCOORD * 32, rec.data() + i * NCOORD * 32, NCOORD * 32);
        kept++;
    }
    if (with_count) put_be32(out, (uint32_t)kept);
    out.insert(out.end(), rec.begin(), rec.begin() + kept * NCOORD * 32);
}

static std::vector<uint8_t> generate_parameters(Ctx* ctx, const Circuit* c, const uint8_t* alpha_b, const uint8_t* beta_b, const uint8_t* gamma_b,
                                                const uint8_t* delta_b, const uint8_t* tau_b, const uint8_t* g1_b, const uint8_t* g2_b) {
    cudaStream_t st = ctx->stream;
    const uint32_t ni = c->ni, na = c->na, nc = c->nc, nv = ni + na;
    const Fr alpha = fr_from_le(alpha_b, "alpha"), beta = fr_from_le(beta_b, "beta"), gamma = fr_from_le(gamma_b, "gamma"),
             delta = fr_from_le(delta_b, "delta"), tau = fr_from_le(tau_b, "tau");
    if (gamma.is_zero() || delta.is_zero()) throw ZaError(ZA_ERR_INVALID, "setup: gamma and delta must be non-zero");
    G1Affine g1; G2Affine g2;
    g1.x = fq_from_le2(g1_b); g1.y = fq_from_le2(g1_b + 32);
    g2.x.c0 = fq_from_le2(g2_b); g2.x.c1 = fq_from_le2(g2_b + 32); g2.y.c0 = fq_from_le2(g2_b + 64); g2.y.c1 = fq_from_le2(g2_b + 96);
    if (!affine_on_curve<Fq>(g1, host_g1_b()) || !affine_on_curve<Fq2>(g2, host_g2_b())) throw ZaError(ZA_ERR_NOT_ON_CURVE, "setup: generator not on the curve");
    const size_t len = (size_t)nc + ni;
    size_t m = 1; int log_m = 0;
    while (m < len) { m *= 2; log_m++; if (log_m >= 28) throw ZaError(ZA_ERR_POLY_DEGREE_TOO_LARGE, "setup: domain of 2^28 or more elements"); }

    // powers of tau, h scalars, Lagrange coefficients
    DevBuf pw(m * sizeof(Fr)), hs(m * sizeof(Fr));
    launch_powers_public(ctx, pw.as<Fr>(), m, tau, Fr::one());      // tau^i
    uint32_t me[8] = {(uint32_t)m, (uint32_t)((uint64_t)m >> 32), 0, 0, 0, 0, 0, 0};
    const Fr z = fp_pow<FrParams>(tau, me) - Fr::one();
    const Fr delta_inv = inv(delta), gamma_inv = inv(gamma);
    if (m > 1) {
        fr_scale_const_kernel<<<nblk(m - 1, 256), 256, 0, st>>>(pw.as<Fr>(), hs.as<Fr>(), m - 1, z * delta_inv);
        ctx->launches++;
    }
    ntt_mode(ctx, pw.as<Fr>(), log_m, ZA_NTT_IFFT, 1);                 // pw now holds L_k(tau)

    // column views and the per-variable evaluations
    DevBuf evals((size_t)4 * nv * sizeof(Fr));
    Fr* d_at = evals.as<Fr>(); Fr* d_bt = d_at + nv; Fr* d_ct = d_bt + nv; Fr* d_ext = d_ct + nv;
    for (int w = 0; w < 3; w++) {
        const uint32_t nt = c->h_ptr[w][nc];
        std::vector<uint32_t> colptr(nv + 1, 0), term(nt ? nt : 1), row(nt ? nt : 1);
        for (uint32_t t = 0; t < nt; t++) colptr[c->h_col[w][t] + 1]++;
        for (uint32_t v = 0; v < nv; v++) colptr[v + 1] += colptr[v];
        std::vector<uint32_t> cur(colptr.begin(), colptr.end() - 1);
        for (uint32_t k = 0; k < nc; k++)
            for (uint32_t t = c->h_ptr[w][k]; t < c->h_ptr[w][k + 1]; t++) { uint32_t p = cur[c->h_col[w][t]]++; term[p] = t; row[p] = k; }
        DevBuf d_colptr((nv + 1) * 4), d_term((size_t)(nt ? nt : 1) * 4), d_row((size_t)(nt ? nt : 1) * 4);
        ZA_CUDA(cudaMemcpyAsync(d_colptr.p, colptr.data(), (nv + 1) * 4, cudaMemcpyHostToDevice, st));
        ZA_CUDA(cudaMemcpyAsync(d_term.p, term.data(), (size_t)(nt ? nt : 1) * 4, cudaMemcpyHostToDevice, st));
        ZA_CUDA(cudaMemcpyAsync(d_row.p, row.data(), (size_t)(nt ? nt : 1) * 4, cudaMemcpyHostToDevice, st));
        qap_column_kernel<<<nblk(nv, 128), 128, 0, st>>>(d_colptr.as<uint32_t>(), d_term.as<uint32_t>(), d_row.as<uint32_t>(), c->coeff[w].as<Fr>(),
                                                        pw.as<Fr>(), nv, w == 0 ? d_at : w == 1 ? d_bt : d_ct);
        ctx->launches++;
        ZA_CUDA(cudaGetLastError());
        ZA_CUDA(cudaStreamSynchronize(st));       // host vectors go out of scope
    }
    qap_ext_kernel<<<nblk(nv, 128), 128, 0, st>>>(d_at, d_bt, d_ct, pw.as<Fr>(), ni, nv, nc, alpha, beta, gamma_inv, delta_inv, d_ext);
    ctx->launches++;
    ZA_CUDA(cudaGetLastError());

    // the fixed-base scalar multiplications
    DevBuf h_pts((m ? m : 1) * sizeof(G1Affine)), a_pts((size_t)nv * sizeof(G1Affine)), b1_pts((size_t)nv * sizeof(G1Affine)),
        ext_pts((size_t)nv * sizeof(G1Affine)), b2_pts((size_t)nv * sizeof(G2Affine));
    fixed_base_mul<Fq>(ctx, g1, hs.as<Fr>(), m - 1, h_pts.as<G1Affine>());
    fixed_base_mul<Fq>(ctx, g1, d_at, nv, a_pts.as<G1Affine>());
    fixed_base_mul<Fq>(ctx, g1, d_bt, nv, b1_pts.as<G1Affine>());
    fixed_base_mul<Fq>(ctx, g1, d_ext, nv, ext_pts.as<G1Affine>());
    fixed_base_mul<Fq2>(ctx, g2, d_bt, nv, b2_pts.as<G2Affine>());

    // the verifying key (host) and the byte stream
    auto h1 = [&](const Fr& k) { Fr kc = fp_from_mont<FrParams>(k); return xyzz_to_affine<Fq>(xyzz_mul<Fq>(G1XYZZ::from_affine(g1), kc.v)); };
    auto h2 = [&](const Fr& k) { Fr kc = fp_from_mont<FrParams>(k); return xyzz_to_affine<Fq2>(xyzz_mul<Fq2>(G2XYZZ::from_affine(g2), kc.v)); };
    G1Affine vk1[3] = {h1(alpha), h1(beta), h1(delta)};
    G2Affine vk2[3] = {h2(beta), h2(gamma), h2(delta)};
    DevBuf d1(sizeof vk1), d2(sizeof vk2);
    ZA_CUDA(cudaMemcpy(d1.p, vk1, sizeof vk1, cudaMemcpyHostToDevice));
    ZA_CUDA(cudaMemcpy(d2.p, vk2, sizeof vk2, cudaMemcpyHostToDevice));
    std::vector<uint8_t> out;
    append_points<Fq, 2>(ctx, out, d1.as<G1Affine>(), 1, false, false, nullptr);            // alpha_g1
    append_points<Fq, 2>(ctx, out, d1.as<G1Affine>() + 1, 1, false, false, nullptr);        // beta_g1
    append_points<Fq2, 4>(ctx, out, d2.as<G2Affine>(), 1, false, false, nullptr);           // beta_g2
    append_points<Fq2, 4>(ctx, out, d2.as<G2Affine>() + 1, 1, false, false, nullptr);       // gamma_g2
    append_points<Fq, 2>(ctx, out, d1.as<G1Affine>() + 2, 1, false, false, nullptr);        // delta_g1
    append_points<Fq2, 4>(ctx, out, d2.as<G2Affine>() + 2, 1, false, false, nullptr);       // delta_g2
    append_points<Fq, 2>(ctx, out, ext_pts.as<G1Affine>(), ni, true, false, nullptr);       // ic
    append_points<Fq, 2>(ctx, out, h_pts.as<G1Affine>(), m - 1, true, false, nullptr);      // h
    bool l_inf = false;
    append_points<Fq, 2>(ctx, out, ext_pts.as<G1Affine>() + ni, na, true, false, &l_inf);   // l
    if (l_inf) throw ZaError(ZA_ERR_UNCONSTRAINED_VARIABLE, "setup: unconstrained variable (an L query element is the point at infinity)");
    append_points<Fq, 2>(ctx, out, a_pts.as<G1Affine>(), nv, true, true, nullptr);          // a  (infinity filtered out)
    append_points<Fq, 2>(ctx, out, b1_pts.as<G1Affine>(), nv, true, true, nullptr);         // b_g1
    append_points<Fq2, 4>(ctx, out, b2_pts.as<G2Affine>(), nv, true, true, nullptr);        // b_g2
    return out;
}

}  // namespace za

using namespace za;

extern "C" {

size_t za_parameters_max_size(const za_circuit* circuit) {
    if (!circuit) return 0;
    const Circuit* c = circuit->c.get();
    size_t len = (size_t)c->nc + c->ni, m = 1;
    while (m < len) m *= 2;
    size_t nv = (size_t)c->ni + c->na;
    return 576 + 4 + 64 * (size_t)c->ni + 4 + 64 * m + 4 + 64 * (size_t)c->na + 4 + 64 * nv + 4 + 64 * nv + 4 + 128 * nv;
}

int za_generate_parameters(za_ctx* ctx, const za_circuit* circuit, const uint8_t* alpha, const uint8_t* beta, const uint8_t* gamma,
                           const uint8_t* delta, const uint8_t* tau, const uint8_t* g1, const uint8_t* g2, uint8_t* out, size_t size, size_t* out_len) {
    if (!ctx || !circuit || !alpha || !beta || !gamma || !delta || !tau || !g1 || !g2 || !out || !out_len) return fail(ZA_ERR_INVALID, "NULL argument");
    try {
        ZA_CUDA(cudaSetDevice(ctx->c.device));
        std::vector<uint8_t> blob = generate_parameters(&ctx->c, circuit->c.get(), alpha, beta, gamma, delta, tau, g1, g2);
        *out_len = blob.size();
        if (blob.size() > size) return fail(ZA_ERR_BUFFER_TOO_SMALL, "Parameters need %zu bytes", blob.size());
        memcpy(out, blob.data(), blob.size());
        return ZA_OK;
    }
    catch (const ZaError& e) { return fail(e.code, "%s", e.what()); }
    catch (const CudaError& e) { return fail(ZA_ERR_CUDA, "%s", e.what()); }
    catch (const std::exception& e) { return fail(ZA_ERR_INVALID, "%s", e.what()); }
}

}  // extern "C"
