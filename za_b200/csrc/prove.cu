// Groth16 create_proof on the GPU, plus the proving-key / circuit / multiexp handles of the C ABI.
//
// Replaces (un-vendored) bellman_ce groth16/prover.rs `create_proof` + `ProvingAssignment`, the
// `ParameterSource for &Parameters` cursor logic, and `Parameters::read` — everything
// /root/reference/prover/src/groth16/prover.rs:173 (`create_random_proof`) and format.rs:285
// (`Parameters::read(pk, true)`) call into.  Step numbers refer to SURVEY.md §3.2.
#include "common.cuh"
#include "api_internal.cuh"
#include "objects.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <memory>
#include <chrono>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <algorithm>

namespace za {

// ------------------------------------------------------------------ host <-> byte encodings
static Fq fq_from_le(const uint8_t* p) { Fq c; memcpy(c.v, p, 32); return fp_to_mont<FqParams>(c); }
static void fq_to_le(const Fq& a, uint8_t* p) { Fq c = fp_from_mont<FqParams>(a); memcpy(p, c.v, 32); }
static bool all_zero(const uint8_t* p, size_t n) { for (size_t i = 0; i < n; i++) if (p[i]) return false; return true; }

static G1Affine g1_from_le(const uint8_t* p) {
    G1Affine a;
    if (all_zero(p, 64)) return G1Affine::inf();
    a.x = fq_from_le(p); a.y = fq_from_le(p + 32);
    return a;
}
static G2Affine g2_from_le(const uint8_t* p) {
    G2Affine a;
    if (all_zero(p, 128)) return G2Affine::inf();
    a.x.c0 = fq_from_le(p); a.x.c1 = fq_from_le(p + 32); a.y.c0 = fq_from_le(p + 64); a.y.c1 = fq_from_le(p + 96);
    return a;
}
static void g1_to_le(const G1Affine& a, uint8_t* p) {
    if (a.is_inf()) { memset(p, 0, 64); return; }
    fq_to_le(a.x, p); fq_to_le(a.y, p + 32);
}
static void g2_to_le(const G2Affine& a, uint8_t* p) {
    if (a.is_inf()) { memset(p, 0, 128); return; }
    fq_to_le(a.x.c0, p); fq_to_le(a.x.c1, p + 32); fq_to_le(a.y.c0, p + 64); fq_to_le(a.y.c1, p + 96);
}
static void xyzz_to_le(const G1XYZZ& a, uint8_t* p) { fq_to_le(a.X, p); fq_to_le(a.Y, p + 32); fq_to_le(a.ZZ, p + 64); fq_to_le(a.ZZZ, p + 96); }
static void xyzz_to_le(const G2XYZZ& a, uint8_t* p) {
    const Fq* c = reinterpret_cast<const Fq*>(&a);
    for (int i = 0; i < 8; i++) fq_to_le(c[i], p + 32 * i);
}
static G1XYZZ g1_xyzz_from_le(const uint8_t* p) { G1XYZZ a; a.X = fq_from_le(p); a.Y = fq_from_le(p + 32); a.ZZ = fq_from_le(p + 64); a.ZZZ = fq_from_le(p + 96); return a; }
static G2XYZZ g2_xyzz_from_le(const uint8_t* p) {
    G2XYZZ a; Fq* c = reinterpret_cast<Fq*>(&a);
    for (int i = 0; i < 8; i++) c[i] = fq_from_le(p + 32 * i);
    return a;
}

// Window size of the fixed-base table: with one bucket space for all windows the bucket count is 2^(c-1)
// regardless of W, so c grows with log2(n) up to 20 (13 windows).  0 = no table (small or too large).
// Upper bound of ONE query's table: 64 GiB (a 2^26-point G1 query at c = 20 is 56 GB of the 180 GB) and never more than
// 45 % of the memory that is free when the table is built — the queries of a large key are built one after the other, so
// the later ones fall back to the table-free path by themselves once the card fills up (there used to be a fixed 24 GiB
// cap that dropped the table of every 2^26 query).  ZA_MSM_TABLE_CAP_GB overrides the 64.
static size_t table_cap_bytes() {
    size_t cap = (size_t)64 << 30;
    if (const char* e = getenv("ZA_MSM_TABLE_CAP_GB")) { long v = atol(e); if (v >= 0 && v <= 1024) cap = (size_t)v << 30; }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) cap = std::min(cap, (size_t)((double)free_b * 0.45));
    else cudaGetLastError();
    return cap;
}
static int table_window_bits(size_t n, size_t point_bytes) {
    if (const char* e = getenv("ZA_MSM_TABLE")) { int v = atoi(e); if (v == 0) return 0; if (v >= 8 && v <= 22) return v; }
    if (n < 4096) return 0;
    int lg = 0; while (((size_t)1 << (lg + 1)) <= n) lg++;
    // measured on B200 (scratch/sweep_c.py): 2^16 -> 16, 2^18 -> 17, 2^20 and up -> 20
    int c = lg >= 20 ? 20 : lg >= 18 ? 17 : lg >= 15 ? 16 : lg + 1;
    size_t W = (255 + c - 1) / c;
    if (n * W * point_bytes > table_cap_bytes()) return 0;
    return c;
}
// Build the table for bases [lo, lo + n) (default: the whole array); the window size follows n unless `c_forced` > 0.
static void bases_build_table(Ctx* ctx, Bases* b, size_t lo = 0, size_t n = (size_t)-1, int c_forced = 0) {
    if (n == (size_t)-1) n = b->n - lo;
    const size_t pb = b->group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
    b->table.release(); b->tab_c = 0; b->tab_W = 0; b->tab_lo = 0; b->tab_n = 0;
    int c = table_window_bits(n, pb);
    if (c_forced > 0 && n > 64 && n * (size_t)((255 + c_forced - 1) / c_forced) * pb <= table_cap_bytes()) c = c_forced;
    if (!c || b->has_infinity) return;
    const int W = (255 + c - 1) / c;
    b->table.alloc(n * (size_t)W * pb);
    if (b->group == 1) bases_table_build<Fq>(ctx, b->pts.as<G1Affine>() + lo, n, c, W, b->table.as<G1Affine>());
    else bases_table_build<Fq2>(ctx, b->pts.as<G2Affine>() + lo, n, c, W, b->table.as<G2Affine>());
    b->tab_c = c; b->tab_W = W; b->tab_lo = lo; b->tab_n = n;
}
// Tables of the four witness queries (L, A, B in G1, B in G2) over the given ranges.  When the ranges are of similar
// size the window is chosen JOINTLY (from the largest), so that B (G1), L and A can run as one multiexp with three bucket
// spaces (K4') and the two B multiexps can share a sort; ranges more than 4x apart keep their own windows (a small query
// should not pay for 2^19 buckets).
static void pk_build_witness_tables(Ctx* ctx, Pk* pk, const size_t* lo, const size_t* n) {      // order: l, a, b
    Bases* q[3] = {pk->l.get(), pk->a.get(), pk->b_g1.get()};
    size_t nmax = 0, nmin = (size_t)-1;
    for (int i = 0; i < 3; i++) { nmax = std::max(nmax, n[i]); nmin = std::min(nmin, n[i]); }
    int joint = 0;
    if (nmin > 64 && nmin * 4 >= nmax && !getenv("ZA_MSM_TABLE")) joint = table_window_bits(nmax, sizeof(G1Affine));
    for (int i = 0; i < 3; i++) bases_build_table(ctx, q[i], lo[i], n[i], joint);
    bases_build_table(ctx, pk->b_g2.get(), lo[2], n[2], joint);
}

template <class F>
static const Affine<F>* table_for(const Bases* b, size_t offset, size_t n) {
    if (!b->tab_c || n <= 64 || offset < b->tab_lo || offset + n > b->tab_lo + b->tab_n) return nullptr;
    return b->table.as<Affine<F>>() + (offset - b->tab_lo) * (size_t)b->tab_W;
}

static std::unique_ptr<Bases> bases_from_le(Ctx* ctx, int group, const uint8_t* le, size_t n, bool allow_infinity, const char* what) {
    std::unique_ptr<Bases> b(new Bases());
    b->ctx = ctx; b->group = group; b->n = n; b->has_infinity = false;
    const size_t sz = group == 1 ? 64 : 128;
    b->pts.alloc(n * sz);
    if (n) ZA_CUDA(cudaMemcpyAsync(b->pts.p, le, n * sz, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t flags = group == 1 ? bases_import<Fq>(ctx, b->pts.p, n) : bases_import<Fq2>(ctx, b->pts.p, n);
    char msg[160];
    if (flags & 1) { snprintf(msg, sizeof msg, "%s: coordinate not in canonical form (>= q)", what); throw ZaError(ZA_ERR_BAD_ENCODING, msg); }
    if (flags & 2) { snprintf(msg, sizeof msg, "%s: point not on the curve", what); throw ZaError(ZA_ERR_NOT_ON_CURVE, msg); }
    if (flags & 4) {
        if (!allow_infinity) { snprintf(msg, sizeof msg, "%s: point at infinity", what); throw ZaError(ZA_ERR_UNEXPECTED_IDENTITY, msg); }
        b->has_infinity = true;
    }
    return b;
}

// ------------------------------------------------------------------ multiexp front-ends
__global__ void gather_scalars_kernel(const uint4* __restrict__ src, const uint32_t* __restrict__ idx, size_t n, uint4* dst) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t s = idx[i];
    dst[2 * i] = __ldg(src + 2 * s);
    dst[2 * i + 1] = __ldg(src + 2 * s + 1);
}
static inline unsigned nblk(size_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

static void check_scalars_canonical(const uint8_t* p, size_t n, const char* what) {
    for (size_t i = 0; i < n; i++) {
        uint32_t w[8];
        memcpy(w, p + 32 * i, 32);
        if (!fp_is_canonical<FrParams>(w)) {
            char b[128];
            snprintf(b, sizeof b, "%s[%zu] is not a canonical Fr element (>= r)", what, i);
            throw ZaError(ZA_ERR_NOT_CANONICAL, b);
        }
    }
}

// "Every scalar < r" for vectors that are already on the device (the digit recoding and the Montgomery conversion both
// assume it): one pass on the context's stream, the verdict — index of the first element >= r, or ~0 — travels to
// pinned host memory behind it and is read by whoever synchronises next.  Slots: 0 witness, 1 h scalars, 2 multiexp.
__global__ void fr_canonical_check_kernel(const uint32_t* __restrict__ v, size_t n, unsigned long long* first_bad);
enum { CHK_WITNESS = 0, CHK_H = 1, CHK_MSM = 2 };
static void canonical_check_enqueue(Ctx* ctx, const void* d, size_t n, int which) {
    DevBuf& bad = ctx->scratch[14];
    bad.ensure(64);
    if (!ctx->host_flag) {
        ZA_CUDA(cudaHostAlloc((void**)&ctx->host_flag, 64, cudaHostAllocDefault));
        for (int i = 0; i < 8; i++) ctx->host_flag[i] = ~0ull;
    }
    unsigned long long* d_flag = bad.as<unsigned long long>() + which;
    ZA_CUDA(cudaMemsetAsync(d_flag, 0xff, 8, ctx->stream));
    if (n) {
        fr_canonical_check_kernel<<<nblk(n, 256), 256, 0, ctx->stream>>>((const uint32_t*)d, n, d_flag);
        ctx->launches++;
        ZA_CUDA(cudaGetLastError());
    }
    ZA_CUDA(cudaMemcpyAsync(ctx->host_flag + which, d_flag, 8, cudaMemcpyDeviceToHost, ctx->stream));
}
// after the stream has been synchronised past the copy above
static unsigned long long canonical_check_verdict(Ctx* ctx, int which) {
    if (!ctx->host_flag) return ~0ull;
    const unsigned long long v = ctx->host_flag[which];
    ctx->host_flag[which] = ~0ull;
    return v;
}

template <class F>
static XYZZ<F> multiexp_dev(Ctx* ctx, const Bases* b, size_t offset, const uint32_t* d_scalars, size_t n, bool check = false) {
    if (offset > b->n || n > b->n - offset)
        throw ZaError(ZA_ERR_IO, "multiexp: the base query is shorter than the exponent vector (bellman: unexpected EOF)");
    if (check) canonical_check_enqueue(ctx, d_scalars, n, CHK_MSM);
    XYZZ<F> r;
    try {
        msm_enqueue<F>(ctx, 0, b->pts.as<Affine<F>>() + offset, d_scalars, n, b->has_infinity, -1, table_for<F>(b, offset, n), b->tab_c, b->tab_W);
        r = msm_finish<F>(ctx, 0);
    } catch (...) { msm_abort(ctx); throw; }
    if (check) {
        ZA_CUDA(cudaStreamSynchronize(ctx->stream));
        const unsigned long long bad = canonical_check_verdict(ctx, CHK_MSM);
        if (bad != ~0ull) {
            char m[128];
            snprintf(m, sizeof m, "scalars[%llu] is not a canonical Fr element (>= r)", bad);
            throw ZaError(ZA_ERR_NOT_CANONICAL, m);
        }
    }
    return r;
}

// ------------------------------------------------------------------ proving key
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// pairing_ce G1Uncompressed/G2Uncompressed: 32-byte big-endian coordinates, bit 7 of byte 0 =
// compressed flag (must be clear), bit 6 = infinity.  Output: the LE interchange layout.
static void be_point_to_le(const uint8_t* src, uint8_t* dst, int ncoord, bool swap_pairs) {
    if (src[0] & 0x80) throw ZaError(ZA_ERR_BAD_ENCODING, "Parameters: compressed point where an uncompressed one is expected");
    if (src[0] & 0x40) {
        if ((src[0] & 0x3f) || !all_zero(src + 1, 32 * ncoord - 1)) throw ZaError(ZA_ERR_BAD_ENCODING, "Parameters: malformed point at infinity");
        memset(dst, 0, 32 * ncoord);
        return;
    }
    for (int k = 0; k < ncoord; k++) {
        // G2 is stored c1 || c0 per coordinate; interchange wants c0 || c1
        int dk = swap_pairs ? (k ^ 1) : k;
        for (int i = 0; i < 32; i++) dst[32 * dk + i] = src[32 * k + 31 - i];
    }
    // A finite point whose bytes are all zero would alias the infinity encoding; (0,0) is off-curve anyway.
    if (all_zero(dst, 32 * ncoord)) throw ZaError(ZA_ERR_NOT_ON_CURVE, "Parameters: point (0,0) is not on the curve");
}

struct Reader {
    const uint8_t* p; const uint8_t* end;
    void need(size_t n) { if ((size_t)(end - p) < n) throw ZaError(ZA_ERR_IO, "Parameters: unexpected end of stream"); }
    uint32_t u32() { need(4); uint32_t v = be32(p); p += 4; return v; }
};

__global__ void g2_subgroup_check_kernel(const G2Affine* __restrict__ pts, size_t n, uint32_t* flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G2Affine P;
    {
        const uint4* s = reinterpret_cast<const uint4*>(pts + i);
        uint4* d = reinterpret_cast<uint4*>(&P);
#pragma unroll
        for (int k = 0; k < 8; k++) d[k] = __ldg(s + k);
    }
    if (P.is_inf()) return;
    // r * P == infinity ?   (pairing_ce is_in_correct_subgroup_assuming_on_curve)
    G2XYZZ acc = G2XYZZ::inf();
    for (int b = 253; b >= 0; b--) {
        acc = xyzz_dbl<Fq2>(acc);
        if ((FrParams::mod(b >> 5) >> (b & 31)) & 1u) xyzz_madd<Fq2>(acc, P.x, P.y, false);
    }
    if (!acc.is_inf()) atomicOr(flag, 1u);
}

static void g2_subgroup_check(Ctx* ctx, const Bases* b, const char* what) {
    if (!b->n) return;
    DevBuf flag(4);
    ZA_CUDA(cudaMemsetAsync(flag.p, 0, 4, ctx->stream));
    g2_subgroup_check_kernel<<<nblk(b->n, 64), 64, 0, ctx->stream>>>(b->pts.as<G2Affine>(), b->n, flag.as<uint32_t>());
    ctx->launches++;
    ZA_CUDA(cudaGetLastError());
    uint32_t h = 0;
    ZA_CUDA(cudaMemcpyAsync(&h, flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ZA_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h) { char m[128]; snprintf(m, sizeof m, "%s: point not in the r-torsion subgroup", what); throw ZaError(ZA_ERR_NOT_IN_SUBGROUP, m); }
}

static std::unique_ptr<Bases> read_query(Ctx* ctx, Reader& r, int group, bool checked, const char* what) {
    uint32_t n = r.u32();
    const size_t sz = group == 1 ? 64 : 128;
    r.need((size_t)n * sz);
    std::vector<uint8_t> le((size_t)n * sz + 1);
    for (uint32_t i = 0; i < n; i++) be_point_to_le(r.p + (size_t)i * sz, le.data() + (size_t)i * sz, group == 1 ? 2 : 4, group == 2);
    r.p += (size_t)n * sz;
    // Parameters::read rejects points at infinity in every query
    std::unique_ptr<Bases> b = bases_from_le(ctx, group, le.data(), n, false, what);
    if (checked && group == 2) g2_subgroup_check(ctx, b.get(), what);
    return b;
}

static std::unique_ptr<Pk> pk_load(Ctx* ctx, const uint8_t* data, size_t len, bool checked) {
    std::unique_ptr<Pk> pk(new Pk());
    pk->ctx = ctx;
    Reader r{data, data + len};
    // VerifyingKey::read: alpha_g1, beta_g1, beta_g2, gamma_g2, delta_g1, delta_g2, ic.  The vk points go
    // through the same import kernel (curve check on the GPU), then come back to the host for the assembly.
    uint8_t le[6 * 128];   // [0,192): alpha_g1, beta_g1, delta_g1   [256,640): beta_g2, gamma_g2, delta_g2
    {
        r.need(64 + 64 + 128 + 128 + 64 + 128);
        const uint8_t* p = r.p;
        be_point_to_le(p, le, 2, false); p += 64;            // alpha_g1
        be_point_to_le(p, le + 64, 2, false); p += 64;       // beta_g1
        be_point_to_le(p, le + 256, 4, true); p += 128;      // beta_g2
        be_point_to_le(p, le + 384, 4, true); p += 128;      // gamma_g2
        be_point_to_le(p, le + 128, 2, false); p += 64;      // delta_g1
        be_point_to_le(p, le + 512, 4, true); p += 128;      // delta_g2
        r.p = p;
    }
    {
        std::unique_ptr<Bases> v1 = bases_from_le(ctx, 1, le, 3, true, "vk G1");
        std::unique_ptr<Bases> v2 = bases_from_le(ctx, 2, le + 256, 3, true, "vk G2");
        if (checked) g2_subgroup_check(ctx, v2.get(), "vk G2");
        G1Affine h1[3]; G2Affine h2[3];
        ZA_CUDA(cudaMemcpy(h1, v1->pts.p, sizeof h1, cudaMemcpyDeviceToHost));
        ZA_CUDA(cudaMemcpy(h2, v2->pts.p, sizeof h2, cudaMemcpyDeviceToHost));
        pk->alpha_g1 = h1[0]; pk->beta_g1 = h1[1]; pk->delta_g1 = h1[2];
        pk->beta_g2 = h2[0]; pk->gamma_g2 = h2[1]; pk->delta_g2 = h2[2];
    }
    {
        uint32_t n = r.u32();
        r.need((size_t)n * 64);
        std::vector<uint8_t> icle((size_t)n * 64 + 1);
        for (uint32_t i = 0; i < n; i++) be_point_to_le(r.p + (size_t)i * 64, icle.data() + (size_t)i * 64, 2, false);
        r.p += (size_t)n * 64;
        std::unique_ptr<Bases> v = bases_from_le(ctx, 1, icle.data(), n, true, "vk ic");
        pk->ic.resize(n);
        if (n) ZA_CUDA(cudaMemcpy(pk->ic.data(), v->pts.p, (size_t)n * sizeof(G1Affine), cudaMemcpyDeviceToHost));
    }
    pk->h = read_query(ctx, r, 1, checked, "h query");
    pk->l = read_query(ctx, r, 1, checked, "l query");
    pk->a = read_query(ctx, r, 1, checked, "a query");
    pk->b_g1 = read_query(ctx, r, 1, checked, "b_g1 query");
    pk->b_g2 = read_query(ctx, r, 2, checked, "b_g2 query");
    bases_build_table(ctx, pk->h.get());
    const size_t lo[3] = {0, 0, 0}, n[3] = {pk->l->n, pk->a->n, pk->b_g1->n};
    pk_build_witness_tables(ctx, pk.get(), lo, n);
    return pk;
}

// ------------------------------------------------------------------ circuit (R1CS on device)
// Everything that depends only on the constraint system is computed once here: the CSR matrices in
// Montgomery form, the three density maps of bellman's ProvingAssignment (they depend on which
// variables occur in A / B rows, never on the witness: density.inc(i) fires for every term)
// and the compacted index lists the density-filtered multiexps need (K10).
__global__ void circuit_coeff_import_kernel(Fr* c, size_t n, uint32_t* flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v;
    {
        const uint4* s = reinterpret_cast<const uint4*>(c + i);
        uint4 a = s[0], b = s[1];
        v.v[0] = a.x; v.v[1] = a.y; v.v[2] = a.z; v.v[3] = a.w; v.v[4] = b.x; v.v[5] = b.y; v.v[6] = b.z; v.v[7] = b.w;
    }
    if (!fp_is_canonical<FrParams>(v.v)) atomicOr(flag, 1u);
    v = fp_to_mont<FrParams>(v);
    uint4* d = reinterpret_cast<uint4*>(c + i);
    d[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    d[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}

// ProvingAssignment::enforce -> eval: one thread per constraint row, out[row] = sum coeff * w[col]
__global__ void r1cs_eval_kernel(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ col, const Fr* __restrict__ coeff,
                                 const Fr* __restrict__ w, uint32_t nrows, Fr* out) {
    uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    Fr acc = Fr::zero();
    for (uint32_t t = ptr[row]; t < ptr[row + 1]; t++) {
        const uint4* cp = reinterpret_cast<const uint4*>(coeff + t);
        const uint4* wp = reinterpret_cast<const uint4*>(w + col[t]);
        uint4 c0 = __ldg(cp), c1 = __ldg(cp + 1), w0 = __ldg(wp), w1 = __ldg(wp + 1);
        Fr c, v;
        c.v[0] = c0.x; c.v[1] = c0.y; c.v[2] = c0.z; c.v[3] = c0.w; c.v[4] = c1.x; c.v[5] = c1.y; c.v[6] = c1.z; c.v[7] = c1.w;
        v.v[0] = w0.x; v.v[1] = w0.y; v.v[2] = w0.z; v.v[3] = w0.w; v.v[4] = w1.x; v.v[5] = w1.y; v.v[6] = w1.z; v.v[7] = w1.w;
        acc = acc + c * v;
    }
    uint4* d = reinterpret_cast<uint4*>(out + row);
    d[0] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
    d[1] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
}

// constraints.satisfies_with_signals (compiler/src/types/constraint.rs:29-67, called at prover.rs:155): row k holds
// iff A_k(w) * B_k(w) == C_k(w).  first_bad receives the smallest failing row (atomicMin).
__global__ void r1cs_check_kernel(const Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ c, uint32_t nrows, uint32_t* first_bad) {
    uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    Fr x, y, z;
    const uint4* pa = reinterpret_cast<const uint4*>(a + row);
    const uint4* pb = reinterpret_cast<const uint4*>(b + row);
    const uint4* pc = reinterpret_cast<const uint4*>(c + row);
    uint4 a0 = pa[0], a1 = pa[1], b0 = pb[0], b1 = pb[1], c0 = pc[0], c1 = pc[1];
    x.v[0] = a0.x; x.v[1] = a0.y; x.v[2] = a0.z; x.v[3] = a0.w; x.v[4] = a1.x; x.v[5] = a1.y; x.v[6] = a1.z; x.v[7] = a1.w;
    y.v[0] = b0.x; y.v[1] = b0.y; y.v[2] = b0.z; y.v[3] = b0.w; y.v[4] = b1.x; y.v[5] = b1.y; y.v[6] = b1.z; y.v[7] = b1.w;
    z.v[0] = c0.x; z.v[1] = c0.y; z.v[2] = c0.z; z.v[3] = c0.w; z.v[4] = c1.x; z.v[5] = c1.y; z.v[6] = c1.z; z.v[7] = c1.w;
    if (x * y != z) atomicMin(first_bad, row);
}

static std::unique_ptr<Circuit> circuit_upload(Ctx* ctx, const za_r1cs* cs) {
    std::unique_ptr<Circuit> c(new Circuit());
    c->ctx = ctx; c->ni = cs->num_inputs; c->na = cs->num_aux; c->nc = cs->num_constraints;
    if (c->ni == 0) throw ZaError(ZA_ERR_INVALID, "circuit: num_inputs must include the constant `one`");
    c->a_aux_density.assign(c->na, 0); c->b_in_density.assign(c->ni, 0); c->b_aux_density.assign(c->na, 0);
    DevBuf flag(4);
    ZA_CUDA(cudaMemsetAsync(flag.p, 0, 4, ctx->stream));
    for (int w = 0; w < 3; w++) {
        const uint32_t* ptr = cs->ptr[w];
        if (!ptr) throw ZaError(ZA_ERR_INVALID, "circuit: NULL row pointer array");
        if (ptr[0] != 0) throw ZaError(ZA_ERR_INVALID, "circuit: ptr[0] must be 0");
        for (uint32_t k = 0; k < c->nc; k++) if (ptr[k + 1] < ptr[k]) throw ZaError(ZA_ERR_INVALID, "circuit: row offsets must be non-decreasing");
        const uint32_t nt = ptr[c->nc];
        std::vector<uint32_t> col(nt ? nt : 1);
        for (uint32_t t = 0; t < nt; t++) {
            uint32_t v = cs->var[w][t];
            if (v & ZA_VAR_AUX) {
                uint32_t i = v & ~ZA_VAR_AUX;
                if (i >= c->na) throw ZaError(ZA_ERR_INVALID, "circuit: aux variable index out of range");
                col[t] = c->ni + i;
                if (w == 0) c->a_aux_density[i] = 1;
                if (w == 1) c->b_aux_density[i] = 1;
            } else {
                if (v >= c->ni) throw ZaError(ZA_ERR_INVALID, "circuit: input variable index out of range");
                col[t] = v;
                if (w == 1) c->b_in_density[v] = 1;
            }
        }
        c->h_ptr[w].assign(ptr, ptr + c->nc + 1);
        c->h_col[w].assign(col.begin(), col.begin() + nt);
        c->ptr[w].alloc(((size_t)c->nc + 1) * 4);
        c->col[w].alloc((size_t)nt * 4);
        c->coeff[w].alloc((size_t)nt * 32);
        ZA_CUDA(cudaMemcpyAsync(c->ptr[w].p, ptr, ((size_t)c->nc + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (nt) {
            ZA_CUDA(cudaMemcpyAsync(c->col[w].p, col.data(), (size_t)nt * 4, cudaMemcpyHostToDevice, ctx->stream));
            ZA_CUDA(cudaMemcpyAsync(c->coeff[w].p, cs->coeff[w], (size_t)nt * 32, cudaMemcpyHostToDevice, ctx->stream));
            circuit_coeff_import_kernel<<<nblk(nt, 256), 256, 0, ctx->stream>>>(c->coeff[w].as<Fr>(), nt, flag.as<uint32_t>());
            ctx->launches++;
        }
        ZA_CUDA(cudaStreamSynchronize(ctx->stream));  // `col` is a local
    }
    uint32_t h = 0;
    ZA_CUDA(cudaMemcpy(&h, flag.p, 4, cudaMemcpyDeviceToHost));
    if (h) throw ZaError(ZA_ERR_NOT_CANONICAL, "circuit: a coefficient is not a canonical Fr element");
    auto make_idx = [&](const std::vector<uint8_t>& dens, DevBuf& out, uint32_t& total) {
        std::vector<uint32_t> idx;
        for (uint32_t i = 0; i < dens.size(); i++) if (dens[i]) idx.push_back(i);
        total = (uint32_t)idx.size();
        out.alloc((size_t)total * 4);
        if (total) {
            // on the context's (non-blocking) stream, like every kernel that reads the list; `idx` is a local
            ZA_CUDA(cudaMemcpyAsync(out.p, idx.data(), (size_t)total * 4, cudaMemcpyHostToDevice, ctx->stream));
            ZA_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    };
    make_idx(c->a_aux_density, c->a_aux_idx, c->a_aux_total);
    make_idx(c->b_in_density, c->b_in_idx, c->b_in_total);
    make_idx(c->b_aux_density, c->b_aux_idx, c->b_aux_total);
    {
        std::vector<uint32_t> a_cat, b_cat;
        for (uint32_t i = 0; i < c->ni; i++) a_cat.push_back(i);
        for (uint32_t i = 0; i < c->na; i++) if (c->a_aux_density[i]) a_cat.push_back(c->ni + i);
        for (uint32_t i = 0; i < c->ni; i++) if (c->b_in_density[i]) b_cat.push_back(i);
        for (uint32_t i = 0; i < c->na; i++) if (c->b_aux_density[i]) b_cat.push_back(c->ni + i);
        c->a_cat_total = (uint32_t)a_cat.size(); c->b_cat_total = (uint32_t)b_cat.size();
        c->a_cat_idx.alloc((size_t)c->a_cat_total * 4); c->b_cat_idx.alloc((size_t)c->b_cat_total * 4);
        if (c->a_cat_total) ZA_CUDA(cudaMemcpyAsync(c->a_cat_idx.p, a_cat.data(), (size_t)c->a_cat_total * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (c->b_cat_total) ZA_CUDA(cudaMemcpyAsync(c->b_cat_idx.p, b_cat.data(), (size_t)c->b_cat_total * 4, cudaMemcpyHostToDevice, ctx->stream));
        ZA_CUDA(cudaStreamSynchronize(ctx->stream));
        c->h_a_cat.swap(a_cat); c->h_b_cat.swap(b_cat);      // host copies: a rank's witness span follows from its point range
    }
    return c;
}

// ------------------------------------------------------------------ create_proof, in three stages
// Stage 1 (one GPU): witness -> a, b, c -> H coefficients.        SURVEY §3.2 steps 2-4
// Stage 2 (every GPU): the eight multiexps over this rank's point range.   steps 4-5, SURVEY §8e
// Stage 3 (host): add the per-rank partial sums and assemble A, B, C.       steps 6-8
static const uint32_t* gather(Ctx* ctx, const uint8_t* d_src, const DevBuf& idx, uint32_t total, DevBuf& dst) {
    dst.ensure((size_t)total * 32);
    if (total) {
        gather_scalars_kernel<<<nblk(total, 256), 256, 0, ctx->stream>>>((const uint4*)d_src, idx.as<uint32_t>(), total, dst.as<uint4>());
        ctx->launches++;
        ZA_CUDA(cudaGetLastError());
    }
    return dst.as<uint32_t>();
}
// the same for the entries [lo, hi) of the index list only (a rank's point range): dst keeps the layout of the whole
// exponent vector, so entry i is at dst + 32 i, and only the witness positions idx[lo..hi) are read
static const uint32_t* gather_range(Ctx* ctx, const uint8_t* d_src, const DevBuf& idx, uint32_t total, size_t lo, size_t hi, DevBuf& dst) {
    dst.ensure((size_t)total * 32);
    if (hi > lo) {
        gather_scalars_kernel<<<nblk(hi - lo, 256), 256, 0, ctx->stream>>>((const uint4*)d_src, idx.as<uint32_t>() + lo, hi - lo, dst.as<uint4>() + 2 * lo);
        ctx->launches++;
        ZA_CUDA(cudaGetLastError());
    }
    return dst.as<uint32_t>();
}

static size_t domain_size(const Circuit* c, int* log_m_out) {
    const size_t len = (size_t)c->nc + c->ni;      // rows incl. the input-consistency rows (step 3)
    size_t m = 1; int log_m = 0;
    while (m < len) { m *= 2; log_m++; if (log_m >= 28) throw ZaError(ZA_ERR_POLY_DEGREE_TOO_LARGE, "create_proof: domain of 2^28 or more elements"); }
    if (log_m_out) *log_m_out = log_m;
    return m;
}

// d_wit: [inputs | aux] canonical on device.  d_h: m scalars; on return the first m-1 are the canonical h coefficients.
static void prove_h(Ctx* ctx, const Circuit* c, const uint8_t* d_wit, Fr* d_h, za_trace* tr) {
    cudaStream_t st = ctx->stream;
    const uint32_t ni = c->ni, na = c->na, nc = c->nc;
    int log_m; const size_t m = domain_size(c, &log_m);
    const size_t len = (size_t)nc + ni;
    DevBuf& wit_mont = ctx->scratch[7];
    wit_mont.ensure(((size_t)ni + na) * 32 + 3 * m * sizeof(Fr));
    Fr* d_w = wit_mont.as<Fr>();
    // a, b, c lie m elements apart: their transforms run as one batch (h_poly_device); the h scalars go to d_h
    Fr* d_a = d_w + ni + na; Fr* d_b = d_a + m; Fr* d_c = d_b + m;
    ZA_CUDA(cudaMemcpyAsync(d_w, d_wit, ((size_t)ni + na) * 32, cudaMemcpyDeviceToDevice, st));
    fr_convert(ctx, d_w, (size_t)ni + na, 0);
    // steps 2-3: a, b, c = <row, witness>; then rows a = input_i, b = c = 0
    // rows the constraint evaluation does not write: the padding behind the nc + ni rows of a and the nc rows of b and c
    ZA_CUDA(cudaMemsetAsync(d_a + len, 0, (m - len) * sizeof(Fr), st));
    ZA_CUDA(cudaMemsetAsync(d_b + nc, 0, (m - nc) * sizeof(Fr), st));
    ZA_CUDA(cudaMemsetAsync(d_c + nc, 0, (m - nc) * sizeof(Fr), st));
    Fr* outs[3] = {d_a, d_b, d_c};
    if (nc) {
        ProfScope prof(ctx, PROF_R1CS, 3.0 * nc);
        for (int w = 0; w < 3; w++) {
            r1cs_eval_kernel<<<nblk(nc, 128), 128, 0, st>>>(c->ptr[w].as<uint32_t>(), c->col[w].as<uint32_t>(), c->coeff[w].as<Fr>(), d_w, nc, outs[w]);
            ctx->launches++;
        }
        ZA_CUDA(cudaGetLastError());
    }
    ZA_CUDA(cudaMemcpyAsync(d_a + nc, d_w, (size_t)ni * 32, cudaMemcpyDeviceToDevice, st));
    if (tr && (tr->a_eval || tr->b_eval || tr->c_eval)) {
        DevBuf tmp(len * 32);
        uint8_t* dst[3] = {tr->a_eval, tr->b_eval, tr->c_eval};
        for (int w = 0; w < 3; w++) {
            if (!dst[w]) continue;
            ZA_CUDA(cudaMemcpyAsync(tmp.p, outs[w], len * 32, cudaMemcpyDeviceToDevice, st));
            fr_convert(ctx, tmp.as<Fr>(), len, 1);
            ZA_CUDA(cudaMemcpyAsync(dst[w], tmp.p, len * 32, cudaMemcpyDeviceToHost, st));
            ZA_CUDA(cudaStreamSynchronize(st));
        }
    }
    if (tr) {
        if (tr->a_aux_density && na) memcpy(tr->a_aux_density, c->a_aux_density.data(), na);
        if (tr->b_input_density) memcpy(tr->b_input_density, c->b_in_density.data(), ni);
        if (tr->b_aux_density && na) memcpy(tr->b_aux_density, c->b_aux_density.data(), na);
    }
    // step 4
    h_poly_device(ctx, d_a, d_b, d_c, log_m, d_h);
    if (tr && tr->h_coeffs && m > 1) {
        ZA_CUDA(cudaMemcpyAsync(tr->h_coeffs, d_h, (m - 1) * 32, cudaMemcpyDeviceToHost, st));
        ZA_CUDA(cudaStreamSynchronize(st));
    }
}

struct Partials {
    G1XYZZ g1[6];   // h, l, a_inputs, a_aux, b1_inputs, b1_aux
    G2XYZZ g2[2];   // b2_inputs, b2_aux
};
static const size_t PARTIALS_BYTES = 6 * 128 + 2 * 256;
static void partials_to_le(const Partials& p, uint8_t* out) {
    for (int i = 0; i < 6; i++) xyzz_to_le(p.g1[i], out + 128 * i);
    for (int i = 0; i < 2; i++) xyzz_to_le(p.g2[i], out + 768 + 256 * i);
}
static void partials_add_le(Partials& acc, const uint8_t* in) {
    for (int i = 0; i < 6; i++) xyzz_add<Fq>(acc.g1[i], g1_xyzz_from_le(in + 128 * i));
    for (int i = 0; i < 2; i++) xyzz_add<Fq2>(acc.g2[i], g2_xyzz_from_le(in + 768 + 256 * i));
}
static Partials partials_zero() {
    Partials p;
    for (int i = 0; i < 6; i++) p.g1[i] = G1XYZZ::inf();
    for (int i = 0; i < 2; i++) p.g2[i] = G2XYZZ::inf();
    return p;
}

// contiguous share of `cnt` items for `rank` of `world`
static inline void share(size_t cnt, int rank, int world, size_t& lo, size_t& hi) {
    lo = (size_t)(((unsigned __int128)cnt * (unsigned)rank) / (unsigned)world);
    hi = (size_t)(((unsigned __int128)cnt * (unsigned)(rank + 1)) / (unsigned)world);
}
// the same with rank 0 weighted w0 (per mille of an ordinary rank's share): rank 0 also runs the H-polynomial
// pipeline, so it takes a smaller part of the witness multiexps (za_pk_partition_weighted)
static inline void share_weighted(size_t cnt, int rank, int world, uint32_t w0, size_t& lo, size_t& hi) {
    if (world == 1 || w0 == 1000) { share(cnt, rank, world, lo, hi); return; }
    const unsigned __int128 total = (unsigned __int128)w0 + 1000u * (unsigned)(world - 1);
    auto bound = [&](int k) -> size_t {
        if (k <= 0) return 0;
        if (k >= world) return cnt;
        return (size_t)(((unsigned __int128)cnt * ((unsigned __int128)w0 + 1000u * (unsigned)(k - 1))) / total);
    };
    lo = bound(rank);
    hi = bound(rank + 1);
}

template <class F>
static void multiexp_enqueue(Ctx* ctx, int slot, const Bases* b, size_t offset, const uint32_t* d_scalars, size_t n, int share_sort = -1) {
    if (offset > b->n || n > b->n - offset)
        throw ZaError(ZA_ERR_IO, "multiexp: the base query is shorter than the exponent vector (bellman: unexpected EOF)");
    msm_enqueue<F>(ctx, slot, b->pts.as<Affine<F>>() + offset, d_scalars, n, b->has_infinity, share_sort, table_for<F>(b, offset, n), b->tab_c, b->tab_W);
}

static void check_query_lengths(const Pk* pk, const Circuit* c, size_t m) {
    // the whole-query length checks bellman's cursors would trip over (unexpected EOF)
    if (pk->h->n < m - 1 || pk->l->n < c->na || pk->a->n < (size_t)c->ni + c->a_aux_total || pk->b_g1->n < (size_t)c->b_in_total + c->b_aux_total ||
        pk->b_g2->n < (size_t)c->b_in_total + c->b_aux_total)
        throw ZaError(ZA_ERR_IO, "create_proof: a proving-key query is shorter than the circuit needs (bellman: unexpected EOF)");
}

// ParameterSource for &Parameters: get_h -> (h,0); get_l -> (l,0); get_a -> ((a,0),(a,num_inputs));
// get_b_g1/g2 -> ((b,0),(b,b_input_density_total)).  bellman runs the input part and the aux part of the A and B
// queries as separate multiexps and adds the results; the bases are adjacent in the query, so here each query
// is ONE multiexp over the concatenated exponent vector (the sums a_inputs + a_aux etc. are all create_proof
// uses).  Five multiexps are in flight: H, L, A, B(G1), B(G2) — the last two share one digit sort.  Each query
// is cut by point range across ranks.
enum { MSM_WITNESS = 1, MSM_H = 2 };     // za_prove_msm_enqueue `which`
// The point ranges of the five queries this device adds up: the key's explicit plan (za_pk_partition_ranges) or, without
// one, the (weighted) contiguous shares of `rank` of `world`.
struct QueryRanges { size_t lo[5], hi[5]; };
static QueryRanges query_ranges(const Pk* pk, const Circuit* c, size_t m, int rank, int world) {
    QueryRanges q;
    if (pk->plan_set) {
        for (int i = 0; i < 5; i++) { q.lo[i] = pk->plan_lo[i]; q.hi[i] = pk->plan_hi[i]; }
        return q;
    }
    const uint32_t w0 = world > 1 ? pk->rank0_weight : 1000u;
    share(m - 1, rank, world, q.lo[Q_H], q.hi[Q_H]);
    share_weighted(c->na, rank, world, w0, q.lo[Q_L], q.hi[Q_L]);
    share_weighted(c->a_cat_total, rank, world, w0, q.lo[Q_A], q.hi[Q_A]);
    share_weighted(c->b_cat_total, rank, world, w0, q.lo[Q_B1], q.hi[Q_B1]);
    q.lo[Q_B2] = q.lo[Q_B1]; q.hi[Q_B2] = q.hi[Q_B1];
    return q;
}
static void prove_msms_enqueue(Ctx* ctx, const Pk* pk, const Circuit* c, const uint8_t* d_wit, const Fr* d_h, int rank, int world,
                               int which = MSM_WITNESS | MSM_H) {
    const uint32_t ni = c->ni, na = c->na;
    const size_t m = domain_size(c, nullptr);
    check_query_lengths(pk, c, m);
    const uint8_t* d_aux = d_wit + (size_t)ni * 32;
    const QueryRanges qr = query_ranges(pk, c, m, rank, world);
    size_t lo, hi;
    if (which & MSM_WITNESS) {
        // G2 first: its (longest) bucket reduction then overlaps the G1 accumulations on the side stream
        lo = qr.lo[Q_B1]; hi = qr.hi[Q_B1];
        const size_t lo2 = qr.lo[Q_B2], hi2 = qr.hi[Q_B2];
        const bool same_b = lo2 == lo && hi2 == hi;
        const uint32_t* sb = gather_range(ctx, d_wit, c->b_cat_idx, c->b_cat_total, lo, hi, ctx->scratch[13]);
        // B in G2 over a different range than B in G1 (explicit plans): its exponents are gathered on their own
        const uint32_t* sb2 = same_b ? sb : gather_range(ctx, d_wit, c->b_cat_idx, c->b_cat_total, lo2, hi2, ctx->scratch[15]);
        // The G2 multiexp goes to its own (high-priority) stream.  Its kernels hold 2 CTAs of 190+ registers per SM and
        // are latency bound (8 warps/SM, multiply pipe ~60 % busy); the registers they leave fit exactly one CTA of the
        // G1 accumulation, so the G1 multiexps on the main stream run next to it and fill the pipe
        // (profiles/r01_session3.md).  It starts as soon as the shared digit sort of the B (G1) multiexp is done.
        // ZA_G2_INLINE=1: everything on the main stream, one after the other.
        static const bool g2_inline = getenv("ZA_G2_INLINE") != nullptr;
        cudaStream_t main_st = ctx->stream;
        if (!g2_inline) {
            if (!ctx->g2_stream) {
                int lo_prio = 0, hi_prio = 0;
                ZA_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
                ZA_CUDA(cudaStreamCreateWithPriority(&ctx->g2_stream, cudaStreamNonBlocking, hi_prio));
                ZA_CUDA(cudaEventCreateWithFlags(&ctx->g2_fork, cudaEventDisableTiming));
            }
            ZA_CUDA(cudaEventRecord(ctx->g2_fork, main_st));            // the B scalars are gathered
            ZA_CUDA(cudaStreamWaitEvent(ctx->g2_stream, ctx->g2_fork, 0));
        }
        // K4': B (G1), L and A as ONE multiexp with three bucket spaces — one digit sort, one accumulation launch, one
        // reduction chain instead of three of each — when all three ranges have fixed-base tables of the same window layout
        // (pk_build_tables chooses the window jointly for exactly that reason).  ZA_MSM_MERGE=0: three multiexps as before.
        static const bool merge_on = !(getenv("ZA_MSM_MERGE") && atoi(getenv("ZA_MSM_MERGE")) == 0) && !(getenv("ZA_MSM_ACC_SM") && atoi(getenv("ZA_MSM_ACC_SM")) == 0);
        const size_t l_lo = qr.lo[Q_L], l_hi = qr.hi[Q_L], a_lo = qr.lo[Q_A], a_hi = qr.hi[Q_A];
        const Affine<Fq>* tb = table_for<Fq>(pk->b_g1.get(), lo, hi - lo);
        const Affine<Fq>* tl = table_for<Fq>(pk->l.get(), l_lo, l_hi - l_lo);
        const Affine<Fq>* ta = table_for<Fq>(pk->a.get(), a_lo, a_hi - a_lo);
        ctx->witness_merged = merge_on && tb && tl && ta && pk->b_g1->tab_c == pk->l->tab_c && pk->l->tab_c == pk->a->tab_c &&
                              pk->b_g1->tab_W == pk->l->tab_W && pk->l->tab_W == pk->a->tab_W;
        if (ctx->witness_merged) {
            const uint32_t* sa = gather_range(ctx, d_wit, c->a_cat_idx, c->a_cat_total, a_lo, a_hi, ctx->scratch[12]);
            MsmPart more[2] = {{tl, (const uint32_t*)(d_aux + l_lo * 32), l_hi - l_lo}, {ta, sa + a_lo * 8, a_hi - a_lo}};
            if (!g2_inline) {                                              // the fork must lie behind the A gather as well
                ZA_CUDA(cudaEventRecord(ctx->g2_fork, main_st));
                ZA_CUDA(cudaStreamWaitEvent(ctx->g2_stream, ctx->g2_fork, 0));
            }
            msm_enqueue<Fq>(ctx, 3, pk->b_g1->pts.as<G1Affine>() + lo, sb + lo * 8, hi - lo, false, -1, tb, pk->b_g1->tab_c, pk->b_g1->tab_W, more, 2);
        } else
        multiexp_enqueue<Fq>(ctx, 3, pk->b_g1.get(), lo, sb + lo * 8, hi - lo);
        // The two B multiexps run over the same exponents and share one digit sort — but only if both resolve to the
        // same bucket layout: with a fixed-base table on one side only (the G2 table is twice the size and may be over
        // the memory cap when the G1 table is not) or tables of different window sizes, G2 sorts for itself.
        int share = -1;
        if (same_b && hi - lo > 64) {
            const bool t1 = table_for<Fq>(pk->b_g1.get(), lo, hi - lo) != nullptr, t2 = table_for<Fq2>(pk->b_g2.get(), lo, hi - lo) != nullptr;
            if (t1 == t2 && (!t1 || (pk->b_g1->tab_c == pk->b_g2->tab_c && pk->b_g1->tab_W == pk->b_g2->tab_W))) share = 3;
        }
        if (!g2_inline) ctx->stream = ctx->g2_stream;
        try { multiexp_enqueue<Fq2>(ctx, 4, pk->b_g2.get(), lo2, sb2 + lo2 * 8, hi2 - lo2, share); }
        catch (...) { ctx->stream = main_st; throw; }
        ctx->stream = main_st;
    }
    if ((which & MSM_H) && (which & MSM_WITNESS)) {
        lo = qr.lo[Q_H]; hi = qr.hi[Q_H];
        multiexp_enqueue<Fq>(ctx, 0, pk->h.get(), lo, (const uint32_t*)(d_h + lo), hi - lo);
    }
    if ((which & MSM_WITNESS) && !ctx->witness_merged) {
        lo = qr.lo[Q_L]; hi = qr.hi[Q_L];
        multiexp_enqueue<Fq>(ctx, 1, pk->l.get(), lo, (const uint32_t*)(d_aux + lo * 32), hi - lo);
        lo = qr.lo[Q_A]; hi = qr.hi[Q_A];
        const uint32_t* sa = gather_range(ctx, d_wit, c->a_cat_idx, c->a_cat_total, lo, hi, ctx->scratch[12]);
        multiexp_enqueue<Fq>(ctx, 2, pk->a.get(), lo, sa + lo * 8, hi - lo);
    }
    if ((which & MSM_H) && !(which & MSM_WITNESS)) {
        lo = qr.lo[Q_H]; hi = qr.hi[Q_H];
        // ZA_H_AFTER_G2=1: the H multiexp starts behind the G2 accumulation of this device instead of next to it, so that
        // the (long, latency-bound) G2 bucket reduction runs under the H accumulation instead of after it
        // (measured per device of an 8-GPU plan: 4.47 -> 4.25 ms, 4 GPUs: 7.33 -> 7.02; on ONE GPU, where the G1 multiexps
        // fill the card next to G2, it costs 0.4 ms — so it follows the explicit plan of a multi-device key unless set)
        static const int h_after_env = getenv("ZA_H_AFTER_G2") ? atoi(getenv("ZA_H_AFTER_G2")) : -1;
        const bool h_after_g2 = h_after_env >= 0 ? h_after_env != 0 : pk->plan_set;
        if (h_after_g2 && ctx->slots[4].busy && ctx->slots[4].kind == 2 && ctx->slots[4].acc_done) ZA_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->slots[4].acc_done, 0));
        multiexp_enqueue<Fq>(ctx, 0, pk->h.get(), lo, (const uint32_t*)(d_h + lo), hi - lo);
    }
}
// second half of prove_msms: wait for the five multiexps in completion order and combine their windows
static void prove_msms_collect(Ctx* ctx, Partials& out) {
    out = partials_zero();
    if (ctx->witness_merged && ctx->slots[3].busy && ctx->slots[3].nparts == 3) {
        G1XYZZ parts[3];
        msm_finish<Fq>(ctx, 3, parts);
        out.g1[5] = parts[0]; out.g1[1] = parts[1]; out.g1[3] = parts[2];      // B, L, A
        out.g2[1] = msm_finish<Fq2>(ctx, 4);
        out.g1[0] = msm_finish<Fq>(ctx, 0);
        return;
    }
    out.g1[5] = msm_finish<Fq>(ctx, 3);      // b1_inputs + b1_aux
    out.g2[1] = msm_finish<Fq2>(ctx, 4);     // b2_inputs + b2_aux
    out.g1[0] = msm_finish<Fq>(ctx, 0);
    out.g1[1] = msm_finish<Fq>(ctx, 1);
    out.g1[3] = msm_finish<Fq>(ctx, 2);      // a_inputs + a_aux
}

// The same eight multiexps bellman runs, one by one (parity of every intermediate result; trace mode only).
static void prove_msms_separate(Ctx* ctx, const Pk* pk, const Circuit* c, const uint8_t* d_wit, const Fr* d_h, Partials& out) {
    const uint32_t ni = c->ni, na = c->na;
    const size_t m = domain_size(c, nullptr);
    out = partials_zero();
    check_query_lengths(pk, c, m);
    const uint8_t* d_in = d_wit;
    const uint8_t* d_aux = d_wit + (size_t)ni * 32;
    out.g1[0] = multiexp_dev<Fq>(ctx, pk->h.get(), 0, (const uint32_t*)d_h, m - 1);
    out.g1[1] = multiexp_dev<Fq>(ctx, pk->l.get(), 0, (const uint32_t*)d_aux, na);
    DevBuf& gbuf = ctx->scratch[1];
    out.g1[2] = multiexp_dev<Fq>(ctx, pk->a.get(), 0, (const uint32_t*)d_in, ni);
    const uint32_t* sc_bin = gather(ctx, d_in, c->b_in_idx, c->b_in_total, gbuf);
    out.g1[4] = multiexp_dev<Fq>(ctx, pk->b_g1.get(), 0, sc_bin, c->b_in_total);
    out.g2[0] = multiexp_dev<Fq2>(ctx, pk->b_g2.get(), 0, sc_bin, c->b_in_total);
    const uint32_t* sc = gather(ctx, d_aux, c->a_aux_idx, c->a_aux_total, gbuf);
    out.g1[3] = multiexp_dev<Fq>(ctx, pk->a.get(), ni, sc, c->a_aux_total);
    sc = gather(ctx, d_aux, c->b_aux_idx, c->b_aux_total, gbuf);
    out.g1[5] = multiexp_dev<Fq>(ctx, pk->b_g1.get(), c->b_in_total, sc, c->b_aux_total);
    out.g2[1] = multiexp_dev<Fq2>(ctx, pk->b_g2.get(), c->b_in_total, sc, c->b_aux_total);
}

// what create_proof refuses before it touches the GPU: r, s >= the modulus, delta at infinity
static void prove_validate(const Pk* pk, const uint8_t* r_le, const uint8_t* s_le) {
    if (pk->delta_g1.is_inf() || pk->delta_g2.is_inf()) throw ZaError(ZA_ERR_UNEXPECTED_IDENTITY, "create_proof: delta is the point at infinity");
    check_scalars_canonical(r_le, 1, "r"); check_scalars_canonical(s_le, 1, "s");
}

// The part of the assembly that does not depend on the multiexps (steps 6-7: delta*r + alpha, delta2*s + beta2,
// delta*rs + alpha*s + beta*r); computed on the host while the GPU runs the multiexps.
struct AssemblePre { G1XYZZ g_a, g_c; G2XYZZ g_b; uint32_t r[8], s[8]; };
static AssemblePre prove_assemble_pre(const Pk* pk, const uint8_t* r_le, const uint8_t* s_le) {
    prove_validate(pk, r_le, s_le);
    AssemblePre A;
    memcpy(A.r, r_le, 32); memcpy(A.s, s_le, 32);
    Fr rf, sf; memcpy(rf.v, A.r, 32); memcpy(sf.v, A.s, 32);
    Fr rs_c = fp_from_mont<FrParams>(fp_to_mont<FrParams>(rf) * fp_to_mont<FrParams>(sf));
    G1XYZZ d1 = G1XYZZ::from_affine(pk->delta_g1), al = G1XYZZ::from_affine(pk->alpha_g1), be1 = G1XYZZ::from_affine(pk->beta_g1);
    G2XYZZ d2 = G2XYZZ::from_affine(pk->delta_g2);
    A.g_a = xyzz_mul<Fq>(d1, A.r); xyzz_madd<Fq>(A.g_a, pk->alpha_g1);
    A.g_b = xyzz_mul<Fq2>(d2, A.s); xyzz_madd<Fq2>(A.g_b, pk->beta_g2);
    A.g_c = xyzz_mul<Fq>(d1, rs_c.v);
    xyzz_add<Fq>(A.g_c, xyzz_mul<Fq>(al, A.s));
    xyzz_add<Fq>(A.g_c, xyzz_mul<Fq>(be1, A.r));
    return A;
}
static void prove_assemble_post(const AssemblePre& A, const Partials& P, uint8_t* proof_out) {
    G1XYZZ g_a = A.g_a, g_c = A.g_c; G2XYZZ g_b = A.g_b;
    G1XYZZ a_ans = P.g1[2]; xyzz_add<Fq>(a_ans, P.g1[3]);
    xyzz_add<Fq>(g_a, a_ans);
    xyzz_add<Fq>(g_c, xyzz_mul<Fq>(a_ans, A.s));
    G1XYZZ b1_ans = P.g1[4]; xyzz_add<Fq>(b1_ans, P.g1[5]);
    G2XYZZ b2_ans = P.g2[0]; xyzz_add<Fq2>(b2_ans, P.g2[1]);
    xyzz_add<Fq2>(g_b, b2_ans);
    xyzz_add<Fq>(g_c, xyzz_mul<Fq>(b1_ans, A.r));
    xyzz_add<Fq>(g_c, P.g1[0]);
    xyzz_add<Fq>(g_c, P.g1[1]);
    g1_to_le(xyzz_to_affine<Fq>(g_a), proof_out);
    g2_to_le(xyzz_to_affine<Fq2>(g_b), proof_out + 64);
    g1_to_le(xyzz_to_affine<Fq>(g_c), proof_out + 192);
}
static void prove_assemble(const Pk* pk, const Partials& P, const uint8_t* r_le, const uint8_t* s_le, uint8_t* proof_out) {
    prove_assemble_post(prove_assemble_pre(pk, r_le, s_le), P, proof_out);
}
static void prove_msms(Ctx* ctx, const Pk* pk, const Circuit* c, const uint8_t* d_wit, const Fr* d_h, int rank, int world, Partials& out) {
    prove_msms_enqueue(ctx, pk, c, d_wit, d_h, rank, world);
    prove_msms_collect(ctx, out);
}

// ZA_DEBUG_TIMELINE: when each multiexp of the last proof started, finished accumulating and finished reducing
static void print_timeline(Ctx* ctx) {
    static const bool timeline = getenv("ZA_DEBUG_TIMELINE") != nullptr;
    if (!timeline) return;
    const int order[5] = {3, 4, 0, 1, 2};
    const char* name[5] = {"B(G1)", "B(G2)", "H", "L", "A"};
    cudaEvent_t base = ctx->dbg_t0_valid ? ctx->dbg_t0 : ctx->slots[3].dbg_start;
    ctx->dbg_t0_valid = false;
    for (int k = 0; k < 5; k++) {
        MsmSlot& sl = ctx->slots[order[k]];
        float a = 0, b = 0, d = 0, so = 0;
        if (sl.dbg_start && sl.kind == 2) {
            cudaEventSynchronize(sl.dbg_done);
            cudaEventElapsedTime(&a, base, sl.dbg_start); cudaEventElapsedTime(&b, base, sl.dbg_acc); cudaEventElapsedTime(&d, base, sl.dbg_done);
            if (sl.dbg_sort) cudaEventElapsedTime(&so, base, sl.dbg_sort);
        }
        fprintf(stderr, "[za timeline] %-6s start %.3f  sorted %.3f  accumulate done %.3f  reduce done %.3f\n", name[k], a, so, b, d);
    }
}

static void create_proof_device_body(Ctx* ctx, const Pk* pk, const Circuit* c, const uint8_t* d_wit, const uint8_t* r_le, const uint8_t* s_le,
                                     uint8_t* proof_out, za_trace* tr);

// whole proof on one GPU, witness [inputs | aux] already on the device.  "Every witness element < r" is checked on the
// device copy (a pass over 32 MB on the host would sit in front of every proof); the verdict travels back behind the
// check and is read once the proof is done.  Any failure leaves the context reusable (msm_abort).
static void create_proof_device(Ctx* ctx, const Pk* pk, const Circuit* c, const uint8_t* d_wit, const uint8_t* r_le, const uint8_t* s_le,
                                uint8_t* proof_out, za_trace* tr) {
    prove_validate(pk, r_le, s_le);
    check_query_lengths(pk, c, domain_size(c, nullptr));
    cudaStream_t st = ctx->stream;
    const uint32_t ni = c->ni, na = c->na;
    canonical_check_enqueue(ctx, d_wit, (size_t)ni + na, CHK_WITNESS);
    bool failed = false;
    try {
        create_proof_device_body(ctx, pk, c, d_wit, r_le, s_le, proof_out, tr);
    } catch (...) {
        ctx->stream = st;
        msm_abort(ctx);
        if (!ctx->host_flag || ctx->host_flag[CHK_WITNESS] == ~0ull) throw;      // otherwise the non-canonical element below is the root cause
        failed = true;
    }
    if (!failed) ZA_CUDA(cudaStreamSynchronize(st));
    const unsigned long long first_bad = canonical_check_verdict(ctx, CHK_WITNESS);
    if (first_bad != ~0ull) {
        memset(proof_out, 0, 256);
        char b[128];
        if (first_bad < ni) snprintf(b, sizeof b, "inputs[%llu] is not a canonical Fr element (>= r)", first_bad);
        else snprintf(b, sizeof b, "aux[%llu] is not a canonical Fr element (>= r)", first_bad - ni);
        throw ZaError(ZA_ERR_NOT_CANONICAL, b);
    }
}

static void create_proof_device_body(Ctx* ctx, const Pk* pk, const Circuit* c, const uint8_t* d_wit, const uint8_t* r_le, const uint8_t* s_le,
                                     uint8_t* proof_out, za_trace* tr) {
    const size_t m = domain_size(c, nullptr);
    DevBuf& h = ctx->scratch[9];
    h.ensure(m * sizeof(Fr));
    // The H pipeline (r1cs evaluation, seven NTTs) runs on its own high-priority stream NEXT TO the witness multiexps:
    // B, L and A do not depend on it, its short kernels leave the SMs half empty at every launch boundary and wave
    // tail, and the accumulation CTAs fill those holes.  The NTTs stretch (2.2 -> ~13 ms, in the shadow), the proof
    // gets 0.8 ms shorter (profiles/r01_session3.md); the H multiexp is enqueued behind the pipeline.
    // ZA_H_INLINE=1: H pipeline first, on the main stream.
    static const bool h_overlap = getenv("ZA_H_INLINE") == nullptr;
    if (h_overlap && !tr) {
        cudaStream_t main_st = ctx->stream;
        if (!ctx->h_stream) {
            int lo_prio = 0, hi_prio = 0;
            ZA_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
            ZA_CUDA(cudaStreamCreateWithPriority(&ctx->h_stream, cudaStreamNonBlocking, hi_prio));
            ZA_CUDA(cudaEventCreateWithFlags(&ctx->h_fork, cudaEventDisableTiming));
            ZA_CUDA(cudaEventCreateWithFlags(&ctx->h_join, cudaEventDisableTiming));
        }
        ZA_CUDA(cudaEventRecord(ctx->h_fork, main_st));
        ZA_CUDA(cudaStreamWaitEvent(ctx->h_stream, ctx->h_fork, 0));
        ctx->stream = ctx->h_stream;
        try { prove_h(ctx, c, d_wit, h.as<Fr>(), nullptr); } catch (...) { ctx->stream = main_st; throw; }
        ZA_CUDA(cudaEventRecord(ctx->h_join, ctx->h_stream));
        ctx->stream = main_st;
        prove_msms_enqueue(ctx, pk, c, d_wit, h.as<Fr>(), 0, 1, MSM_WITNESS);
        ZA_CUDA(cudaStreamWaitEvent(main_st, ctx->h_join, 0));
        prove_msms_enqueue(ctx, pk, c, d_wit, h.as<Fr>(), 0, 1, MSM_H);
        AssemblePre pre = prove_assemble_pre(pk, r_le, s_le);
        Partials Q;
        prove_msms_collect(ctx, Q);
        prove_assemble_post(pre, Q, proof_out);
        print_timeline(ctx);
        return;
    }
    prove_h(ctx, c, d_wit, h.as<Fr>(), tr);
    Partials P;
    if (tr && (tr->msm_g1 || tr->msm_g2)) {
        // trace mode: bellman's eight multiexps individually, and the proof must not depend on the path taken
        Partials S;
        prove_msms_separate(ctx, pk, c, d_wit, h.as<Fr>(), S);
        if (tr->msm_g1) for (int i = 0; i < 6; i++) g1_to_le(xyzz_to_affine<Fq>(S.g1[i]), tr->msm_g1 + 64 * i);
        if (tr->msm_g2) for (int i = 0; i < 2; i++) g2_to_le(xyzz_to_affine<Fq2>(S.g2[i]), tr->msm_g2 + 128 * i);
        uint8_t check[256];
        prove_assemble(pk, S, r_le, s_le, check);
        prove_msms(ctx, pk, c, d_wit, h.as<Fr>(), 0, 1, P);
        prove_assemble(pk, P, r_le, s_le, proof_out);
        if (memcmp(check, proof_out, 256) != 0) throw ZaError(ZA_ERR_INVALID, "internal: fused and separate multiexp paths disagree");
        return;
    }
    static const bool timeline = getenv("ZA_DEBUG_TIMELINE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    if (timeline) { cudaStreamSynchronize(ctx->stream); t0 = now(); }
    prove_msms_enqueue(ctx, pk, c, d_wit, h.as<Fr>(), 0, 1);
    if (timeline) t1 = now();
    AssemblePre pre = prove_assemble_pre(pk, r_le, s_le);      // host work under the GPU's multiexps
    if (timeline) t2 = now();
    prove_msms_collect(ctx, P);
    if (timeline) t3 = now();
    prove_assemble_post(pre, P, proof_out);
    print_timeline(ctx);
    if (timeline) { t4 = now(); fprintf(stderr, "[za timeline] after H: enqueue %.3f ms, host pre %.3f, wait+combine %.3f, host post %.3f\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3); }
}

// index of the first element >= r in a vector of canonical-form candidates (atomicMin; ~0 if all are canonical)
__global__ void fr_canonical_check_kernel(const uint32_t* __restrict__ v, size_t n, unsigned long long* first_bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* p = reinterpret_cast<const uint4*>(v + 8 * i);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (!fp_is_canonical<FrParams>(w)) atomicMin(first_bad, (unsigned long long)i);
}

static void create_proof(Ctx* ctx, const Pk* pk, const Circuit* c, const uint8_t* inputs, const uint8_t* aux, const uint8_t* r_le,
                         const uint8_t* s_le, uint8_t* proof_out, za_trace* tr) {
    cudaStream_t st = ctx->stream;
    const uint32_t ni = c->ni, na = c->na;
    DevBuf& wit = ctx->scratch[10];
    wit.ensure(((size_t)ni + na) * 32);
    ZA_CUDA(cudaMemcpyAsync(wit.p, inputs, (size_t)ni * 32, cudaMemcpyHostToDevice, st));
    if (na) ZA_CUDA(cudaMemcpyAsync((uint8_t*)wit.p + (size_t)ni * 32, aux, (size_t)na * 32, cudaMemcpyHostToDevice, st));
    create_proof_device(ctx, pk, c, (const uint8_t*)wit.p, r_le, s_le, proof_out, tr);
}

// synthetic proving key: every base is a known multiple of the generator (bench + full-size property tests)
static Affine<Fq> host_g1_gen() { Affine<Fq> g; g.x = fp_from_u64<FqParams>(1); g.y = fp_from_u64<FqParams>(2); return g; }
static Affine<Fq2> host_g2_gen() {
    // /root/reference/prover/src/groth16/ethereum.rs:28-31
    static const uint32_t X0[8] = {0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu};
    static const uint32_t X1[8] = {0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u};
    static const uint32_t Y0[8] = {0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u};
    static const uint32_t Y1[8] = {0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u};
    Affine<Fq2> g; Fq t;
    memcpy(t.v, X0, 32); g.x.c0 = fp_to_mont<FqParams>(t); memcpy(t.v, X1, 32); g.x.c1 = fp_to_mont<FqParams>(t);
    memcpy(t.v, Y0, 32); g.y.c0 = fp_to_mont<FqParams>(t); memcpy(t.v, Y1, 32); g.y.c1 = fp_to_mont<FqParams>(t);
    return g;
}
static std::unique_ptr<Bases> bases_generated(Ctx* ctx, int group, size_t n, uint64_t first) {
    std::unique_ptr<Bases> b(new Bases());
    b->ctx = ctx; b->group = group; b->n = n; b->has_infinity = false;
    if (first == 0 || first + n >= ((uint64_t)1 << 62)) throw ZaError(ZA_ERR_INVALID, "bases_generate: multiplier range must be in [1, 2^62)");
    if (group == 1) { b->pts.alloc(n * sizeof(G1Affine)); bases_generate<Fq>(ctx, b->pts.as<G1Affine>(), n, first, host_g1_gen()); }
    else { b->pts.alloc(n * sizeof(G2Affine)); bases_generate<Fq2>(ctx, b->pts.as<G2Affine>(), n, first, host_g2_gen()); }
    return b;
}
template <class F>
static Affine<F> host_multiple(const Affine<F>& g, uint64_t k) {
    uint32_t w[8] = {(uint32_t)k, (uint32_t)(k >> 32), 0, 0, 0, 0, 0, 0};
    return xyzz_to_affine<F>(xyzz_mul<F>(XYZZ<F>::from_affine(g), w));
}


// ------------------------------------------------------------------ several GPUs, one process
// The reference's caller is ONE process (create_random_proof at prover.rs:173), so the multi-GPU path sits behind one
// call as well: a Prover owns a context, the proving key and the circuit on every device, and one host thread per
// device that issues that device's work.  Per proof:
//   device 0   witness H2D -> range check -> constraint evaluation -> H pipeline (alone: it is the critical path of
//              every device) -> the h slice of device k straight into k's memory (cudaMemcpyPeerAsync over NVLink, an
//              event per peer) -> its (smaller, za_pk_partition_weighted) share of the witness multiexps -> its H share
//   device k   the witness span its point ranges need, H2D over its own PCIe link -> witness multiexps (L, A, B in G1,
//              B in G2) -> waits for its h slice on the stream -> H multiexp
//   host       delta r + alpha etc. while the GPUs work; then adds the partial sums (8 points per device) and
//              assembles.  No collective and no torch / NCCL anywhere on the path.
struct ProverWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool pending = false, quit = false;
    int err_code = 0;
    std::string err;
};
struct Prover {
    int n = 0;
    std::vector<int> devices;
    std::vector<za_ctx*> ctx;
    std::vector<za_pk*> pk;
    std::vector<za_circuit*> circ;
    std::vector<std::unique_ptr<ProverWorker>> workers;
    std::vector<DevBuf> wit, h;                      // per device: witness [inputs | aux]; h scalars (m entries)
    std::vector<cudaEvent_t> h_ready;                // on device 0's stream, behind the copy of device k's h slice
    std::vector<std::vector<std::pair<size_t, size_t>>> spans;   // per device: witness positions [first, last) it reads
    std::mutex h_mu;
    std::condition_variable h_cv;
    uint64_t h_issued = 0, generation = 0;           // device 0 has issued the h copies of proof `generation`
    bool h_failed = false;
    bool peers_mapped = false;                       // device 0 can store into every other device's memory
    uint32_t rank0_weight = 1000;
    std::vector<Partials> partials;
    ~Prover();
};

static void prover_worker_loop(ProverWorker* w) {
    for (;;) {
        std::function<void()> job;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->pending || w->quit; });
            if (w->quit) return;
            job = w->job;
        }
        int code = 0; std::string text;
        try { job(); }
        catch (const ZaError& e) { code = e.code; text = e.what(); }
        catch (const CudaError& e) { code = ZA_ERR_CUDA; text = e.what(); }
        catch (const std::exception& e) { code = ZA_ERR_INVALID; text = e.what(); }
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->err_code = code; w->err = text; w->pending = false;
        }
        w->cv.notify_all();
    }
}
// run fn(k) on device k's thread for every device; the caller's thread runs `meanwhile`.  Throws the first error.
static void prover_run_all(Prover* P, const std::function<void(int)>& fn, const std::function<void()>& meanwhile = nullptr) {
    for (int k = 0; k < P->n; k++) {
        ProverWorker* w = P->workers[k].get();
        { std::lock_guard<std::mutex> lk(w->mu); w->job = [fn, k] { fn(k); }; w->pending = true; w->err_code = 0; w->err.clear(); }
        w->cv.notify_all();
    }
    int mcode = 0; std::string mtext;
    if (meanwhile) {
        try { meanwhile(); }
        catch (const ZaError& e) { mcode = e.code; mtext = e.what(); }
        catch (const std::exception& e) { mcode = ZA_ERR_INVALID; mtext = e.what(); }
    }
    int code = 0; std::string text;
    for (int k = 0; k < P->n; k++) {
        ProverWorker* w = P->workers[k].get();
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return !w->pending; });
        if (w->err_code && !code) { code = w->err_code; text = "device " + std::to_string(P->devices[k]) + ": " + w->err; }
    }
    if (mcode) throw ZaError(mcode, mtext);
    if (code) throw ZaError(code, text);
}
Prover::~Prover() {
    for (auto& w : workers) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
    }
    if (!devices.empty()) cudaSetDevice(devices[0]);
    for (cudaEvent_t e : h_ready) if (e) cudaEventDestroy(e);
    for (int k = 0; k < (int)ctx.size(); k++) {
        cudaSetDevice(devices[k]);
        if (k < (int)wit.size()) { wit[k].release(); h[k].release(); }
        if (k < (int)circ.size() && circ[k]) delete circ[k];
        if (k < (int)pk.size() && pk[k]) delete pk[k];
    }
    for (int k = 0; k < (int)ctx.size(); k++) if (ctx[k]) za_ctx_destroy(ctx[k]);
}

// Which part of which query every device adds up (one proof = the sum over the devices).
// Round 1 gave every device one (weighted) slice of EVERY query: five short multiexps per device, each with its own sort,
// launch chain and bucket-reduction tail — on 8 GPUs one rank's share took 5.3 ms of which about 3 ms were bulk work
// (profiles/r02_shards.md).  The plan below works on estimated time instead (a cost model, ms per 2^20 points on B200:
// H pipeline hp, G1 multiexp g1, G2 multiexp g2, plus a latency-bound tail per device after its last accumulation):
//   * devices 1 .. n2 are "pure G2" devices: they split the front of the B (G2) query evenly and run nothing else — the G2
//     kernels do not share an SM well with anything (§4.4 of DESIGN.md), and the h exponents they would have to wait for
//     do not exist before hp;
//   * the other devices share what is left of B (G2), then B (G1), A and L laid end to end (one contiguous piece each, so a
//     device touches one or two queries with long ranges) and the H query, each in proportion to the time it has: device 0
//     runs the H pipeline first and gets less;
//   * n2 and the common finish time T are found by trying every n2 and bisecting T.
// ZA_PROVER_COSTS="hp,g1,g2,tail_g1,tail_g2" overrides the model (defaults 2.1, 3.9, 8.6, 0.9, 1.2: a device's share of a
// query costs more per point than the whole query — smaller window, shorter waves);
// ZA_PROVER_PLAN=0 restores the slices, ZA_PROVER_PLAN=1 the first version of the line (no pure G2 devices, H split evenly).
struct DevicePlan { size_t lo[5], hi[5]; };
struct PlanModel { double hp = 2.1, g1 = 3.9, g2 = 8.6, t1 = 0.9, t2 = 1.2; };      // fitted to profiles/r02_shards.md (per-device times; 8 GPUs: 4.67 -> 4.37 ms against 3.7 / 9.5 / 0.8 / 1.3)
static PlanModel plan_model() {
    PlanModel pm;
    if (const char* e = getenv("ZA_PROVER_COSTS")) {
        double v[5];
        const int n = sscanf(e, "%lf,%lf,%lf,%lf,%lf", &v[0], &v[1], &v[2], &v[3], &v[4]);
        if (n >= 3 && v[0] >= 0 && v[1] > 0 && v[2] > 0) { pm.hp = v[0]; pm.g1 = v[1]; pm.g2 = v[2]; }
        if (n == 5 && v[3] >= 0 && v[4] >= 0) { pm.t1 = v[3]; pm.t2 = v[4]; }
    }
    return pm;
}
// cut `cnt` items of a query that occupies [start, start + len) of a line at position x
static size_t line_index(double start, double len, size_t cnt, double x) {
    if (len <= 0 || x <= start) return 0;
    if (x >= start + len) return cnt;
    const size_t v = (size_t)((x - start) / len * (double)cnt);
    return v > cnt ? cnt : v;
}
static std::vector<DevicePlan> prover_make_plans_v1(const size_t* cnt, size_t m, int N);
static std::vector<DevicePlan> prover_make_plans(const size_t* cnt, size_t m, int N) {
    static const int version = getenv("ZA_PROVER_PLAN") ? atoi(getenv("ZA_PROVER_PLAN")) : 2;
    if (version == 1 || N < 2) return prover_make_plans_v1(cnt, m, N);
    const PlanModel pm = plan_model();
    const double unit = 1.0 / 1048576.0, scale = (double)m * unit;          // tails and the pipeline scale with the domain
    const double c_b2 = pm.g2 * cnt[Q_B2] * unit, c_b1 = pm.g1 * cnt[Q_B1] * unit, c_a = pm.g1 * cnt[Q_A] * unit, c_l = pm.g1 * cnt[Q_L] * unit;
    const double c_h = pm.g1 * cnt[Q_H] * unit, hp = pm.hp * scale, t1 = pm.t1 * scale, t2 = pm.t2 * scale;
    const double w_g1 = c_b1 + c_a + c_l;
    struct Trial { double T; int n2; double pure; };        // pure: bulk of B (G2) on each pure device
    // shared devices in line order: n2 + 1, ..., N - 1, 0
    auto line_order = [&](int n2) { std::vector<int> o; for (int k = n2 + 1; k < N; k++) o.push_back(k); o.push_back(0); return o; };
    // capacities (time left for witness + H work) of the shared devices if everything is to finish by T: T minus the tail of
    // the device's last multiexp (the G2 tail for every device whose piece of the line reaches into B (G2): found by
    // iterating, the pieces depend on the capacities) minus the H pipeline on device 0
    auto shared_caps = [&](double T, int n2, double g2_left, std::vector<double>& cap) -> double {
        const std::vector<int> order = line_order(n2);
        std::vector<double> tail(N, t1);
        double total = 0;
        for (int it = 0; it < 4; it++) {
            cap.assign(N, 0.0);
            total = 0;
            for (int k : order) { const double v = T - tail[k] - (k == 0 ? hp : 0.0); cap[k] = v > 0 ? v : 0; total += cap[k]; }
            if (total <= 0 || g2_left <= 0) break;
            const double w_line = g2_left + w_g1;
            double x = 0;
            for (int k : order) { tail[k] = x < g2_left - 1e-12 && cap[k] > 0 ? t2 : t1; x += w_line * cap[k] / total; }
        }
        return total;
    };
    Trial best{1e300, 0, 0};
    std::vector<double> cap;
    for (int n2 = 0; n2 <= N - 2; n2++) {
        double lo = 0, hi = hp + c_b2 + w_g1 + c_h + t1 + t2 + 1.0;
        for (int it = 0; it < 60; it++) {
            const double T = 0.5 * (lo + hi);
            double pure = n2 ? std::min(std::max(T - t2, 0.0), c_b2 / n2) : 0.0;
            if (n2 && c_b2 - n2 * pure < 0.05 * c_b2) pure = c_b2 / n2;      // no sliver of B (G2) for the shared devices
            const double g2_left = c_b2 - n2 * pure;
            const double need = g2_left + w_g1 + c_h;
            const double total = shared_caps(T, n2, g2_left, cap);
            bool ok = total >= need;
            // a device other than 0 must have witness work until the h exponents exist (its H share would idle otherwise)
            if (ok && c_h > 0) {
                const double frac_w = (g2_left + w_g1) / need;
                for (int k = 1; k < N && ok; k++) if (cap[k] > 0 && cap[k] * frac_w < hp * 0.9) ok = false;
            }
            if (ok) hi = T; else lo = T;
        }
        if (hi < best.T - 1e-9) {
            best.T = hi; best.n2 = n2; best.pure = n2 ? std::min(std::max(hi - t2, 0.0), c_b2 / n2) : 0.0;
            if (n2 && c_b2 - n2 * best.pure < 0.05 * c_b2) best.pure = c_b2 / n2;
        }
    }
    const int n2 = best.n2;
    const double g2_pure_total = n2 * best.pure, g2_left = c_b2 - g2_pure_total;
    shared_caps(best.T, n2, g2_left, cap);
    double cap_total = 0; for (int k = 0; k < N; k++) cap_total += cap[k];
    // H ranges go in device order (0 first): the scatter of the last NTT pass needs ascending bounds
    const std::vector<int> order = line_order(n2);
    const double w_line = g2_left + w_g1;                       // the shared witness line: [rest of B2 | B1 | A | L]
    const double st_b2 = -g2_pure_total, st_b1 = g2_left, st_a = st_b1 + c_b1, st_l = st_a + c_a;      // B2 starts before 0: its front is the pure part
    std::vector<double> bound(order.size() + 1, 0.0);
    for (size_t i = 0; i < order.size(); i++) bound[i + 1] = bound[i] + (cap_total > 0 ? w_line * cap[order[i]] / cap_total : 0.0);
    bound[order.size()] = w_line;
    const double qs[3] = {st_b1, st_a, st_l}, ql[3] = {c_b1, c_a, c_l};
    const size_t qc[3] = {cnt[Q_B1], cnt[Q_A], cnt[Q_L]};
    for (size_t i = 1; i < order.size(); i++)                   // no slivers (a piece has a fixed cost): snap to query ends within 8 %
        for (int q = 0; q < 3; q++) {
            const double thr = 0.08 * pm.g1 * qc[q] * unit;
            if (fabs(bound[i] - qs[q]) < thr) bound[i] = qs[q];
            if (fabs(bound[i] - (qs[q] + ql[q])) < thr) bound[i] = qs[q] + ql[q];
        }
    std::vector<DevicePlan> plans(N);
    for (int k = 0; k < N; k++) for (int i = 0; i < 5; i++) { plans[k].lo[i] = 0; plans[k].hi[i] = 0; }
    for (int j = 1; j <= n2; j++) {                             // pure devices: even pieces of the front of B (G2)
        plans[j].lo[Q_B2] = line_index(st_b2, c_b2, cnt[Q_B2], st_b2 + (j - 1) * best.pure);
        plans[j].hi[Q_B2] = line_index(st_b2, c_b2, cnt[Q_B2], st_b2 + j * best.pure);
    }
    const size_t b2_pure_end = n2 ? plans[n2].hi[Q_B2] : 0;
    for (size_t i = 0; i < order.size(); i++) {
        DevicePlan& d = plans[order[i]];
        const double x0 = bound[i], x1 = bound[i + 1];
        const bool last = i + 1 == order.size();
        d.lo[Q_B2] = std::max(b2_pure_end, line_index(st_b2, c_b2, cnt[Q_B2], x0));
        d.hi[Q_B2] = last ? cnt[Q_B2] : std::max(d.lo[Q_B2], line_index(st_b2, c_b2, cnt[Q_B2], x1));
        if (i == 0) d.lo[Q_B2] = b2_pure_end;
        const int qq[3] = {Q_B1, Q_A, Q_L};
        for (int q = 0; q < 3; q++) {
            d.lo[qq[q]] = line_index(qs[q], ql[q], qc[q], x0);
            d.hi[qq[q]] = last ? qc[q] : line_index(qs[q], ql[q], qc[q], x1);
            if (d.hi[qq[q]] < d.lo[qq[q]]) d.hi[qq[q]] = d.lo[qq[q]];
        }
    }
    // a leftover of B (G2) too short to be worth a multiexp of its own (rounding): the last pure device takes it
    if (n2) {
        const int first_shared = order[0];
        DevicePlan& f = plans[first_shared];
        bool only_first = true;
        for (size_t i = 1; i < order.size(); i++) if (plans[order[i]].hi[Q_B2] > plans[order[i]].lo[Q_B2]) only_first = false;
        if (only_first && f.hi[Q_B2] > f.lo[Q_B2] && f.hi[Q_B2] - f.lo[Q_B2] < 4096) { plans[n2].hi[Q_B2] = f.hi[Q_B2]; f.lo[Q_B2] = f.hi[Q_B2]; }
    }
    // H in proportion to the same capacities, in device order
    {
        double acc = 0;
        size_t prev = 0;
        for (int k = 0; k < N; k++) {
            acc += cap[k];
            size_t hi = cap_total > 0 ? (size_t)((double)cnt[Q_H] * (acc / cap_total)) : (k == N - 1 ? cnt[Q_H] : 0);
            if (k == N - 1 || hi > cnt[Q_H]) hi = cnt[Q_H];
            if (hi < prev) hi = prev;
            plans[k].lo[Q_H] = prev; plans[k].hi[Q_H] = hi;
            prev = hi;
        }
        // the last device with a share takes the rounding rest
        for (int k = N - 1; k >= 0; k--) if (cap[k] > 0) { for (int j = k; j < N; j++) { plans[j].hi[Q_H] = cnt[Q_H]; if (j > k) plans[j].lo[Q_H] = cnt[Q_H]; } break; }
    }
    return plans;
}
static std::vector<DevicePlan> prover_make_plans_v1(const size_t* cnt, size_t m, int N) {
    double hp = 2.0, g1 = 3.0, g2 = 8.3;
    if (const char* e = getenv("ZA_PROVER_COSTS")) { double a, b, d; if (sscanf(e, "%lf,%lf,%lf", &a, &b, &d) == 3 && a >= 0 && b > 0 && d > 0) { hp = a; g1 = b; g2 = d; } }
    const int line[4] = {Q_B2, Q_B1, Q_A, Q_L};                     // order on the work line
    double len[4], start[4], total = 0;
    for (int i = 0; i < 4; i++) { len[i] = (line[i] == Q_B2 ? g2 : g1) * (double)cnt[line[i]]; start[i] = total; total += len[i]; }
    const double c_h = g1 * (double)cnt[Q_H], c_pipe = hp * (double)m;
    // every device: h share c_h / N; device 0 additionally the pipeline; equal finish times
    double w0 = (c_pipe + total + c_h) / N - c_pipe - c_h / N;
    if (w0 < 0) w0 = 0;
    if (N == 1) w0 = total;
    const double wk = N > 1 ? (total - w0) / (N - 1) : 0;
    std::vector<double> bound(N + 1, 0.0);                          // device k >= 1 owns [bound[k-1], bound[k]); device 0 the rest
    for (int k = 1; k < N; k++) bound[k] = bound[k - 1] + wk;
    bound[N] = total;
    // no slivers: a cut closer than 8 % of a G1 query's time to an end of a query moves to that end (a piece has a fixed cost of
    // a few hundred microseconds: sort, launch chain, reduction tail)
    for (int k = 1; k < N; k++)
        for (int i = 0; i < 4; i++) {
            if (len[i] <= 0) continue;
            const double thr = 0.08 * g1 * (double)cnt[line[i]];        // 8 % of a G1 query of that length, also for the G2 query
            if (fabs(bound[k] - start[i]) < thr) bound[k] = start[i];
            if (fabs(bound[k] - (start[i] + len[i])) < thr) bound[k] = start[i] + len[i];
        }
    auto index_at = [&](int i, double x, bool last) -> size_t {
        const size_t n = cnt[line[i]];
        if (last || x >= start[i] + len[i]) return n;
        if (x <= start[i] || len[i] <= 0) return 0;
        size_t v = (size_t)((x - start[i]) / len[i] * (double)n);
        return v > n ? n : v;
    };
    std::vector<DevicePlan> plans(N);
    for (int k = 0; k < N; k++) {
        DevicePlan& d = plans[k];
        share(cnt[Q_H], k, N, d.lo[Q_H], d.hi[Q_H]);
        const double x0 = k == 0 ? bound[N - 1] : bound[k - 1], x1 = k == 0 ? bound[N] : bound[k];
        for (int i = 0; i < 4; i++) {
            d.lo[line[i]] = index_at(i, x0, false);
            d.hi[line[i]] = index_at(i, x1, k == 0);
            if (d.hi[line[i]] < d.lo[line[i]]) d.hi[line[i]] = d.lo[line[i]];
        }
        if (N == 1) for (int i = 0; i < 5; i++) { d.lo[i] = 0; d.hi[i] = cnt[i]; }
    }
    return plans;
}

// witness positions device k reads: the aux range of its L share and the spans of its A and B index ranges (merged)
static void prover_plan_spans(Prover* P) {
    const Circuit* c = P->circ[0]->c.get();
    const size_t m = domain_size(c, nullptr);
    P->spans.assign(P->n, {});
    for (int k = 0; k < P->n; k++) {
        std::vector<std::pair<size_t, size_t>> iv;
        if (k == 0) { iv.push_back({0, (size_t)c->ni + c->na}); P->spans[k] = iv; continue; }      // device 0 evaluates the constraints
        const QueryRanges q = query_ranges(P->pk[k]->p.get(), c, m, k, P->n);
        if (q.hi[Q_L] > q.lo[Q_L]) iv.push_back({c->ni + q.lo[Q_L], c->ni + q.hi[Q_L]});
        if (q.hi[Q_A] > q.lo[Q_A]) iv.push_back({c->h_a_cat[q.lo[Q_A]], (size_t)c->h_a_cat[q.hi[Q_A] - 1] + 1});
        if (q.hi[Q_B1] > q.lo[Q_B1]) iv.push_back({c->h_b_cat[q.lo[Q_B1]], (size_t)c->h_b_cat[q.hi[Q_B1] - 1] + 1});
        if (q.hi[Q_B2] > q.lo[Q_B2]) iv.push_back({c->h_b_cat[q.lo[Q_B2]], (size_t)c->h_b_cat[q.hi[Q_B2] - 1] + 1});
        std::sort(iv.begin(), iv.end());
        std::vector<std::pair<size_t, size_t>> merged;
        for (auto& x : iv) {
            if (!merged.empty() && x.first <= merged.back().second) merged.back().second = std::max(merged.back().second, x.second);
            else merged.push_back(x);
        }
        P->spans[k] = merged;
    }
}

static void prover_upload_witness(Prover* P, int k, const uint8_t* inputs, const uint8_t* aux) {
    Ctx* cx = &P->ctx[k]->c;
    const Circuit* c = P->circ[k]->c.get();
    const size_t ni = c->ni;
    uint8_t* d = (uint8_t*)P->wit[k].p;
    for (auto& sp : P->spans[k]) {
        size_t a = sp.first, b = sp.second;
        if (a < ni) {
            const size_t e = b < ni ? b : ni;
            ZA_CUDA(cudaMemcpyAsync(d + a * 32, inputs + a * 32, (e - a) * 32, cudaMemcpyHostToDevice, cx->stream));
            a = e;
        }
        if (b > a) ZA_CUDA(cudaMemcpyAsync(d + a * 32, aux + (a - ni) * 32, (b - a) * 32, cudaMemcpyHostToDevice, cx->stream));
    }
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// one device's part of one proof (runs on that device's thread)
static void prover_device_step(Prover* P, int k, uint64_t gen, const uint8_t* inputs, const uint8_t* aux) {
    static const bool tl = getenv("ZA_DEBUG_TIMELINE") != nullptr;
    double t_start = tl ? now_ms() : 0, t_wit = 0, t_h = 0, t_enq = 0, t_col = 0;
    ZA_CUDA(cudaSetDevice(P->devices[k]));
    Ctx* cx = &P->ctx[k]->c;
    if (tl) {
        if (!cx->dbg_t0) cudaEventCreate(&cx->dbg_t0);
        cudaEventRecord(cx->dbg_t0, cx->stream);
        cx->dbg_t0_valid = true;
    }
    const Pk* pk = P->pk[k]->p.get();
    const Circuit* c = P->circ[k]->c.get();
    const uint8_t* d_wit = (const uint8_t*)P->wit[k].p;
    Fr* d_h = P->h[k].as<Fr>();
    const size_t m = domain_size(c, nullptr);
    const int n = P->n;
    auto announce = [&](bool failed) {
        { std::lock_guard<std::mutex> lk(P->h_mu); P->h_issued = gen; if (failed) P->h_failed = true; }
        P->h_cv.notify_all();
    };
    try {
        if (inputs) prover_upload_witness(P, k, inputs, aux);
        if (k == 0) {
            try {
                canonical_check_enqueue(cx, d_wit, (size_t)c->ni + c->na, CHK_WITNESS);
                // The h slice of device j is WRITTEN into j's memory by the last pass of the last transform (stores over
                // NVLink through the peer mapping): no copy sits between the H pipeline and the peers' H multiexps.
                // Without peer access to every device (or ZA_PROVER_H_COPY=1) the slices travel as copies afterwards.
                static const bool force_copy = getenv("ZA_PROVER_H_COPY") != nullptr;
                int log_m = 0; domain_size(c, &log_m);
                const bool direct = !force_copy && P->peers_mapped && n <= ZA_H_SCATTER_MAX && log_m > 11;
                if (direct) {
                    cx->h_scatter.n = n;
                    for (int j = 0; j < n; j++) {                  // the plan's H ranges ascend with the device index
                        const QueryRanges qj = query_ranges(P->pk[j]->p.get(), c, m, j, n);
                        cx->h_scatter.out[j] = P->h[j].as<Fr>();
                        cx->h_scatter.hi[j] = (uint32_t)qj.hi[Q_H];
                    }
                }
                try { prove_h(cx, c, d_wit, d_h, nullptr); } catch (...) { cx->h_scatter.n = 0; throw; }
                cx->h_scatter.n = 0;
                for (int j = 1; j < n; j++) {
                    const QueryRanges qj = query_ranges(P->pk[j]->p.get(), c, m, j, n);
                    const size_t lo = qj.lo[Q_H], hi = qj.hi[Q_H];
                    if (!direct && hi > lo) ZA_CUDA(cudaMemcpyPeerAsync(P->h[j].as<Fr>() + lo, P->devices[j], d_h + lo, P->devices[0], (hi - lo) * sizeof(Fr), cx->stream));
                    ZA_CUDA(cudaEventRecord(P->h_ready[j], cx->stream));
                }
            } catch (...) { announce(true); throw; }
            announce(false);
            if (tl) t_h = now_ms();
            prove_msms_enqueue(cx, pk, c, d_wit, d_h, 0, n, MSM_WITNESS);
            if (tl) t_wit = now_ms();
            prove_msms_enqueue(cx, pk, c, d_wit, d_h, 0, n, MSM_H);
        } else {
            prove_msms_enqueue(cx, pk, c, d_wit, d_h, k, n, MSM_WITNESS);
            if (tl) t_wit = now_ms();
            bool failed;
            {
                std::unique_lock<std::mutex> lk(P->h_mu);
                P->h_cv.wait(lk, [&] { return P->h_issued >= gen; });
                failed = P->h_failed;
            }
            if (failed) throw ZaError(ZA_ERR_INVALID, "the H pipeline on the first device failed");
            if (tl) t_h = now_ms();
            ZA_CUDA(cudaStreamWaitEvent(cx->stream, P->h_ready[k], 0));
            prove_msms_enqueue(cx, pk, c, d_wit, d_h, k, n, MSM_H);
        }
        if (tl) t_enq = now_ms();
        prove_msms_collect(cx, P->partials[k]);
        if (tl) {
            t_col = now_ms();
            fprintf(stderr, "[za prover] dev %d host ms since job start: witness msms enqueued %.3f, h issued/seen %.3f, all enqueued %.3f, collected %.3f\n", k,
                    t_wit - t_start, t_h - t_start, t_enq - t_start, t_col - t_start);
            print_timeline(cx);
        }
        if (k == 0) {
            ZA_CUDA(cudaStreamSynchronize(cx->stream));
            const unsigned long long bad = canonical_check_verdict(cx, CHK_WITNESS);
            if (bad != ~0ull) {
                char b[128];
                if (bad < c->ni) snprintf(b, sizeof b, "inputs[%llu] is not a canonical Fr element (>= r)", bad);
                else snprintf(b, sizeof b, "aux[%llu] is not a canonical Fr element (>= r)", bad - c->ni);
                throw ZaError(ZA_ERR_NOT_CANONICAL, b);
            }
        }
    } catch (...) {
        msm_abort(cx);
        if (k == 0 && cx->host_flag && cx->host_flag[CHK_WITNESS] != ~0ull) {
            const unsigned long long bad = canonical_check_verdict(cx, CHK_WITNESS);
            char b[128];
            if (bad < c->ni) snprintf(b, sizeof b, "inputs[%llu] is not a canonical Fr element (>= r)", bad);
            else snprintf(b, sizeof b, "aux[%llu] is not a canonical Fr element (>= r)", bad - c->ni);
            throw ZaError(ZA_ERR_NOT_CANONICAL, b);
        }
        throw;
    }
}

static void prover_create_proof(Prover* P, const uint8_t* inputs, const uint8_t* aux, const uint8_t* r_le, const uint8_t* s_le, uint8_t* proof_out) {
    if (P->circ.empty() || P->pk.empty()) throw ZaError(ZA_ERR_INVALID, "za_prover: load a proving key and a circuit first");
    const Pk* pk0 = P->pk[0]->p.get();
    const Circuit* c0 = P->circ[0]->c.get();
    prove_validate(pk0, r_le, s_le);
    if (P->n == 1) {
        Ctx* cx = &P->ctx[0]->c;
        ZA_CUDA(cudaSetDevice(P->devices[0]));
        if (inputs) prover_upload_witness(P, 0, inputs, aux);
        create_proof_device(cx, pk0, c0, (const uint8_t*)P->wit[0].p, r_le, s_le, proof_out, nullptr);
        return;
    }
    const uint64_t gen = ++P->generation;
    { std::lock_guard<std::mutex> lk(P->h_mu); P->h_failed = false; }
    static const bool tl = getenv("ZA_DEBUG_TIMELINE") != nullptr;
    const double t0 = tl ? now_ms() : 0;
    AssemblePre pre;
    double t_pre = 0;
    prover_run_all(P, [&](int k) { prover_device_step(P, k, gen, inputs, aux); }, [&] { pre = prove_assemble_pre(pk0, r_le, s_le); if (tl) t_pre = now_ms(); });
    const double t1 = tl ? now_ms() : 0;
    Partials sum = P->partials[0];
    for (int k = 1; k < P->n; k++) {
        for (int i = 0; i < 6; i++) xyzz_add<Fq>(sum.g1[i], P->partials[k].g1[i]);
        for (int i = 0; i < 2; i++) xyzz_add<Fq2>(sum.g2[i], P->partials[k].g2[i]);
    }
    prove_assemble_post(pre, sum, proof_out);
    if (tl) fprintf(stderr, "[za prover] proof %llu: host pre %.3f ms, devices done %.3f, combine + assemble %.3f\n", (unsigned long long)gen, t_pre - t0, t1 - t0, now_ms() - t1);
}

// after the key and the circuit are on every device: point ranges, tables of the ranges, witness / h buffers
static void prover_finish_setup(Prover* P) {
    if (P->circ.empty() || P->pk.empty()) return;
    const Circuit* c0 = P->circ[0]->c.get();
    const size_t m = domain_size(c0, nullptr);
    if (P->n > 1) {
        // device 0 runs the H pipeline (rho = its time / the time of the witness multiexps on one GPU; measured 2.1 ms /
        // 18 ms at 2^20) while the others already accumulate: its share of the witness multiexps shrinks accordingly
        double rho = 0.115;
        if (const char* e = getenv("ZA_PROVER_H_RATIO")) { double v = atof(e); if (v >= 0 && v < 1) rho = v; }
        double w = (1.0 - rho * (P->n - 1)) / (1.0 + rho);
        if (w < 0.05) w = 0.05;
        if (w > 1.0) w = 1.0;
        P->rank0_weight = (uint32_t)(w * 1000.0 + 0.5);
        if (P->rank0_weight < 1) P->rank0_weight = 1;
    } else P->rank0_weight = 1000;
    static const bool use_plan = !(getenv("ZA_PROVER_PLAN") && atoi(getenv("ZA_PROVER_PLAN")) == 0);
    std::vector<DevicePlan> plans;
    if (P->n > 1 && use_plan) {
        const size_t qcnt[5] = {m - 1, c0->na, c0->a_cat_total, c0->b_cat_total, c0->b_cat_total};
        plans = prover_make_plans(qcnt, m, P->n);
        const double l0 = (double)(plans[0].hi[Q_L] - plans[0].lo[Q_L]) + (plans[0].hi[Q_A] - plans[0].lo[Q_A]) + (plans[0].hi[Q_B1] - plans[0].lo[Q_B1]) +
                          2.8 * (double)(plans[0].hi[Q_B2] - plans[0].lo[Q_B2]);
        const double l1 = (double)(plans[1].hi[Q_L] - plans[1].lo[Q_L]) + (plans[1].hi[Q_A] - plans[1].lo[Q_A]) + (plans[1].hi[Q_B1] - plans[1].lo[Q_B1]) +
                          2.8 * (double)(plans[1].hi[Q_B2] - plans[1].lo[Q_B2]);
        P->rank0_weight = (uint32_t)std::max(1.0, std::min(1000.0, l1 > 0 ? 1000.0 * l0 / l1 : 1000.0));      // reported by za_prover_info only
    }
    prover_run_all(P, [&](int k) {
        ZA_CUDA(cudaSetDevice(P->devices[k]));
        const Circuit* c = P->circ[k]->c.get();
        if (P->n > 1 && use_plan) {
            uint64_t lo[5], hi[5];
            for (int i = 0; i < 5; i++) { lo[i] = plans[k].lo[i]; hi[i] = plans[k].hi[i]; }
            int rc = za_pk_partition_ranges(P->ctx[k], P->pk[k], P->circ[k], lo, hi);
            if (rc != ZA_OK) throw ZaError(rc, last_error());
        } else if (P->n > 1) {
            int rc = za_pk_partition_weighted(P->ctx[k], P->pk[k], P->circ[k], k, P->n, P->rank0_weight);
            if (rc != ZA_OK) throw ZaError(rc, last_error());
        } else check_query_lengths(P->pk[k]->p.get(), c, m);
        P->wit[k].ensure(((size_t)c->ni + c->na) * 32);
        P->h[k].ensure(m * sizeof(Fr));
    });
    prover_plan_spans(P);
}

}  // namespace za

struct za_prover { za::Prover p; };

using namespace za;

#define ZA_TRY try {
#define ZA_CATCH                                                                   \
    }                                                                              \
    catch (const ZaError& e) { return fail(e.code, "%s", e.what()); }              \
    catch (const CudaError& e) { return fail(ZA_ERR_CUDA, "%s", e.what()); }       \
    catch (const std::bad_alloc&) { return fail(ZA_ERR_INVALID, "out of host memory"); } \
    catch (const std::exception& e) { return fail(ZA_ERR_INVALID, "%s", e.what()); }

extern "C" {

int za_bases_upload(za_ctx* ctx, int group, const uint8_t* bases, size_t n, za_bases** out) {
    if (!ctx || !out || (n && !bases)) return fail(ZA_ERR_INVALID, "NULL argument");
    if (group != 1 && group != 2) return fail(ZA_ERR_INVALID, "group must be 1 (G1) or 2 (G2)");
    *out = nullptr;
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    za_bases* h = new za_bases();
    try { h->b = bases_from_le(&ctx->c, group, bases, n, true, group == 1 ? "G1 bases" : "G2 bases"); }
    catch (...) { delete h; throw; }
    *out = h;
    return ZA_OK;
    ZA_CATCH
}
void za_bases_free(za_bases* b) { delete b; }
size_t za_bases_len(const za_bases* b) { return b ? b->b->n : 0; }

static int multiexp_common(za_ctx* ctx, const za_bases* bases, size_t offset, const uint32_t* d_scalars, size_t n, uint8_t* out, bool partial, bool check) {
    Ctx* c = &ctx->c;
    if (bases->b->group == 1) {
        G1XYZZ r = multiexp_dev<Fq>(c, bases->b.get(), offset, d_scalars, n, check);
        if (partial) xyzz_to_le(r, out); else g1_to_le(xyzz_to_affine<Fq>(r), out);
    } else {
        G2XYZZ r = multiexp_dev<Fq2>(c, bases->b.get(), offset, d_scalars, n, check);
        if (partial) xyzz_to_le(r, out); else g2_to_le(xyzz_to_affine<Fq2>(r), out);
    }
    return ZA_OK;
}

int za_multiexp(za_ctx* ctx, const za_bases* bases, size_t offset, const uint8_t* scalars, size_t n_exp, const uint8_t* density, uint8_t* out) {
    if (!ctx || !bases || !out || (n_exp && !scalars)) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    Ctx* c = &ctx->c;
    ZA_CUDA(cudaSetDevice(c->device));
    check_scalars_canonical(scalars, n_exp, "scalars");
    // density-filtered exponent vector: exponent i is paired with the k(i)-th base of the query
    std::vector<uint8_t> compact;
    const uint8_t* src = scalars;
    size_t n = n_exp;
    if (density) {
        size_t cnt = 0;
        for (size_t i = 0; i < n_exp; i++) cnt += density[i] != 0;
        compact.resize(cnt * 32 + 1);
        size_t k = 0;
        for (size_t i = 0; i < n_exp; i++) if (density[i]) { memcpy(compact.data() + 32 * k, scalars + 32 * i, 32); k++; }
        src = compact.data(); n = cnt;
    }
    DevBuf& d = c->scratch[8];
    d.ensure(n * 32);
    if (n) ZA_CUDA(cudaMemcpyAsync(d.p, src, n * 32, cudaMemcpyHostToDevice, c->stream));
    return multiexp_common(ctx, bases, offset, d.as<uint32_t>(), n, out, false, false);
    ZA_CATCH
}

int za_multiexp_device(za_ctx* ctx, const za_bases* bases, size_t offset, const void* d_scalars, size_t n, uint8_t* out) {
    if (!ctx || !bases || !out || (n && !d_scalars)) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    return multiexp_common(ctx, bases, offset, (const uint32_t*)d_scalars, n, out, false, true);
    ZA_CATCH
}

int za_multiexp_partial_device(za_ctx* ctx, const za_bases* bases, size_t offset, const void* d_scalars, size_t n, uint8_t* out_xyzz) {
    if (!ctx || !bases || !out_xyzz || (n && !d_scalars)) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    return multiexp_common(ctx, bases, offset, (const uint32_t*)d_scalars, n, out_xyzz, true, true);
    ZA_CATCH
}

int za_point_sum(int group, const uint8_t* xyzz, size_t count, uint8_t* out_affine) {
    if (!out_affine || (count && !xyzz)) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    if (group == 1) {
        G1XYZZ acc = G1XYZZ::inf();
        for (size_t i = 0; i < count; i++) xyzz_add<Fq>(acc, g1_xyzz_from_le(xyzz + 128 * i));
        g1_to_le(xyzz_to_affine<Fq>(acc), out_affine);
    } else if (group == 2) {
        G2XYZZ acc = G2XYZZ::inf();
        for (size_t i = 0; i < count; i++) xyzz_add<Fq2>(acc, g2_xyzz_from_le(xyzz + 256 * i));
        g2_to_le(xyzz_to_affine<Fq2>(acc), out_affine);
    } else return fail(ZA_ERR_INVALID, "group must be 1 or 2");
    return ZA_OK;
    ZA_CATCH
}

int za_pk_load(za_ctx* ctx, const uint8_t* params, size_t len, int checked, za_pk** out) {
    if (!ctx || !params || !out) return fail(ZA_ERR_INVALID, "NULL argument");
    *out = nullptr;
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    za_pk* h = new za_pk();
    try { h->p = pk_load(&ctx->c, params, len, checked != 0); }
    catch (...) { delete h; throw; }
    *out = h;
    return ZA_OK;
    ZA_CATCH
}
void za_pk_free(za_pk* pk) { delete pk; }
int za_pk_counts(const za_pk* pk, uint32_t* counts) {
    if (!pk || !counts) return fail(ZA_ERR_INVALID, "NULL argument");
    const Pk* p = pk->p.get();
    counts[0] = (uint32_t)p->ic.size(); counts[1] = (uint32_t)p->h->n; counts[2] = (uint32_t)p->l->n;
    counts[3] = (uint32_t)p->a->n; counts[4] = (uint32_t)p->b_g1->n; counts[5] = (uint32_t)p->b_g2->n;
    return ZA_OK;
}
int za_pk_vk(const za_pk* pk, uint8_t* vk_out, size_t size) {
    if (!pk || !vk_out) return fail(ZA_ERR_INVALID, "NULL argument");
    const Pk* p = pk->p.get();
    size_t need = 64 * 3 + 128 * 3 + 64 * p->ic.size();
    if (size < need) return fail(ZA_ERR_BUFFER_TOO_SMALL, "vk buffer too small: need %zu bytes", need);
    uint8_t* w = vk_out;
    g1_to_le(p->alpha_g1, w); w += 64; g1_to_le(p->beta_g1, w); w += 64; g2_to_le(p->beta_g2, w); w += 128;
    g2_to_le(p->gamma_g2, w); w += 128; g1_to_le(p->delta_g1, w); w += 64; g2_to_le(p->delta_g2, w); w += 128;
    for (const G1Affine& q : p->ic) { g1_to_le(q, w); w += 64; }
    return ZA_OK;
}

int za_circuit_upload(za_ctx* ctx, const za_r1cs* cs, za_circuit** out) {
    if (!ctx || !cs || !out) return fail(ZA_ERR_INVALID, "NULL argument");
    *out = nullptr;
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    za_circuit* h = new za_circuit();
    try { h->c = circuit_upload(&ctx->c, cs); }
    catch (...) { delete h; throw; }
    *out = h;
    return ZA_OK;
    ZA_CATCH
}
void za_circuit_free(za_circuit* c) { delete c; }

int za_bases_generate(za_ctx* ctx, int group, size_t n, uint64_t first_multiple, za_bases** out) {
    if (!ctx || !out) return fail(ZA_ERR_INVALID, "NULL argument");
    if (group != 1 && group != 2) return fail(ZA_ERR_INVALID, "group must be 1 (G1) or 2 (G2)");
    *out = nullptr;
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    za_bases* h = new za_bases();
    try { h->b = bases_generated(&ctx->c, group, n, first_multiple); }
    catch (...) { delete h; throw; }
    *out = h;
    return ZA_OK;
    ZA_CATCH
}

int za_bases_precompute(za_ctx* ctx, za_bases* bases) {
    if (!ctx || !bases) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    if (!bases->b->tab_c) bases_build_table(&ctx->c, bases->b.get());
    return bases->b->tab_c;
    ZA_CATCH
}

int za_bases_download(za_ctx* ctx, const za_bases* bases, size_t offset, size_t n, uint8_t* out) {
    if (!ctx || !bases || !out) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    const Bases* b = bases->b.get();
    if (offset > b->n || n > b->n - offset) return fail(ZA_ERR_INVALID, "range outside the bases array");
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    if (b->group == 1) {
        std::vector<G1Affine> h(n ? n : 1);
        ZA_CUDA(cudaMemcpy(h.data(), b->pts.as<G1Affine>() + offset, n * sizeof(G1Affine), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++) g1_to_le(h[i], out + 64 * i);
    } else {
        std::vector<G2Affine> h(n ? n : 1);
        ZA_CUDA(cudaMemcpy(h.data(), b->pts.as<G2Affine>() + offset, n * sizeof(G2Affine), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++) g2_to_le(h[i], out + 128 * i);
    }
    return ZA_OK;
    ZA_CATCH
}

int za_pk_synthetic(za_ctx* ctx, const uint32_t* counts, za_pk** out) {
    if (!ctx || !counts || !out) return fail(ZA_ERR_INVALID, "NULL argument");
    *out = nullptr;
    ZA_TRY
    Ctx* c = &ctx->c;
    ZA_CUDA(cudaSetDevice(c->device));
    za_pk* h = new za_pk();
    try {
        h->p.reset(new Pk());
        Pk* pk = h->p.get();
        pk->ctx = c;
        Affine<Fq> g1 = host_g1_gen(); Affine<Fq2> g2 = host_g2_gen();
        pk->alpha_g1 = host_multiple<Fq>(g1, 3); pk->beta_g1 = host_multiple<Fq>(g1, 5); pk->delta_g1 = host_multiple<Fq>(g1, 11);
        pk->beta_g2 = host_multiple<Fq2>(g2, 5); pk->gamma_g2 = host_multiple<Fq2>(g2, 7); pk->delta_g2 = host_multiple<Fq2>(g2, 11);
        pk->ic.resize(counts[0]);
        for (uint32_t i = 0; i < counts[0]; i++) pk->ic[i] = host_multiple<Fq>(g1, 13 + i);
        // query q, entry i is the multiple  (q+1) * 2^32 + i + 1  of the generator
        pk->h = bases_generated(c, 1, counts[1], ((uint64_t)1 << 32) + 1);
        pk->l = bases_generated(c, 1, counts[2], ((uint64_t)2 << 32) + 1);
        pk->a = bases_generated(c, 1, counts[3], ((uint64_t)3 << 32) + 1);
        pk->b_g1 = bases_generated(c, 1, counts[4], ((uint64_t)4 << 32) + 1);
        pk->b_g2 = bases_generated(c, 2, counts[5], ((uint64_t)4 << 32) + 1);
        bases_build_table(c, pk->h.get());
        const size_t lo3[3] = {0, 0, 0}, n3[3] = {pk->l->n, pk->a->n, pk->b_g1->n};
        pk_build_witness_tables(c, pk, lo3, n3);
    } catch (...) { delete h; throw; }
    *out = h;
    return ZA_OK;
    ZA_CATCH
}

// One process per GPU: rebuild the fixed-base tables of `pk` for the point range rank `rank` of `world` owns in
// every query (the shares za_prove_msm_partials uses), with the window size chosen for the share.
int za_share_weighted(uint64_t count, int rank, int world, uint32_t rank0_weight_permille, uint64_t* lo, uint64_t* hi) {
    if (!lo || !hi || world < 1 || rank < 0 || rank >= world || rank0_weight_permille < 1 || rank0_weight_permille > 1000)
        return fail(ZA_ERR_INVALID, "bad argument");
    size_t a, b;
    share_weighted((size_t)count, rank, world, rank0_weight_permille, a, b);
    *lo = a; *hi = b;
    return ZA_OK;
}

int za_pk_partition_weighted(za_ctx* ctx, za_pk* pk, const za_circuit* circuit, int rank, int world, uint32_t rank0_weight_permille) {
    if (!ctx || !pk || !circuit) return fail(ZA_ERR_INVALID, "NULL argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(ZA_ERR_INVALID, "bad rank/world");
    if (rank0_weight_permille < 1 || rank0_weight_permille > 1000) return fail(ZA_ERR_INVALID, "rank-0 weight must be in [1, 1000] per mille");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    Pk* p = pk->p.get();
    const Circuit* c = circuit->c.get();
    const size_t m = domain_size(c, nullptr);
    check_query_lengths(p, c, m);
    p->rank0_weight = rank0_weight_permille;
    p->plan_set = false;
    const uint32_t w0 = world > 1 ? rank0_weight_permille : 1000u;
    size_t lo, hi;
    share(m - 1, rank, world, lo, hi); bases_build_table(&ctx->c, p->h.get(), lo, hi - lo);
    size_t los[3], ns[3];
    share_weighted(c->na, rank, world, w0, lo, hi); los[0] = lo; ns[0] = hi - lo;
    share_weighted(c->a_cat_total, rank, world, w0, lo, hi); los[1] = lo; ns[1] = hi - lo;
    share_weighted(c->b_cat_total, rank, world, w0, lo, hi); los[2] = lo; ns[2] = hi - lo;
    pk_build_witness_tables(&ctx->c, p, los, ns);
    return ZA_OK;
    ZA_CATCH
}
int za_pk_partition_ranges(za_ctx* ctx, za_pk* pk, const za_circuit* circuit, const uint64_t* lo, const uint64_t* hi) {
    if (!ctx || !pk || !circuit) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    Pk* p = pk->p.get();
    const Circuit* c = circuit->c.get();
    const size_t m = domain_size(c, nullptr);
    check_query_lengths(p, c, m);
    if (!lo || !hi) {                               // back to the whole queries
        p->plan_set = false;
        return za_pk_partition_weighted(ctx, pk, circuit, 0, 1, 1000);
    }
    const size_t total[5] = {m - 1, c->na, c->a_cat_total, c->b_cat_total, c->b_cat_total};
    for (int i = 0; i < 5; i++)
        if (lo[i] > hi[i] || hi[i] > total[i]) return fail(ZA_ERR_INVALID, "za_pk_partition_ranges: range %d is [%llu, %llu) of %zu", i, (unsigned long long)lo[i], (unsigned long long)hi[i], total[i]);
    for (int i = 0; i < 5; i++) { p->plan_lo[i] = (size_t)lo[i]; p->plan_hi[i] = (size_t)hi[i]; }
    p->plan_set = true;
    bases_build_table(&ctx->c, p->h.get(), p->plan_lo[Q_H], p->plan_hi[Q_H] - p->plan_lo[Q_H]);
    // L, A and B (G1) get a joint window when their ranges are of similar size (pk_build_witness_tables); B (G2) follows
    // the B (G1) window when it covers the same range (shared digit sort) and gets its own otherwise
    size_t los[3] = {p->plan_lo[Q_L], p->plan_lo[Q_A], p->plan_lo[Q_B1]};
    size_t ns[3] = {p->plan_hi[Q_L] - p->plan_lo[Q_L], p->plan_hi[Q_A] - p->plan_lo[Q_A], p->plan_hi[Q_B1] - p->plan_lo[Q_B1]};
    pk_build_witness_tables(&ctx->c, p, los, ns);
    if (p->plan_lo[Q_B2] != p->plan_lo[Q_B1] || p->plan_hi[Q_B2] != p->plan_hi[Q_B1])
        bases_build_table(&ctx->c, p->b_g2.get(), p->plan_lo[Q_B2], p->plan_hi[Q_B2] - p->plan_lo[Q_B2]);
    return ZA_OK;
    ZA_CATCH
}
int za_prover_plan(const za_circuit* circuit, int n_devices, uint64_t* lo_out, uint64_t* hi_out) {
    if (!circuit || !lo_out || !hi_out || n_devices < 1 || n_devices > 64) return fail(ZA_ERR_INVALID, "bad argument");
    ZA_TRY
    const Circuit* c = circuit->c.get();
    const size_t m = domain_size(c, nullptr);
    const size_t qcnt[5] = {m - 1, c->na, c->a_cat_total, c->b_cat_total, c->b_cat_total};
    const std::vector<DevicePlan> plans = prover_make_plans(qcnt, m, n_devices);
    for (int k = 0; k < n_devices; k++)
        for (int i = 0; i < 5; i++) { lo_out[5 * k + i] = plans[k].lo[i]; hi_out[5 * k + i] = plans[k].hi[i]; }
    return ZA_OK;
    ZA_CATCH
}
int za_prover_plan_counts(const uint64_t* counts, uint64_t domain, int n_devices, uint64_t* lo_out, uint64_t* hi_out) {
    if (!counts || !lo_out || !hi_out || n_devices < 1 || n_devices > 64 || domain < 2) return fail(ZA_ERR_INVALID, "bad argument");
    ZA_TRY
    const size_t qcnt[5] = {(size_t)counts[0], (size_t)counts[1], (size_t)counts[2], (size_t)counts[3], (size_t)counts[4]};
    const std::vector<DevicePlan> plans = prover_make_plans(qcnt, (size_t)domain, n_devices);
    for (int k = 0; k < n_devices; k++)
        for (int i = 0; i < 5; i++) { lo_out[5 * k + i] = plans[k].lo[i]; hi_out[5 * k + i] = plans[k].hi[i]; }
    return ZA_OK;
    ZA_CATCH
}
int za_pk_partition(za_ctx* ctx, za_pk* pk, const za_circuit* circuit, int rank, int world) {
    return za_pk_partition_weighted(ctx, pk, circuit, rank, world, 1000);
}

int za_circuit_satisfied(za_ctx* ctx, const za_circuit* circuit, const uint8_t* inputs, const uint8_t* aux, int64_t* first_bad) {
    if (!ctx || !circuit || !inputs || !first_bad) return fail(ZA_ERR_INVALID, "NULL argument");
    if (circuit->c->na && !aux) return fail(ZA_ERR_INVALID, "aux is NULL");
    ZA_TRY
    Ctx* cx = &ctx->c;
    const Circuit* c = circuit->c.get();
    ZA_CUDA(cudaSetDevice(cx->device));
    cudaStream_t st = cx->stream;
    const uint32_t ni = c->ni, na = c->na, nc = c->nc;
    check_scalars_canonical(inputs, ni, "inputs"); check_scalars_canonical(aux, na, "aux");
    *first_bad = -1;
    if (!nc) return ZA_OK;
    DevBuf w(((size_t)ni + na) * 32), abc(3 * (size_t)nc * 32), flag(4);
    ZA_CUDA(cudaMemcpyAsync(w.p, inputs, (size_t)ni * 32, cudaMemcpyHostToDevice, st));
    if (na) ZA_CUDA(cudaMemcpyAsync((uint8_t*)w.p + (size_t)ni * 32, aux, (size_t)na * 32, cudaMemcpyHostToDevice, st));
    fr_convert(cx, w.as<Fr>(), (size_t)ni + na, 0);
    Fr* outs[3] = {abc.as<Fr>(), abc.as<Fr>() + nc, abc.as<Fr>() + 2 * (size_t)nc};
    for (int k = 0; k < 3; k++)
        r1cs_eval_kernel<<<nblk(nc, 128), 128, 0, st>>>(c->ptr[k].as<uint32_t>(), c->col[k].as<uint32_t>(), c->coeff[k].as<Fr>(), w.as<Fr>(), nc, outs[k]);
    ZA_CUDA(cudaMemsetAsync(flag.p, 0xff, 4, st));
    r1cs_check_kernel<<<nblk(nc, 128), 128, 0, st>>>(outs[0], outs[1], outs[2], nc, flag.as<uint32_t>());
    cx->launches += 4;
    ZA_CUDA(cudaGetLastError());
    uint32_t h = 0;
    ZA_CUDA(cudaMemcpyAsync(&h, flag.p, 4, cudaMemcpyDeviceToHost, st));
    ZA_CUDA(cudaStreamSynchronize(st));
    if (h != 0xffffffffu) *first_bad = (int64_t)h;
    return ZA_OK;
    ZA_CATCH
}

int za_circuit_info(const za_circuit* circuit, uint32_t* info) {
    if (!circuit || !info) return fail(ZA_ERR_INVALID, "NULL argument");
    const Circuit* c = circuit->c.get();
    info[0] = c->ni; info[1] = c->na; info[2] = c->nc; info[3] = c->a_aux_total; info[4] = c->b_in_total; info[5] = c->b_aux_total;
    int log_m = 0;
    ZA_TRY
    domain_size(c, &log_m);
    info[6] = (uint32_t)log_m;
    return ZA_OK;
    ZA_CATCH
}

int za_imad_peak(za_ctx* ctx, double* imads_per_second) {
    if (!ctx || !imads_per_second) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    *imads_per_second = imad_peak(&ctx->c);
    return ZA_OK;
    ZA_CATCH
}

int za_prove_h_device(za_ctx* ctx, const za_circuit* circuit, const void* d_witness, void* d_h) {
    if (!ctx || !circuit || !d_witness || !d_h) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    const Circuit* c = circuit->c.get();
    struct Clear { Ctx* c; ~Clear() { c->h_scatter.n = 0; } } clear{&ctx->c};
    if (ctx->c.h_scatter.n) {
        int log_m = 0; domain_size(c, &log_m);
        if (log_m <= 11) throw ZaError(ZA_ERR_INVALID, "za_ctx_set_h_scatter needs a domain of 2^12 or more elements");
    }
    canonical_check_enqueue(&ctx->c, d_witness, (size_t)c->ni + c->na, CHK_WITNESS);      // verdict: za_prove_msm_collect / _partials
    prove_h(&ctx->c, c, (const uint8_t*)d_witness, (Fr*)d_h, nullptr);
    return ZA_OK;
    ZA_CATCH
}
int za_ctx_set_h_scatter(za_ctx* ctx, int n, void* const* outs, const uint64_t* his) {
    if (!ctx || n < 0 || n > ZA_H_SCATTER_MAX || (n && (!outs || !his))) return fail(ZA_ERR_INVALID, "bad argument");
    for (int j = 0; j < n; j++) {
        if (!outs[j] || his[j] > 0xffffffffull || (j && his[j] < his[j - 1])) return fail(ZA_ERR_INVALID, "h scatter: NULL destination or bounds not ascending");
    }
    ctx->c.h_scatter.n = n;
    for (int j = 0; j < n; j++) { ctx->c.h_scatter.out[j] = (Fr*)outs[j]; ctx->c.h_scatter.hi[j] = (uint32_t)his[j]; }
    return ZA_OK;
}

// verdicts of the range checks the staged entry points enqueued on this context (after a synchronisation)
static void staged_check_verdicts(Ctx* ctx, const Circuit* c) {
    ZA_CUDA(cudaStreamSynchronize(ctx->stream));
    const unsigned long long w = canonical_check_verdict(ctx, CHK_WITNESS), h = canonical_check_verdict(ctx, CHK_H);
    char b[128];
    if (w != ~0ull) {
        if (c && w >= c->ni) snprintf(b, sizeof b, "aux[%llu] is not a canonical Fr element (>= r)", w - c->ni);
        else snprintf(b, sizeof b, "witness[%llu] is not a canonical Fr element (>= r)", w);
        throw ZaError(ZA_ERR_NOT_CANONICAL, b);
    }
    if (h != ~0ull) { snprintf(b, sizeof b, "h[%llu] is not a canonical Fr element (>= r)", h); throw ZaError(ZA_ERR_NOT_CANONICAL, b); }
}

int za_prove_msm_partials(za_ctx* ctx, const za_pk* pk, const za_circuit* circuit, const void* d_witness, const void* d_h, int rank, int world,
                          uint8_t* partials_out) {
    if (!ctx || !pk || !circuit || !d_witness || !d_h || !partials_out) return fail(ZA_ERR_INVALID, "NULL argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(ZA_ERR_INVALID, "bad rank/world");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    Partials P;
    const Circuit* c = circuit->c.get();
    canonical_check_enqueue(&ctx->c, d_witness, (size_t)c->ni + c->na, CHK_WITNESS);
    { size_t lo, hi; share(domain_size(c, nullptr) - 1, rank, world, lo, hi); canonical_check_enqueue(&ctx->c, (const Fr*)d_h + lo, hi - lo, CHK_H); }
    try { prove_msms(&ctx->c, pk->p.get(), c, (const uint8_t*)d_witness, (const Fr*)d_h, rank, world, P); }
    catch (...) { msm_abort(&ctx->c); throw; }
    staged_check_verdicts(&ctx->c, c);
    partials_to_le(P, partials_out);
    return ZA_OK;
    ZA_CATCH
}

int za_prove_msm_enqueue(za_ctx* ctx, const za_pk* pk, const za_circuit* circuit, const void* d_witness, const void* d_h, int rank, int world,
                         int which) {
    if (!ctx || !pk || !circuit || !d_witness) return fail(ZA_ERR_INVALID, "NULL argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(ZA_ERR_INVALID, "bad rank/world");
    if (!(which & (MSM_WITNESS | MSM_H)) || ((which & MSM_H) && !d_h)) return fail(ZA_ERR_INVALID, "bad multiexp selection");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    const Circuit* c = circuit->c.get();
    if (which & MSM_WITNESS) canonical_check_enqueue(&ctx->c, d_witness, (size_t)c->ni + c->na, CHK_WITNESS);
    if (which & MSM_H) { size_t lo, hi; share(domain_size(c, nullptr) - 1, rank, world, lo, hi); canonical_check_enqueue(&ctx->c, (const Fr*)d_h + lo, hi - lo, CHK_H); }
    try { prove_msms_enqueue(&ctx->c, pk->p.get(), c, (const uint8_t*)d_witness, (const Fr*)d_h, rank, world, which); }
    catch (...) { msm_abort(&ctx->c); throw; }
    return ZA_OK;
    ZA_CATCH
}

int za_prove_msm_collect(za_ctx* ctx, uint8_t* partials_out) {
    if (!ctx || !partials_out) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    for (int s : {0, 1, 2, 3, 4}) {
        if ((s == 1 || s == 2) && ctx->c.witness_merged && ctx->c.slots[3].busy && ctx->c.slots[3].nparts == 3) continue;   // L and A ride in slot 3
        if (!ctx->c.slots[s].busy) return fail(ZA_ERR_INVALID, "za_prove_msm_collect: multiexp %d was not enqueued", s);
    }
    Partials P;
    try { prove_msms_collect(&ctx->c, P); }
    catch (...) { msm_abort(&ctx->c); throw; }
    staged_check_verdicts(&ctx->c, nullptr);
    partials_to_le(P, partials_out);
    return ZA_OK;
    ZA_CATCH
}

int za_prove_assemble(const za_pk* pk, const uint8_t* partials, int world, const uint8_t* r, const uint8_t* s, uint8_t* proof_out) {
    if (!pk || !partials || !r || !s || !proof_out || world < 1) return fail(ZA_ERR_INVALID, "bad argument");
    ZA_TRY
    Partials P = partials_zero();
    for (int k = 0; k < world; k++) partials_add_le(P, partials + (size_t)k * PARTIALS_BYTES);
    prove_assemble(pk->p.get(), P, r, s, proof_out);
    return ZA_OK;
    ZA_CATCH
}

int za_create_proof_device(za_ctx* ctx, const za_pk* pk, const za_circuit* circuit, const void* d_witness, const uint8_t* r, const uint8_t* s,
                           uint8_t* proof_out) {
    if (!ctx || !pk || !circuit || !d_witness || !r || !s || !proof_out) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    create_proof_device(&ctx->c, pk->p.get(), circuit->c.get(), (const uint8_t*)d_witness, r, s, proof_out, nullptr);
    return ZA_OK;
    ZA_CATCH
}

// ---- several GPUs behind one call
int za_prover_create(const int* devices, int n_devices, za_prover** out) {
    if (!devices || !out || n_devices < 1 || n_devices > 64) return fail(ZA_ERR_INVALID, "bad argument");
    *out = nullptr;
    for (int i = 0; i < n_devices; i++) for (int j = 0; j < i; j++) if (devices[i] == devices[j]) return fail(ZA_ERR_INVALID, "device %d listed twice", devices[i]);
    ZA_TRY
    std::unique_ptr<za_prover> h(new za_prover());
    Prover* P = &h->p;
    P->n = n_devices;
    P->devices.assign(devices, devices + n_devices);
    P->ctx.assign(n_devices, nullptr); P->pk.clear(); P->circ.clear();
    P->wit.resize(n_devices); P->h.resize(n_devices); P->h_ready.assign(n_devices, nullptr); P->partials.resize(n_devices);
    for (int k = 0; k < n_devices; k++) {
        int rc = za_ctx_create(devices[k], &P->ctx[k]);
        if (rc != ZA_OK) return rc;
    }
    // recorded on device 0's stream (an event belongs to the device it was created on); the peers wait on them
    ZA_CUDA(cudaSetDevice(devices[0]));
    for (int k = 0; k < n_devices; k++) ZA_CUDA(cudaEventCreateWithFlags(&P->h_ready[k], cudaEventDisableTiming));
    // device 0 writes the h slices straight into its peers' memory
    P->peers_mapped = n_devices > 1;
    for (int k = 1; k < n_devices; k++) {
        int can = 0;
        bool ok = false;
        if (cudaDeviceCanAccessPeer(&can, devices[0], devices[k]) == cudaSuccess && can) {
            cudaSetDevice(devices[0]);
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[k], 0);
            ok = e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;      // otherwise the h slices travel as (host-staged) copies
            cudaGetLastError();
        }
        if (!ok) P->peers_mapped = false;
    }
    for (int k = 0; k < n_devices; k++) {
        P->workers.emplace_back(new ProverWorker());
        ProverWorker* w = P->workers.back().get();
        w->th = std::thread(prover_worker_loop, w);
    }
    *out = h.release();
    return ZA_OK;
    ZA_CATCH
}
void za_prover_destroy(za_prover* p) { delete p; }
int za_prover_device_count(const za_prover* p) { return p ? p->p.n : 0; }
za_ctx* za_prover_ctx(za_prover* p, int k) { return (p && k >= 0 && k < p->p.n) ? p->p.ctx[k] : nullptr; }

static void prover_drop_pk(Prover* P) {
    for (int k = 0; k < (int)P->pk.size(); k++) { cudaSetDevice(P->devices[k]); delete P->pk[k]; }
    P->pk.clear();
}
int za_prover_load_pk(za_prover* p, const uint8_t* params, size_t len, int checked) {
    if (!p || !params) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    Prover* P = &p->p;
    prover_drop_pk(P);
    P->pk.assign(P->n, nullptr);
    try {
        prover_run_all(P, [&](int k) {
            ZA_CUDA(cudaSetDevice(P->devices[k]));
            std::unique_ptr<za_pk> h(new za_pk());
            h->p = pk_load(&P->ctx[k]->c, params, len, checked != 0);
            P->pk[k] = h.release();
        });
        prover_finish_setup(P);
    } catch (...) { prover_drop_pk(P); throw; }
    return ZA_OK;
    ZA_CATCH
}
int za_prover_synthetic_pk(za_prover* p, const uint32_t* counts) {
    if (!p || !counts) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    Prover* P = &p->p;
    prover_drop_pk(P);
    P->pk.assign(P->n, nullptr);
    try {
        prover_run_all(P, [&](int k) {
            za_pk* h = nullptr;
            int rc = za_pk_synthetic(P->ctx[k], counts, &h);
            if (rc != ZA_OK) throw ZaError(rc, last_error());
            P->pk[k] = h;
        });
        prover_finish_setup(P);
    } catch (...) { prover_drop_pk(P); throw; }
    return ZA_OK;
    ZA_CATCH
}
int za_prover_set_circuit(za_prover* p, const za_r1cs* cs) {
    if (!p || !cs) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    Prover* P = &p->p;
    for (int k = 0; k < (int)P->circ.size(); k++) { cudaSetDevice(P->devices[k]); delete P->circ[k]; }
    P->circ.assign(P->n, nullptr);
    try {
        prover_run_all(P, [&](int k) {
            ZA_CUDA(cudaSetDevice(P->devices[k]));
            std::unique_ptr<za_circuit> h(new za_circuit());
            h->c = circuit_upload(&P->ctx[k]->c, cs);
            P->circ[k] = h.release();
        });
        prover_finish_setup(P);
    } catch (...) {
        for (int k = 0; k < (int)P->circ.size(); k++) { cudaSetDevice(P->devices[k]); delete P->circ[k]; }
        P->circ.clear();
        throw;
    }
    return ZA_OK;
    ZA_CATCH
}
int za_prover_vk(const za_prover* p, uint8_t* vk_out, size_t size) {
    if (!p || p->p.pk.empty()) return fail(ZA_ERR_INVALID, "za_prover: no proving key loaded");
    return za_pk_vk(p->p.pk[0], vk_out, size);
}
int za_prover_pk_counts(const za_prover* p, uint32_t* counts) {
    if (!p || p->p.pk.empty()) return fail(ZA_ERR_INVALID, "za_prover: no proving key loaded");
    return za_pk_counts(p->p.pk[0], counts);
}
int za_prover_upload_witness(za_prover* p, const uint8_t* inputs, const uint8_t* aux) {
    if (!p || !inputs) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    Prover* P = &p->p;
    if (P->circ.empty() || P->pk.empty()) return fail(ZA_ERR_INVALID, "za_prover: load a proving key and a circuit first");
    if (P->circ[0]->c->na && !aux) return fail(ZA_ERR_INVALID, "aux is NULL");
    prover_run_all(P, [&](int k) {
        ZA_CUDA(cudaSetDevice(P->devices[k]));
        prover_upload_witness(P, k, inputs, aux);
        ZA_CUDA(cudaStreamSynchronize(P->ctx[k]->c.stream));
    });
    return ZA_OK;
    ZA_CATCH
}
int za_prover_create_proof(za_prover* p, const uint8_t* inputs, const uint8_t* aux, const uint8_t* r, const uint8_t* s, uint8_t* proof_out) {
    if (!p || !r || !s || !proof_out) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    Prover* P = &p->p;
    if (P->circ.empty() || P->pk.empty()) return fail(ZA_ERR_INVALID, "za_prover: load a proving key and a circuit first");
    if (inputs && P->circ[0]->c->na && !aux) return fail(ZA_ERR_INVALID, "aux is NULL");
    prover_create_proof(P, inputs, aux, r, s, proof_out);
    return ZA_OK;
    ZA_CATCH
}
int za_prover_info(const za_prover* p, uint64_t* info) {
    if (!p || !info) return fail(ZA_ERR_INVALID, "NULL argument");
    const Prover* P = &p->p;
    uint64_t h2d = 0;
    for (auto& dev : P->spans) for (auto& sp : dev) h2d += (sp.second - sp.first) * 32;
    info[0] = h2d;                       // witness bytes uploaded per proof, all devices
    info[1] = P->rank0_weight;           // per mille of an ordinary share that device 0 takes of the witness multiexps
    info[2] = (uint64_t)P->n * (4 * 6 * 128 + 6 * 256) + 8;   // bytes read back per proof: 6 partial results per multiexp and device, one verdict
    return ZA_OK;
}
uint64_t za_prover_launch_count(const za_prover* p) {
    uint64_t t = 0;
    if (p) for (za_ctx* c : p->p.ctx) t += c->c.launches;
    return t;
}

int za_create_proof(za_ctx* ctx, const za_pk* pk, const za_circuit* circuit, const uint8_t* inputs, const uint8_t* aux, const uint8_t* r,
                    const uint8_t* s, uint8_t* proof_out, za_trace* trace) {
    if (!ctx || !pk || !circuit || !inputs || !r || !s || !proof_out) return fail(ZA_ERR_INVALID, "NULL argument");
    if (circuit->c->na && !aux) return fail(ZA_ERR_INVALID, "aux is NULL");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    create_proof(&ctx->c, pk->p.get(), circuit->c.get(), inputs, aux, r, s, proof_out, trace);
    return ZA_OK;
    ZA_CATCH
}

}  // extern "C"
