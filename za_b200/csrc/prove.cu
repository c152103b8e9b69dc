// Groth16 create_proof on the GPU, plus the proving-key / circuit / multiexp handles of the C ABI.
//
// Replaces (un-vendored) bellman_ce groth16/prover.rs `create_proof` + `ProvingAssignment`, the
// `ParameterSource for &Parameters` cursor logic, and `Parameters::read` — everything
// /root/reference/prover/src/groth16/prover.rs:173 (`create_random_proof`) and format.rs:285
// (`Parameters::read(pk, true)`) call into.  Step numbers refer to SURVEY.md §3.2.
#include "common.cuh"
#include "api_internal.cuh"
#include <string.h>
#include <memory>

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

// ------------------------------------------------------------------ device-resident bases
struct Bases {
    Ctx* ctx;
    int group;            // 1 = G1, 2 = G2
    size_t n;
    bool has_infinity;
    DevBuf pts;           // Affine<Fq> or Affine<Fq2>, Montgomery form
};

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

template <class F>
static XYZZ<F> multiexp_dev(Ctx* ctx, const Bases* b, size_t offset, const uint32_t* d_scalars, size_t n) {
    if (offset > b->n || n > b->n - offset)
        throw ZaError(ZA_ERR_IO, "multiexp: the base query is shorter than the exponent vector (bellman: unexpected EOF)");
    return msm_run<F>(ctx, b->pts.as<Affine<F>>() + offset, d_scalars, n, b->has_infinity);
}

// ------------------------------------------------------------------ proving key
struct Pk {
    Ctx* ctx;
    // verifying key, host side, Montgomery form
    G1Affine alpha_g1, beta_g1, delta_g1;
    G2Affine beta_g2, gamma_g2, delta_g2;
    std::vector<G1Affine> ic;
    std::unique_ptr<Bases> h, l, a, b_g1, b_g2;
};

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
    return pk;
}

// ------------------------------------------------------------------ circuit (R1CS on device)
// Everything that depends only on the constraint system is computed once here: the CSR matrices in
// Montgomery form, the three density maps of bellman's ProvingAssignment (they depend on which
// variables occur in A / B rows, never on the witness: density.inc(i) fires for every term)
// and the compacted index lists the density-filtered multiexps need (K10).
struct Circuit {
    Ctx* ctx;
    uint32_t ni, na, nc;
    DevBuf ptr[3], col[3], coeff[3];     // col = slot in the witness vector [inputs | aux]
    std::vector<uint8_t> a_aux_density, b_in_density, b_aux_density;
    DevBuf a_aux_idx, b_in_idx, b_aux_idx;
    uint32_t a_aux_total = 0, b_in_total = 0, b_aux_total = 0;
};

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
        if (total) ZA_CUDA(cudaMemcpy(out.p, idx.data(), (size_t)total * 4, cudaMemcpyHostToDevice));
    };
    make_idx(c->a_aux_density, c->a_aux_idx, c->a_aux_total);
    make_idx(c->b_in_density, c->b_in_idx, c->b_in_total);
    make_idx(c->b_aux_density, c->b_aux_idx, c->b_aux_total);
    return c;
}

// ------------------------------------------------------------------ create_proof
static const uint32_t* gather(Ctx* ctx, const uint8_t* d_src, const DevBuf& idx, uint32_t total, DevBuf& dst) {
    dst.ensure((size_t)total * 32);
    if (total) {
        gather_scalars_kernel<<<nblk(total, 256), 256, 0, ctx->stream>>>((const uint4*)d_src, idx.as<uint32_t>(), total, dst.as<uint4>());
        ctx->launches++;
        ZA_CUDA(cudaGetLastError());
    }
    return dst.as<uint32_t>();
}

static void create_proof(Ctx* ctx, const Pk* pk, const Circuit* c, const uint8_t* inputs, const uint8_t* aux, const uint8_t* r_le,
                         const uint8_t* s_le, uint8_t* proof_out, za_trace* tr) {
    cudaStream_t st = ctx->stream;
    const uint32_t ni = c->ni, na = c->na, nc = c->nc;
    check_scalars_canonical(inputs, ni, "inputs"); check_scalars_canonical(aux, na, "aux");
    check_scalars_canonical(r_le, 1, "r"); check_scalars_canonical(s_le, 1, "s");
    const size_t len = (size_t)nc + ni;          // rows incl. the input-consistency rows (step 3)
    size_t m = 1; int log_m = 0;
    while (m < len) { m *= 2; log_m++; if (log_m >= 28) throw ZaError(ZA_ERR_POLY_DEGREE_TOO_LARGE, "create_proof: domain of 2^28 or more elements"); }

    // witness on device: canonical copy (multiexp exponents) and Montgomery copy (row evaluation)
    DevBuf wit_canon(((size_t)ni + na) * 32), wit_mont(((size_t)ni + na) * 32);
    ZA_CUDA(cudaMemcpyAsync(wit_canon.p, inputs, (size_t)ni * 32, cudaMemcpyHostToDevice, st));
    if (na) ZA_CUDA(cudaMemcpyAsync((uint8_t*)wit_canon.p + (size_t)ni * 32, aux, (size_t)na * 32, cudaMemcpyHostToDevice, st));
    ZA_CUDA(cudaMemcpyAsync(wit_mont.p, wit_canon.p, ((size_t)ni + na) * 32, cudaMemcpyDeviceToDevice, st));
    fr_convert(ctx, wit_mont.as<Fr>(), (size_t)ni + na, 0);

    // steps 2-3: a, b, c = <row, witness>; then rows a = input_i, b = c = 0
    DevBuf abc(3 * m * sizeof(Fr));
    Fr* d_a = abc.as<Fr>(); Fr* d_b = d_a + m; Fr* d_c = d_b + m;
    ZA_CUDA(cudaMemsetAsync(abc.p, 0, 3 * m * sizeof(Fr), st));
    Fr* outs[3] = {d_a, d_b, d_c};
    if (nc) {
        for (int w = 0; w < 3; w++) {
            r1cs_eval_kernel<<<nblk(nc, 128), 128, 0, st>>>(c->ptr[w].as<uint32_t>(), c->col[w].as<uint32_t>(), c->coeff[w].as<Fr>(),
                                                          wit_mont.as<Fr>(), nc, outs[w]);
            ctx->launches++;
        }
        ZA_CUDA(cudaGetLastError());
    }
    ZA_CUDA(cudaMemcpyAsync(d_a + nc, wit_mont.p, (size_t)ni * 32, cudaMemcpyDeviceToDevice, st));
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

    // step 4: H polynomial; result canonical in d_a[0 .. m-1)
    h_poly_device(ctx, d_a, d_b, d_c, log_m);
    if (tr && tr->h_coeffs && m > 1) {
        ZA_CUDA(cudaMemcpyAsync(tr->h_coeffs, d_a, (m - 1) * 32, cudaMemcpyDeviceToHost, st));
        ZA_CUDA(cudaStreamSynchronize(st));
    }

    // steps 4-5: the eight multiexps
    const uint8_t* d_in = (const uint8_t*)wit_canon.p;
    const uint8_t* d_aux = d_in + (size_t)ni * 32;
    G1XYZZ H = multiexp_dev<Fq>(ctx, pk->h.get(), 0, (const uint32_t*)d_a, m - 1);
    G1XYZZ Lq = multiexp_dev<Fq>(ctx, pk->l.get(), 0, (const uint32_t*)d_aux, na);
    G1XYZZ A_in = multiexp_dev<Fq>(ctx, pk->a.get(), 0, (const uint32_t*)d_in, ni);
    DevBuf g1buf, g2buf;
    const uint32_t* sc = gather(ctx, d_aux, c->a_aux_idx, c->a_aux_total, g1buf);
    G1XYZZ A_aux = multiexp_dev<Fq>(ctx, pk->a.get(), ni, sc, c->a_aux_total);
    const uint32_t* sc_bin = gather(ctx, d_in, c->b_in_idx, c->b_in_total, g2buf);
    G1XYZZ B1_in = multiexp_dev<Fq>(ctx, pk->b_g1.get(), 0, sc_bin, c->b_in_total);
    G2XYZZ B2_in = multiexp_dev<Fq2>(ctx, pk->b_g2.get(), 0, sc_bin, c->b_in_total);
    const uint32_t* sc_baux = gather(ctx, d_aux, c->b_aux_idx, c->b_aux_total, g1buf);
    G1XYZZ B1_aux = multiexp_dev<Fq>(ctx, pk->b_g1.get(), c->b_in_total, sc_baux, c->b_aux_total);
    G2XYZZ B2_aux = multiexp_dev<Fq2>(ctx, pk->b_g2.get(), c->b_in_total, sc_baux, c->b_aux_total);
    if (tr && tr->msm_g1) {
        const G1XYZZ* v[6] = {&H, &Lq, &A_in, &A_aux, &B1_in, &B1_aux};
        for (int i = 0; i < 6; i++) g1_to_le(xyzz_to_affine<Fq>(*v[i]), tr->msm_g1 + 64 * i);
    }
    if (tr && tr->msm_g2) {
        g2_to_le(xyzz_to_affine<Fq2>(B2_in), tr->msm_g2);
        g2_to_le(xyzz_to_affine<Fq2>(B2_aux), tr->msm_g2 + 128);
    }

    // steps 6-8: assembly on the host (a handful of group operations)
    if (pk->delta_g1.is_inf() || pk->delta_g2.is_inf()) throw ZaError(ZA_ERR_UNEXPECTED_IDENTITY, "create_proof: delta is the point at infinity");
    uint32_t r[8], s[8];
    memcpy(r, r_le, 32); memcpy(s, s_le, 32);
    Fr rf, sf; memcpy(rf.v, r, 32); memcpy(sf.v, s, 32);
    Fr rs_m = fp_to_mont<FrParams>(rf) * fp_to_mont<FrParams>(sf);
    Fr rs_c = fp_from_mont<FrParams>(rs_m);
    G1XYZZ d1 = G1XYZZ::from_affine(pk->delta_g1), al = G1XYZZ::from_affine(pk->alpha_g1), be1 = G1XYZZ::from_affine(pk->beta_g1);
    G2XYZZ d2 = G2XYZZ::from_affine(pk->delta_g2);
    G1XYZZ g_a = xyzz_mul<Fq>(d1, r); xyzz_madd<Fq>(g_a, pk->alpha_g1);
    G2XYZZ g_b = xyzz_mul<Fq2>(d2, s); xyzz_madd<Fq2>(g_b, pk->beta_g2);
    G1XYZZ g_c = xyzz_mul<Fq>(d1, rs_c.v);
    xyzz_add<Fq>(g_c, xyzz_mul<Fq>(al, s));
    xyzz_add<Fq>(g_c, xyzz_mul<Fq>(be1, r));
    G1XYZZ a_ans = A_in; xyzz_add<Fq>(a_ans, A_aux);
    xyzz_add<Fq>(g_a, a_ans);
    xyzz_add<Fq>(g_c, xyzz_mul<Fq>(a_ans, s));
    G1XYZZ b1_ans = B1_in; xyzz_add<Fq>(b1_ans, B1_aux);
    G2XYZZ b2_ans = B2_in; xyzz_add<Fq2>(b2_ans, B2_aux);
    xyzz_add<Fq2>(g_b, b2_ans);
    xyzz_add<Fq>(g_c, xyzz_mul<Fq>(b1_ans, r));
    xyzz_add<Fq>(g_c, H);
    xyzz_add<Fq>(g_c, Lq);
    g1_to_le(xyzz_to_affine<Fq>(g_a), proof_out);
    g2_to_le(xyzz_to_affine<Fq2>(g_b), proof_out + 64);
    g1_to_le(xyzz_to_affine<Fq>(g_c), proof_out + 192);
}

}  // namespace za

using namespace za;

struct za_bases { std::unique_ptr<Bases> b; };
struct za_pk { std::unique_ptr<Pk> p; };
struct za_circuit { std::unique_ptr<Circuit> c; };

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

static int multiexp_common(za_ctx* ctx, const za_bases* bases, size_t offset, const uint32_t* d_scalars, size_t n, uint8_t* out, bool partial) {
    Ctx* c = &ctx->c;
    if (bases->b->group == 1) {
        G1XYZZ r = multiexp_dev<Fq>(c, bases->b.get(), offset, d_scalars, n);
        if (partial) xyzz_to_le(r, out); else g1_to_le(xyzz_to_affine<Fq>(r), out);
    } else {
        G2XYZZ r = multiexp_dev<Fq2>(c, bases->b.get(), offset, d_scalars, n);
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
    DevBuf& d = c->scratch[7];
    d.ensure(n * 32);
    if (n) ZA_CUDA(cudaMemcpyAsync(d.p, src, n * 32, cudaMemcpyHostToDevice, c->stream));
    return multiexp_common(ctx, bases, offset, d.as<uint32_t>(), n, out, false);
    ZA_CATCH
}

int za_multiexp_device(za_ctx* ctx, const za_bases* bases, size_t offset, const void* d_scalars, size_t n, uint8_t* out) {
    if (!ctx || !bases || !out || (n && !d_scalars)) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    return multiexp_common(ctx, bases, offset, (const uint32_t*)d_scalars, n, out, false);
    ZA_CATCH
}

int za_multiexp_partial_device(za_ctx* ctx, const za_bases* bases, size_t offset, const void* d_scalars, size_t n, uint8_t* out_xyzz) {
    if (!ctx || !bases || !out_xyzz || (n && !d_scalars)) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    return multiexp_common(ctx, bases, offset, (const uint32_t*)d_scalars, n, out_xyzz, true);
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
