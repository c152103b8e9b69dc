// Fr number-theoretic transforms for the H polynomial of Groth16.
//
// Replaces (un-vendored) bellman_ce domain.rs: EvaluationDomain::{fft, ifft, coset_fft, icoset_fft,
// mul_assign, sub_assign, divide_by_z_on_coset} and best_fft/serial_fft/parallel_fft, which
// create_proof drives (entered from /root/reference/prover/src/groth16/prover.rs:173).
//
// Design: natural-order-in / natural-order-out decimation-in-frequency transform,
// split into passes.  One pass = one kernel: a CTA stages a tile of 2^K strided elements x 2^LOGT
// adjacent columns (2^11 elements, 64 KiB) in shared memory, each thread keeps 8 elements in
// registers and runs up to 3 radix-2 stages per shared-memory exchange.  Twiddles come from one
// precomputed table omega^e, e in [0, N/2]; the inverse transform reuses the same table through
// omega^-e = -omega^(N/2-e).  The last pass writes through a bit-reversal so the result is in
// natural order; pre-scaling (coset shift) and post-scaling (1/m, coset un-shift) are fused into the
// first load / last store.  HBM traffic per transform = passes x 64 B per element; the binding
// resource is the integer pipe.
#include "common.cuh"
#include <stdlib.h>

namespace za {

static __device__ __forceinline__ Fr ld_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
static __device__ __forceinline__ Fr ldg_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
static __device__ __forceinline__ void st_fr(Fr* p, const Fr& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

struct NttPass {
    const Fr* in;
    Fr* out;
    size_t batch_stride;  // elements between consecutive vectors of a batch
    const Fr* tw;         // omega^e, e in [0, N/2]
    const Fr* pre;        // pre-scale by natural input index (first pass) or nullptr
    const Fr* post;       // post-scale by natural output index (last pass) or nullptr
    const Fr* post_const; // single post-scale constant (last pass) or nullptr
    const Fr* mul_b;      // first pass: x = (x * mul_b[g] - sub_c[g]) * pre_const  (mul_assign, sub_assign and
    const Fr* sub_c;      //             divide_by_z_on_coset of create_proof fused into the load of the last transform)
    const Fr* pre_const;
    int n;                // log2 N
    int lo;               // lowest global index bit of this pass's tile
    int K;                // tile = 2^K elements at stride 2^lo
    int LOGT;             // 2^LOGT columns per CTA
    int inverse;
    int first;            // first pass of the transform (apply `pre`)
    HScatter sc;          // last pass: output index k goes to sc.out[j] + k, j the first part with k < sc.hi[j] (sc.n == 0: to `out`)
};

static __device__ __forceinline__ unsigned rotr3(unsigned j, unsigned r) { return ((j >> r) | (j << (3 - r))) & 7u; }

// LAST = false: in-place style pass, columns are the low LOGT index bits (contiguous in memory).
// LAST = true : final pass (lo = 0), columns are the top LOGT index bits, output is written at the
//               bit-reversed index so the transform result is in natural order.
#ifndef ZA_NTT_MINB
#define ZA_NTT_MINB 2          // two CTAs of 64 KiB per SM: <= 128 registers (ptxas takes 188 for the last pass if allowed; 3 CTAs spill and lose, profiles/r02_ntt.md)
#endif
template <bool LAST>
__global__ void __launch_bounds__(256, ZA_NTT_MINB) ntt_pass_kernel(NttPass p) {
    extern __shared__ uint4 smem_raw[];
    Fr* sm = reinterpret_cast<Fr*>(smem_raw);
    const int n = p.n, lo = p.lo, K = p.K, LOGT = p.LOGT, L = K + LOGT;
    const unsigned t = threadIdx.x;
    const size_t blk = blockIdx.x;
    const Fr* in = p.in + (size_t)blockIdx.y * p.batch_stride;
    Fr* out = p.out + (size_t)blockIdx.y * p.batch_stride;
    const size_t half_n = (size_t)1 << (n - 1);

    // global index of local element l
    size_t g_base;
    if (!LAST) {
        const int s = lo - LOGT;
        size_t blo = blk & (((size_t)1 << s) - 1), bhi = blk >> s;
        g_base = (bhi << (lo + K)) | (blo << LOGT);
    } else {
        g_base = blk << K;
    }
    auto gidx = [&](unsigned l) -> size_t {
        if (!LAST) return g_base | ((size_t)(l >> LOGT) << lo) | (l & ((1u << LOGT) - 1));
        return g_base | ((size_t)(l >> K) << (n - LOGT)) | (l & ((1u << K) - 1));
    };
    const int tshift = LAST ? 0 : LOGT;

    Fr x[8];
    bool first_round = true;
    // rounds of (K - 1) % 3 + 1, 3, 3, ... stages: the short round comes first, so that the final round of the last pass
    // is always the fixed 8-point network on the three lowest index bits (twiddles 1, w8, w8^2, w8^3: 5 products, not 12)
    for (int top = K; top > 0;) {
        const int nb = first_round ? (K - 1) % 3 + 1 : 3;
        const int rb = top - nb;                 // this round runs the stages on the tile bits [rb, top)
        const int wb = nb < 3 ? top - 3 : rb;    // a thread holds the three bits [wb, wb + 3): in a short round the stages use the upper nb of them
        const int pbit = wb + tshift;
        const unsigned lbase = ((t >> pbit) << (pbit + 3)) | (t & ((1u << pbit) - 1));
        unsigned rot = 0;                        // x[j] is the element whose three local bits are rotr3(j, rot)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            unsigned l = lbase | (rotr3(j, rot) << pbit);
            if (first_round) {
                size_t g = gidx(l);
                x[j] = ld_fr(in + g);
                if (p.first && p.pre) x[j] = x[j] * ldg_fr(p.pre + g);
                if (p.first && p.mul_b) x[j] = (x[j] * ldg_fr(p.mul_b + (size_t)blockIdx.y * p.batch_stride + g) - ldg_fr(p.sub_c + (size_t)blockIdx.y * p.batch_stride + g)) * ldg_fr(p.pre_const);
            } else {
                x[j] = ld_fr(sm + l);
            }
        }
        if (LAST && rb == 0 && nb == 3) {
            // stage b = 2: twiddle w8^j for the pair (j, j + 4); b = 1: 1, 1, w4, w4; b = 0: all 1.  Inverse transform:
            // w^-e = -w^(N/2 - e), the sign goes into the difference (as below).
            const size_t e8 = (size_t)1 << (n - 3);
            const Fr w1 = ldg_fr(p.tw + (p.inverse ? half_n - e8 : e8));
            const Fr w2 = ldg_fr(p.tw + (p.inverse ? half_n - 2 * e8 : 2 * e8));
            const Fr w3 = ldg_fr(p.tw + (p.inverse ? half_n - 3 * e8 : 3 * e8));
#pragma unroll
            for (int k = 0; k < 3; k++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const bool plain = k == 2 || (k == 1 && j < 2) || (k == 0 && j == 0);
                    Fr u = x[j] + x[j + 4];
                    if (plain) x[j + 4] = x[j] - x[j + 4];
                    else {
                        Fr d = p.inverse ? (x[j + 4] - x[j]) : (x[j] - x[j + 4]);
                        x[j + 4] = d * (k == 1 ? w2 : j == 1 ? w1 : j == 2 ? w2 : w3);
                    }
                    x[j] = u;
                }
                Fr y1 = x[1], y2 = x[2], y3 = x[3], y4 = x[4], y5 = x[5], y6 = x[6];
                x[2] = y1; x[4] = y2; x[6] = y3; x[1] = y4; x[3] = y5; x[5] = y6;
            }
        } else
#pragma unroll 1      // measured: fully unrolled stages (no register rotation) run 30 % slower (instruction cache), profiles/r02_ntt.md
        for (int k = 0; k < nb; k++) {
            const int b = lo + wb + 2 - k;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                size_t i0 = gidx(lbase | (rotr3(j, rot) << pbit));
                size_t e = (i0 & (((size_t)1 << b) - 1)) << (n - b - 1);
                Fr w = ldg_fr(p.tw + (p.inverse ? half_n - e : e));
                Fr u = x[j] + x[j + 4];
                Fr d = p.inverse ? (x[j + 4] - x[j]) : (x[j] - x[j + 4]);
                x[j] = u;
                x[j + 4] = d * w;
            }
            // rotate the 3-bit register index left by one: the next stage again pairs (j, j+4)
            Fr y1 = x[1], y2 = x[2], y3 = x[3], y4 = x[4], y5 = x[5], y6 = x[6];
            x[2] = y1; x[4] = y2; x[6] = y3; x[1] = y4; x[3] = y5; x[5] = y6;
            rot++;
        }
        const bool more = rb > 0;
        top = rb;
        if (more || LAST) {
#pragma unroll
            for (int j = 0; j < 8; j++) st_fr(sm + (lbase | (rotr3(j, rot) << pbit)), x[j]);
            __syncthreads();
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) st_fr(out + gidx(lbase | (rotr3(j, rot) << pbit)), x[j]);
        }
        first_round = false;
    }
    if (LAST) {
        // o = (kappa' << LOGT) | col'  ->  natural output index (kappa' << (n-K)) | (blk' << LOGT) | col'
        const int nblkbits = n - K - LOGT;
        const size_t blk_rev = nblkbits ? (size_t)(__brevll((unsigned long long)blk) >> (64 - nblkbits)) : 0;
        const unsigned nthreads = blockDim.x;
#pragma unroll 1
        for (unsigned o = t; o < (1u << L); o += nthreads) {
            unsigned cp = o & ((1u << LOGT) - 1), kp = o >> LOGT;
            unsigned col = LOGT ? (__brev(cp) >> (32 - LOGT)) : 0;
            unsigned kap = __brev(kp) >> (32 - K);
            size_t kidx = ((size_t)kp << (n - K)) | (blk_rev << LOGT) | cp;
            Fr v = ld_fr(sm + ((col << K) | kap));
            if (p.post) v = v * ldg_fr(p.post + kidx);
            else if (p.post_const) v = v * ldg_fr(p.post_const);
            Fr* dst = out;
            if (p.sc.n) {
                int j = 0;
                while (j + 1 < p.sc.n && kidx >= p.sc.hi[j]) j++;
                dst = p.sc.out[j];
            }
            st_fr(dst + kidx, v);
        }
    }
}

// ---- the same pass with the tile's butterflies working OUT OF shared memory -----------------------------------------------
// ntt_pass_kernel keeps 8 elements (64 registers) per thread and needs 118-128 registers: two CTAs = 15 warps per SM, where a
// chain of field products reaches ~0.6 of the integer pipe (profiles/r02_ntt.md).  Here a butterfly loads its two operands
// and its twiddle, and stores its two results, through one non-inlined routine (LDS.128 / STS.128 on the otherwise idle LSU,
// the scheme of msm_accumulate_g1_sm_kernel): ~80 registers, three CTAs = 24 warps per SM.  Same tile geometry, same rounds
// (a thread still owns 8 elements for up to three stages, so there is one barrier per round, not per stage), same fused
// first-load / last-store work, bit-identical results.  Shared-memory layout: the low and the high 16 bytes of element l
// in two planes (a warp's LDS.128 over consecutive l is conflict free), l swizzled by its own bits 3..5 so that the last
// rounds (a thread's 8 elements adjacent: stride-8 accesses across the warp) are conflict free as well.
static __device__ __forceinline__ uint32_t ntt_phys(uint32_t l) { return l ^ ((l >> 3) & 7u); }
static __device__ __forceinline__ Fr ntt_sm_ld(uint32_t base, uint32_t plane, uint32_t l) {
    const uint32_t a = base + ntt_phys(l) * 16u;
    Fr r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "r"(a) : "memory");
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "r"(a + plane) : "memory");
    return r;
}
static __device__ __forceinline__ void ntt_sm_st(uint32_t base, uint32_t plane, uint32_t l, const Fr& r) {
    const uint32_t a = base + ntt_phys(l) * 16u;
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]) : "memory");
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a + plane), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]) : "memory");
}
// one butterfly in place: (a, b) <- (a + b, (a - b) w)   [inverse: (b - a) w: the sign of w^-e = -w^(N/2 - e)];  w == nullptr: no product
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ void ntt_bfly_sm(uint32_t base, uint32_t plane, uint32_t la, uint32_t lb, const Fr* w, int inverse) {
    const Fr a = ntt_sm_ld(base, plane, la), b = ntt_sm_ld(base, plane, lb);
    ntt_sm_st(base, plane, la, a + b);
    if (!w) { ntt_sm_st(base, plane, lb, a - b); return; }
    const Fr d = inverse ? b - a : a - b;
    ntt_sm_st(base, plane, lb, d * ldg_fr(w));
}
#else
static inline void ntt_bfly_sm(uint32_t, uint32_t, uint32_t, uint32_t, const Fr*, int) {}
#endif

#ifndef ZA_NTT_SM_MINB
#define ZA_NTT_SM_MINB 3
#endif
template <bool LAST>
__global__ void __launch_bounds__(256, ZA_NTT_SM_MINB) ntt_pass_sm_kernel(NttPass p) {
    extern __shared__ uint4 smem_raw[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const int n = p.n, lo = p.lo, K = p.K, LOGT = p.LOGT, L = K + LOGT;
    const uint32_t plane = (16u << L);
    const unsigned t = threadIdx.x, nthreads = blockDim.x;
    const size_t blk = blockIdx.x;
    const Fr* in = p.in + (size_t)blockIdx.y * p.batch_stride;
    Fr* out = p.out + (size_t)blockIdx.y * p.batch_stride;
    const size_t half_n = (size_t)1 << (n - 1);
    size_t g_base;
    if (!LAST) {
        const int s = lo - LOGT;
        size_t blo = blk & (((size_t)1 << s) - 1), bhi = blk >> s;
        g_base = (bhi << (lo + K)) | (blo << LOGT);
    } else {
        g_base = blk << K;
    }
    auto gidx = [&](unsigned l) -> size_t {
        if (!LAST) return g_base | ((size_t)(l >> LOGT) << lo) | (l & ((1u << LOGT) - 1));
        return g_base | ((size_t)(l >> K) << (n - LOGT)) | (l & ((1u << K) - 1));
    };
    const int tshift = LAST ? 0 : LOGT;
    // the tile into shared memory (fused first-load work as in ntt_pass_kernel)
#pragma unroll 1
    for (unsigned l = t; l < (1u << L); l += nthreads) {
        const size_t g = gidx(l);
        Fr x = ld_fr(in + g);
        if (p.first && p.pre) x = x * ldg_fr(p.pre + g);
        if (p.first && p.mul_b) x = (x * ldg_fr(p.mul_b + (size_t)blockIdx.y * p.batch_stride + g) - ldg_fr(p.sub_c + (size_t)blockIdx.y * p.batch_stride + g)) * ldg_fr(p.pre_const);
        ntt_sm_st(base, plane, l, x);
    }
    __syncthreads();
    bool first_round = true;
    for (int top = K; top > 0;) {
        const int nb = first_round ? (K - 1) % 3 + 1 : 3;
        const int rb = top - nb;
        const int wb = nb < 3 ? top - 3 : rb;
        const int pbit = wb + tshift;
        const unsigned lbase = ((t >> pbit) << (pbit + 3)) | (t & ((1u << pbit) - 1));
        const bool fixed8 = LAST && rb == 0 && nb == 3;          // the constant 8-point network on the three lowest bits
        const size_t e8 = (size_t)1 << (n - 3);
#pragma unroll 1
        for (int k = 0; k < nb; k++) {
            const unsigned bitj = 4u >> k;                       // the stage pairs the elements whose 3-bit index differs in this bit
            const int b = lo + wb + 2 - k;
#pragma unroll 1
            for (unsigned q = 0; q < 4; q++) {
                // q enumerates the 3-bit indices j with bit `bitj` clear
                const unsigned j = bitj == 4u ? q : bitj == 2u ? ((q & 2u) << 1) | (q & 1u) : q << 1;
                const unsigned la = lbase | (j << pbit), lb = lbase | ((j | bitj) << pbit);
                const Fr* w;
                if (fixed8) {
                    // stage 0: w8^j for the pair (j, j + 4); stage 1: 1, w4 by bit 0 of the pair's position; stage 2: all 1
                    const unsigned ex = k == 0 ? j : k == 1 ? (j & 1u) * 2u : 0u;       // exponent in units of N / 8
                    w = ex ? p.tw + (p.inverse ? half_n - ex * e8 : ex * e8) : nullptr;
                } else {
                    const size_t i0 = gidx(la);
                    const size_t e = (i0 & (((size_t)1 << b) - 1)) << (n - b - 1);
                    w = p.tw + (p.inverse ? half_n - e : e);
                }
                ntt_bfly_sm(base, plane, la, lb, w, p.inverse);
            }
        }
        __syncthreads();
        top = rb;
        first_round = false;
    }
    if (!LAST) {
#pragma unroll 1
        for (unsigned l = t; l < (1u << L); l += nthreads) st_fr(out + gidx(l), ntt_sm_ld(base, plane, l));
    } else {
        const int nblkbits = n - K - LOGT;
        const size_t blk_rev = nblkbits ? (size_t)(__brevll((unsigned long long)blk) >> (64 - nblkbits)) : 0;
#pragma unroll 1
        for (unsigned o = t; o < (1u << L); o += nthreads) {
            unsigned cp = o & ((1u << LOGT) - 1), kp = o >> LOGT;
            unsigned col = LOGT ? (__brev(cp) >> (32 - LOGT)) : 0;
            unsigned kap = __brev(kp) >> (32 - K);
            size_t kidx = ((size_t)kp << (n - K)) | (blk_rev << LOGT) | cp;
            Fr v = ntt_sm_ld(base, plane, (col << K) | kap);
            if (p.post) v = v * ldg_fr(p.post + kidx);
            else if (p.post_const) v = v * ldg_fr(p.post_const);
            Fr* dst = out;
            if (p.sc.n) {
                int j = 0;
                while (j + 1 < p.sc.n && kidx >= p.sc.hi[j]) j++;
                dst = p.sc.out[j];
            }
            st_fr(dst + kidx, v);
        }
    }
}

// N <= 4: the definition, one thread per output
__global__ void ntt_tiny_kernel(const Fr* in, Fr* out, size_t batch_stride, const Fr* tw, const Fr* pre, const Fr* post,
                                const Fr* post_const, int n, int inverse) {
    const unsigned N = 1u << n, k = threadIdx.x;
    in += (size_t)blockIdx.y * batch_stride;
    out += (size_t)blockIdx.y * batch_stride;
    if (k >= N) return;
    Fr acc = Fr::zero();
    for (unsigned j = 0; j < N; j++) {
        Fr v = ld_fr(in + j);
        if (pre) v = v * ldg_fr(pre + j);
        unsigned e = (j * k) & (N - 1);          // omega^(jk); table holds e <= N/2
        if (inverse) e = (N - e) & (N - 1);
        Fr w = (e <= N / 2) ? ldg_fr(tw + e) : -ldg_fr(tw + (e - N / 2));
        acc = acc + v * w;
    }
    if (post) acc = acc * ldg_fr(post + k);
    else if (post_const) acc = acc * ldg_fr(post_const);
    __syncthreads();
    st_fr(out + k, acc);
}

// out[i] = first * base^i, i < count
__global__ void gen_powers_kernel(Fr* out, size_t count, Fr base, Fr first, int chunk) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t start = c * (size_t)chunk;
    if (start >= count) return;
    uint32_t e[8] = {(uint32_t)start, (uint32_t)(start >> 32), 0, 0, 0, 0, 0, 0};
    Fr cur = first * fp_pow<FrParams>(base, e);
    size_t end = start + chunk < count ? start + chunk : count;
    for (size_t i = start; i < end; i++) {
        st_fr(out + i, cur);
        cur = cur * base;
    }
}

// dir 0: canonical -> Montgomery ; 1: Montgomery -> canonical
__global__ void fr_convert_kernel(Fr* d, size_t n, int dir) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = ld_fr(d + i);
    st_fr(d + i, dir == 0 ? fp_to_mont<FrParams>(v) : fp_from_mont<FrParams>(v));
}

// a = (a*b - c) * zinv      (mul_assign, sub_assign, divide_by_z_on_coset fused)
__global__ void h_pointwise_kernel(Fr* a, const Fr* b, const Fr* c, const Fr* zinv, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr z = ldg_fr(zinv);
    st_fr(a + i, (ld_fr(a + i) * ldg_fr(b + i) - ldg_fr(c + i)) * z);
}
// unfused variants for checkpoint parity
__global__ void fr_mul_assign_kernel(Fr* a, const Fr* b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(a + i, ld_fr(a + i) * ldg_fr(b + i));
}
__global__ void fr_sub_assign_kernel(Fr* a, const Fr* b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(a + i, ld_fr(a + i) - ldg_fr(b + i));
}
__global__ void fr_scale_kernel(Fr* a, const Fr* k, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(a + i, ld_fr(a + i) * ldg_fr(k));
}

// ------------------------------------------------------------------------------------------ host
static Fr host_fr_from_u64(uint64_t x) { return fp_from_u64<FrParams>(x); }
static Fr host_fr_pow_u64(const Fr& a, uint64_t e) {
    uint32_t w[8] = {(uint32_t)e, (uint32_t)(e >> 32), 0, 0, 0, 0, 0, 0};
    return fp_pow<FrParams>(a, w);
}
// 2^28-th root of unity 7^((r-1)/2^28) (canonical limbs; SURVEY §8a, machine-checked there)
static Fr host_root_of_unity() {
    Fr c;
    const uint32_t v[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du, 0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
    for (int i = 0; i < 8; i++) c.v[i] = v[i];
    return fp_to_mont<FrParams>(c);
}
Fr host_domain_omega(int log_n) {
    Fr w = host_root_of_unity();
    for (int i = log_n; i < 28; i++) w = sqr(w);
    return w;
}

static void launch_powers(Ctx* ctx, Fr* out, size_t count, const Fr& base, const Fr& first) {
    const int chunk = 64;
    size_t threads = (count + chunk - 1) / chunk;
    unsigned blocks = (unsigned)((threads + 127) / 128);
    gen_powers_kernel<<<blocks, 128, 0, ctx->stream>>>(out, count, base, first, chunk);
    ctx->launches++;
    ZA_CUDA(cudaGetLastError());
}

void launch_powers_public(Ctx* ctx, Fr* out, size_t count, const Fr& base, const Fr& first) { launch_powers(ctx, out, count, base, first); }

NttDomain* get_domain(Ctx* ctx, int log_n, bool need_coset) {
    NttDomain* d;
    auto it = ctx->domains.find(log_n);
    if (it == ctx->domains.end()) {
        d = new NttDomain();
        d->log_n = log_n;
        ctx->domains[log_n] = d;
        size_t N = (size_t)1 << log_n;
        size_t ntw = N / 2 + 1;
        d->tw.alloc(ntw * sizeof(Fr));
        launch_powers(ctx, d->tw.as<Fr>(), ntw, host_domain_omega(log_n), Fr::one());
        Fr g = host_fr_from_u64(7);
        Fr minv = inv(host_fr_from_u64((uint64_t)N));
        Fr zinv = inv(host_fr_pow_u64(g, (uint64_t)N) - Fr::one());
        Fr consts[2] = {minv, zinv};
        d->consts.alloc(sizeof(consts));
        ZA_CUDA(cudaMemcpyAsync(d->consts.p, consts, sizeof(consts), cudaMemcpyHostToDevice, ctx->stream));
        ZA_CUDA(cudaStreamSynchronize(ctx->stream));  // consts[] is a stack buffer
    } else {
        d = it->second;
    }
    if (need_coset && !d->have_coset) {
        size_t N = (size_t)1 << log_n;
        Fr g = host_fr_from_u64(7), ginv = inv(g);
        Fr minv = inv(host_fr_from_u64((uint64_t)N));
        d->pow_g.alloc(N * sizeof(Fr));
        d->pow_g_minv.alloc(N * sizeof(Fr));
        d->pow_ginv_minv.alloc(N * sizeof(Fr));
        launch_powers(ctx, d->pow_g.as<Fr>(), N, g, Fr::one());
        launch_powers(ctx, d->pow_g_minv.as<Fr>(), N, g, minv);
        launch_powers(ctx, d->pow_ginv_minv.as<Fr>(), N, ginv, minv);
        d->have_coset = true;
    }
    return d;
}

static const int NTT_L = 11;  // log2 elements per CTA tile (2^11 x 32 B = 64 KiB shared memory)

// Split log_n into pass sizes.  Every pass needs 3 <= K <= NTT_L (a thread owns 8 elements).  Default: the fewest
// passes, sizes as even as possible (2^20: 10 + 10; 2^24: 8 + 8 + 8); ZA_NTT_MAXK caps K (9 = the three-pass plan of
// round 1 at 2^20).
static std::vector<int> plan_passes(int n) {
    std::vector<int> ks;
    if (n <= NTT_L) { ks.push_back(n); return ks; }
    int maxk = 10;
    if (const char* e = getenv("ZA_NTT_MAXK")) { int v = atoi(e); if (v >= 6 && v <= NTT_L) maxk = v; }
    int P = (n + maxk - 1) / maxk;
    int base = n / P, extra = n % P;
    for (int i = 0; i < P; i++) ks.push_back(base + (i < extra ? 1 : 0));
    return ks;
}

// what is fused into the first load / last store of a transform
struct NttFuse {
    const Fr* pre = nullptr;         // x *= pre[i]                        (coset shift)
    const Fr* post = nullptr;        // y *= post[k]                       (1/m, coset un-shift)
    const Fr* post_const = nullptr;  // y *= *post_const
    const Fr* mul_b = nullptr;       // x = (x * mul_b[i] - sub_c[i]) * *pre_const
    const Fr* sub_c = nullptr;
    const Fr* pre_const = nullptr;
    Fr* out = nullptr;               // the result goes here instead of back into `buf` (batch 1)
    const HScatter* scatter = nullptr; // ... or to several destinations by output index (batch 1, two or more passes)
};

// One transform (batch vectors, `stride` elements apart), data in `buf` (Montgomery), result back in `buf`.
// scratch must hold batch * stride elements.
static void ntt_run(Ctx* ctx, Fr* buf, Fr* scratch, int n, int batch, size_t stride, bool inverse, const Fr* tw, const NttFuse& f) {
    const size_t N = (size_t)1 << n;
    ProfScope prof(ctx, PROF_NTT, (double)batch * (double)N);
    if (!ctx->ntt_attr_set) {          // per device: a context belongs to one device, a process may hold several contexts
        ZA_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << NTT_L) * 32));
        ZA_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << NTT_L) * 32));
        ZA_CUDA(cudaFuncSetAttribute(ntt_pass_sm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << NTT_L) * 32));
        ZA_CUDA(cudaFuncSetAttribute(ntt_pass_sm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << NTT_L) * 32));
        ctx->ntt_attr_set = true;
    }
    if (n < 3) {
        if (f.mul_b) throw ZaError(ZA_ERR_INVALID, "internal: fused pointwise load needs a domain of 8 or more elements");
        ntt_tiny_kernel<<<dim3(1, batch), 4, 0, ctx->stream>>>(buf, scratch, stride, tw, f.pre, f.post, f.post_const, n, inverse ? 1 : 0);
        ctx->launches++;
        ZA_CUDA(cudaGetLastError());
        for (int v = 0; v < batch; v++)
            ZA_CUDA(cudaMemcpyAsync((f.out ? f.out : buf) + (size_t)v * stride, scratch + (size_t)v * stride, N * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
        return;
    }
    std::vector<int> ks = plan_passes(n);
    const int P = (int)ks.size();
    int lo = n;
    for (int pi = 0; pi < P; pi++) {
        const int K = ks[pi];
        lo -= K;
        const bool last = pi == P - 1;
        NttPass p;
        p.batch_stride = stride;
        p.tw = tw; p.pre = f.pre; p.post = f.post; p.post_const = f.post_const;
        p.mul_b = f.mul_b; p.sub_c = f.sub_c; p.pre_const = f.pre_const;
        p.n = n; p.lo = lo; p.K = K; p.inverse = inverse ? 1 : 0; p.first = pi == 0;
        p.sc.n = 0;
        if (pi == P - 1 && f.scatter && f.scatter->n) {
            if (P < 2 || batch != 1) throw ZaError(ZA_ERR_INVALID, "internal: scattered NTT output needs one vector and two or more passes");
            p.sc = *f.scatter;
        }
        int logt = NTT_L - K;
        if (last) { if (logt > n - K) logt = n - K; } else { if (logt > lo) logt = lo; }
        p.LOGT = logt;
        const int L = K + logt;
        // data flow: buf -> scratch (pass 0), scratch in place (middle), scratch -> buf (last); P == 1: buf -> scratch + copy
        if (P == 1) { p.in = buf; p.out = scratch; }
        else if (pi == 0) { p.in = buf; p.out = scratch; }
        else if (last) { p.in = scratch; p.out = f.out ? f.out : buf; }
        else { p.in = scratch; p.out = scratch; }
        dim3 grid((unsigned)(N >> L), batch);
        unsigned threads = 1u << (L - 3);
        size_t smem = ((size_t)1 << L) * sizeof(Fr);
        static const bool use_sm = getenv("ZA_NTT_SM") && atoi(getenv("ZA_NTT_SM")) != 0;
        if (use_sm) {
            if (last) ntt_pass_sm_kernel<true><<<grid, threads, smem, ctx->stream>>>(p);
            else ntt_pass_sm_kernel<false><<<grid, threads, smem, ctx->stream>>>(p);
        } else if (last) ntt_pass_kernel<true><<<grid, threads, smem, ctx->stream>>>(p);
        else ntt_pass_kernel<false><<<grid, threads, smem, ctx->stream>>>(p);
        ctx->launches++;
        ZA_CUDA(cudaGetLastError());
    }
    if (P == 1)
        for (int v = 0; v < batch; v++)
            ZA_CUDA(cudaMemcpyAsync((f.out ? f.out : buf) + (size_t)v * stride, scratch + (size_t)v * stride, N * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
}

static inline unsigned nblk(size_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

void fr_convert(Ctx* ctx, Fr* d, size_t n, int dir) {
    if (!n) return;
    fr_convert_kernel<<<nblk(n, 256), 256, 0, ctx->stream>>>(d, n, dir);
    ctx->launches++;
    ZA_CUDA(cudaGetLastError());
}

// EvaluationDomain::{fft, ifft, coset_fft, icoset_fft} on device-resident Montgomery vectors
void ntt_mode(Ctx* ctx, Fr* buf, int log_n, int mode, int batch) {
    if (log_n >= 28) throw ZaError(ZA_ERR_POLY_DEGREE_TOO_LARGE, "domain of 2^28 or more elements (Fr two-adicity is 28)");
    if (log_n == 0) {
        // N = 1: fft/ifft are the identity, coset shifts multiply by g^0 = 1, 1/m = 1
        return;
    }
    NttDomain* d = get_domain(ctx, log_n, mode >= 2);
    size_t N = (size_t)1 << log_n;
    DevBuf& scratch = ctx->scratch[0];
    scratch.ensure((size_t)batch * N * sizeof(Fr));
    const Fr* tw = d->tw.as<Fr>();
    NttFuse f;
    switch (mode) {
    case ZA_NTT_FFT: break;
    case ZA_NTT_IFFT: f.post_const = d->consts.as<Fr>(); break;
    case ZA_NTT_COSET_FFT: f.pre = d->pow_g.as<Fr>(); break;
    case ZA_NTT_ICOSET_FFT: f.post = d->pow_ginv_minv.as<Fr>(); break;
    default: throw ZaError(ZA_ERR_INVALID, "unknown NTT mode");
    }
    ntt_run(ctx, buf, scratch.as<Fr>(), log_n, batch, N, mode == ZA_NTT_IFFT || mode == ZA_NTT_ICOSET_FFT, tw, f);
}

// Fused H pipeline on device (create_proof step 4).  Seven transforms in three groups: a, b and c go through ifft
// and coset_fft TOGETHER (one launch per pass, blockIdx.y = vector) when they lie `m` elements apart; the scalings ride
// on the transforms' last stores (ifft -> coset shift: one multiplication by m^-1 g^i), mul_assign / sub_assign /
// divide_by_z_on_coset on the first load of the last transform, and its last store multiplies by the CANONICAL
// m^-1 g^-i, which also takes the result out of Montgomery form.  Result: canonical h coefficients in a[0..m-1).
void h_poly_device(Ctx* ctx, Fr* a, Fr* b, Fr* c, int log_m, Fr* h_out) {
    if (log_m >= 28) throw ZaError(ZA_ERR_POLY_DEGREE_TOO_LARGE, "domain of 2^28 or more elements");
    size_t m = (size_t)1 << log_m;
    if (log_m == 0) {
        // m = 1: ifft/coset_fft are identities; h has m-1 = 0 coefficients
        return;
    }
    if (h_out == a) h_out = nullptr;
    NttDomain* d = get_domain(ctx, log_m, true);
    if (!d->pow_ginv_minv_canon.p) {
        // the same table out of Montgomery form: montmul(x R, t) = x t
        d->pow_ginv_minv_canon.alloc(m * sizeof(Fr));
        ZA_CUDA(cudaMemcpyAsync(d->pow_ginv_minv_canon.p, d->pow_ginv_minv.p, m * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
        fr_convert(ctx, d->pow_ginv_minv_canon.as<Fr>(), m, 1);
    }
    const bool together = b == a + m && c == b + m;
    DevBuf& scratch = ctx->scratch[0];
    scratch.ensure((together ? 3 : 1) * m * sizeof(Fr));
    const Fr* tw = d->tw.as<Fr>();
    NttFuse inv_shift; inv_shift.post = d->pow_g_minv.as<Fr>();       // ifft then coset shift: one post-scale by m^-1 g^i
    NttFuse plain;
    if (together) {
        ntt_run(ctx, a, scratch.as<Fr>(), log_m, 3, m, true, tw, inv_shift);
        ntt_run(ctx, a, scratch.as<Fr>(), log_m, 3, m, false, tw, plain);
    } else {
        Fr* vecs[3] = {a, b, c};
        for (int v = 0; v < 3; v++) {
            ntt_run(ctx, vecs[v], scratch.as<Fr>(), log_m, 1, m, true, tw, inv_shift);
            ntt_run(ctx, vecs[v], scratch.as<Fr>(), log_m, 1, m, false, tw, plain);
        }
    }
    if (log_m >= 3) {
        NttFuse last;
        last.mul_b = b; last.sub_c = c; last.pre_const = d->consts.as<Fr>() + 1;
        last.post = d->pow_ginv_minv_canon.as<Fr>();
        last.out = h_out;
        if (ctx->h_scatter.n) {
            if (log_m <= NTT_L) throw ZaError(ZA_ERR_INVALID, "internal: scattered h output needs a domain of more than one NTT tile");
            last.scatter = &ctx->h_scatter;
        }
        ntt_run(ctx, a, scratch.as<Fr>(), log_m, 1, m, true, tw, last);
    } else {
        {
            ProfScope prof(ctx, PROF_POINTWISE, (double)m);
            h_pointwise_kernel<<<nblk(m, 256), 256, 0, ctx->stream>>>(a, b, c, d->consts.as<Fr>() + 1, m);
            ctx->launches++;
            ZA_CUDA(cudaGetLastError());
        }
        NttFuse last; last.post = d->pow_ginv_minv_canon.as<Fr>(); last.out = h_out;
        ntt_run(ctx, a, scratch.as<Fr>(), log_m, 1, m, true, tw, last);
    }
}

// Unfused H pipeline that materialises every vector bellman materialises (parity checkpoints).
// ck (host, 8*m*32 bytes, canonical) may be null.
void h_poly_checkpointed(Ctx* ctx, Fr* a, Fr* b, Fr* c, int log_m, uint8_t* ck) {
    size_t m = (size_t)1 << log_m;
    DevBuf tmp(m * sizeof(Fr));
    auto snap = [&](int k, Fr* v) {
        if (!ck) return;
        ZA_CUDA(cudaMemcpyAsync(tmp.p, v, m * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
        fr_convert(ctx, tmp.as<Fr>(), m, 1);
        ZA_CUDA(cudaMemcpyAsync(ck + (size_t)k * m * 32, tmp.p, m * 32, cudaMemcpyDeviceToHost, ctx->stream));
        ZA_CUDA(cudaStreamSynchronize(ctx->stream));
    };
    Fr* vecs[3] = {a, b, c};
    for (int v = 0; v < 3; v++) {
        ntt_mode(ctx, vecs[v], log_m, ZA_NTT_IFFT, 1); snap(2 * v, vecs[v]);
        ntt_mode(ctx, vecs[v], log_m, ZA_NTT_COSET_FFT, 1); snap(2 * v + 1, vecs[v]);
    }
    if (log_m > 0) {
        NttDomain* d = get_domain(ctx, log_m, true);
        fr_mul_assign_kernel<<<nblk(m, 256), 256, 0, ctx->stream>>>(a, b, m);
        fr_sub_assign_kernel<<<nblk(m, 256), 256, 0, ctx->stream>>>(a, c, m);
        fr_scale_kernel<<<nblk(m, 256), 256, 0, ctx->stream>>>(a, d->consts.as<Fr>() + 1, m);
        ctx->launches += 3;
        ZA_CUDA(cudaGetLastError());
    } else {
        // m = 1: Z on the coset is g - 1 = 6
        Fr zinv = inv(host_fr_from_u64(6));
        DevBuf k(sizeof(Fr));
        ZA_CUDA(cudaMemcpyAsync(k.p, &zinv, sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
        fr_mul_assign_kernel<<<1, 32, 0, ctx->stream>>>(a, b, m);
        fr_sub_assign_kernel<<<1, 32, 0, ctx->stream>>>(a, c, m);
        fr_scale_kernel<<<1, 32, 0, ctx->stream>>>(a, k.as<Fr>(), m);
        ctx->launches += 3;
        ZA_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    snap(6, a);
    ntt_mode(ctx, a, log_m, ZA_NTT_ICOSET_FFT, 1); snap(7, a);
    fr_convert(ctx, a, m, 1);
}

}  // namespace za
