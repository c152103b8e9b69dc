// Pippenger multi-scalar multiplication over BN254 G1 / G2.
//
// Replaces (un-vendored) bellman_ce multiexp.rs `multiexp` / `multiexp_inner`, which create_proof
// calls eight times per proof (entered from
// /root/reference/prover/src/groth16/prover.rs:173).  bellman: unsigned c = ceil(ln n) windows,
// one CPU task per window, Jacobian buckets.  Here:
//
//   K3  msm_digits_*      signed-digit recoding of every scalar (W = ceil(255/c) windows,
//                         2^(c-1) buckets per window, negation is free)
//   K4  counting sort     histogram (global atomics) -> exclusive scan -> scatter of
//                         (point index | sign) into bucket order
//   K5  msm_accumulate    the sorted entry stream is cut into equal chunks, one per thread; a thread
//                         walks its chunk with XYZZ mixed additions (8M+2S); buckets that straddle
//                         chunk boundaries leave partial sums that a warp-cooperative fix-up kernel
//                         folds together — load balance is independent of the scalar distribution
//                         (witness vectors are dominated by 0 and 1)
//   K6  msm_bucket_reduce running-sum over segments of buckets + weight fix-up, then a warp
//                         reduction per window; the W window sums go to the host, which does the
//                         c doublings per window (a few hundred field operations)
//
// All results are exact group elements; equality with the reference is checked after affine
// normalisation (representation independent).
#include "common.cuh"
#include "api_internal.cuh"
#include <string.h>
#include <stdlib.h>

#ifndef ZA_ACC_G1_BLOCKS
#define ZA_ACC_G1_BLOCKS 4
#endif
#ifndef ZA_ACC_G2_BLOCKS
#define ZA_ACC_G2_BLOCKS 2
#endif
#ifndef ZA_PAIR_G2_BLOCKS
#define ZA_PAIR_G2_BLOCKS 2
#endif
// register cap of the kernels over Fq2 (accumulation and pair rounds): 255 = two CTAs of 128 threads per SM and nothing
// else; 192 leaves the registers of one G1 accumulation CTA next to them
#ifndef ZA_G2_MAXNREG
#define ZA_G2_MAXNREG 255
#endif

namespace za {

// ------------------------------------------------------------------------------ device helpers
template <class T>
static __device__ __forceinline__ T ld_vec(const T* p) {
    static_assert(sizeof(T) % 16 == 0, "16-byte multiple");
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) d[i] = s[i];
    return r;
}
template <class T>
static __device__ __forceinline__ T ldg_vec(const T* p) {
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) d[i] = __ldg(s + i);
    return r;
}
template <class T>
static __device__ __forceinline__ void st_vec(T* p, const T& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) d[i] = s[i];
}

// signed digit of window w given the running carry; updates the carry.
// digit in [-(2^(c-1) - 1), 2^(c-1)]
static __device__ __forceinline__ int next_digit(const uint32_t* s, int w, int c, uint32_t& carry) {
    const int bit = w * c;
    const int word = bit >> 5, sh = bit & 31;
    uint64_t v = 0;
    if (word < 8) v = s[word];
    if (word + 1 < 8) v |= (uint64_t)s[word + 1] << 32;
    uint32_t raw = (uint32_t)((v >> sh) & ((1u << c) - 1)) + carry;
    if (raw > (1u << (c - 1))) { carry = 1; return (int)raw - (1 << c); }
    carry = 0;
    return (int)raw;
}

// Warp-aggregated atomics: lanes of a warp that hit the same bucket (witness vectors are full of 0/1 scalars, and
// the top window of a 254-bit scalar has few significant bits) are combined into ONE atomic on that counter —
// same-address atomics serialise in L2 and used to dominate the sort for such inputs.
static __device__ __forceinline__ void warp_count_add(uint32_t* counts, uint32_t key) {
    const unsigned peers = __match_any_sync(__activemask(), key);
    const unsigned lane = threadIdx.x & 31;
    if ((unsigned)(__ffs(peers) - 1) == lane) atomicAdd(&counts[key], (uint32_t)__popc(peers));
}
static __device__ __forceinline__ uint32_t warp_cursor_take(uint32_t* cursors, uint32_t key) {
    const unsigned peers = __match_any_sync(__activemask(), key);
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((unsigned)leader == lane) base = atomicAdd(&cursors[key], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + (uint32_t)__popc(peers & ((1u << lane) - 1));
}

// K3 + K4a: per-bucket entry counts
// `single` != 0: fixed-base table mode — one bucket space for all windows, entry = (i * W + w) | sign.
// `key_base`: first key of this part's bucket space when several queries are sorted into one key range (K4').
__global__ void msm_digits_hist_kernel(const uint32_t* __restrict__ scalars, size_t n, int c, int W, uint32_t B, uint32_t* counts, int single, uint32_t key_base) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    const uint4* p = reinterpret_cast<const uint4*>(scalars + 8 * i);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    if ((s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) == 0) return;
    uint32_t carry = 0;
    for (int w = 0; w < W; w++) {
        int d = next_digit(s, w, c, carry);
        if (d != 0) warp_count_add(counts, key_base + (single ? 0u : (uint32_t)w * B) + (uint32_t)(d < 0 ? -d : d) - 1);
    }
}

// K4c: scatter (point index | sign << 31) into bucket order.  Order inside a bucket is arbitrary;
// the bucket sum is not (group addition is commutative and exact).
__global__ void msm_digits_scatter_kernel(const uint32_t* __restrict__ scalars, size_t n, int c, int W, uint32_t B, uint32_t* cursors,
                                          uint32_t* entries, int single, uint32_t key_base) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    const uint4* p = reinterpret_cast<const uint4*>(scalars + 8 * i);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    if ((s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) == 0) return;
    uint32_t carry = 0;
    for (int w = 0; w < W; w++) {
        int d = next_digit(s, w, c, carry);
        if (d != 0) {
            uint32_t pos = warp_cursor_take(cursors, key_base + (single ? 0u : (uint32_t)w * B) + (uint32_t)(d < 0 ? -d : d) - 1);
            entries[pos] = (single ? (uint32_t)i * (uint32_t)W + (uint32_t)w : (uint32_t)i) | (d < 0 ? 0x80000000u : 0u);
        }
    }
}

// K4b: exclusive scan of the per-bucket counts in three small kernels (4096 counters per CTA):
// per-CTA totals -> scan of the totals -> per-CTA rescan with the CTA's base offset.
#define SCAN_PER_CTA 4096
static __device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t* smem /*32*/, uint32_t& total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (unsigned)d) x += y; }
    if (lane == 31) smem[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = smem[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= (unsigned)d) w += y; }
        smem[lane] = w;
    }
    __syncthreads();
    total = smem[31];
    uint32_t base = warp ? smem[warp - 1] : 0;
    __syncthreads();
    return base + x - v;
}
// `halve`: the counters are not read from `counts` but derived from the previous pair round's offsets:
// a bucket holding k points holds ceil(k / 2) after one round of pairwise additions.
static __device__ __forceinline__ uint32_t scan_count_at(const uint32_t* counts, uint32_t i, int halve) {
    return halve ? (counts[i + 1] - counts[i] + 1) >> 1 : counts[i];
}
__global__ void __launch_bounds__(1024) msm_scan_totals_kernel(const uint32_t* counts, uint32_t count, uint32_t* cta_totals, int halve) {
    __shared__ uint32_t sm[32];
    const uint32_t base = blockIdx.x * SCAN_PER_CTA + threadIdx.x * 4;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) if (base + k < count) s += scan_count_at(counts, base + k, halve);
    uint32_t total;
    block_excl_scan_1024(s, sm, total);
    if (threadIdx.x == 0) cta_totals[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) msm_scan_ctas_kernel(uint32_t* cta_totals, uint32_t nctas, uint32_t* grand_total) {
    __shared__ uint32_t sm[32];
    uint32_t run = 0;
    for (uint32_t b0 = 0; b0 < nctas; b0 += 1024) {
        uint32_t i = b0 + threadIdx.x;
        uint32_t v = i < nctas ? cta_totals[i] : 0;
        uint32_t total;
        uint32_t ex = block_excl_scan_1024(v, sm, total);
        if (i < nctas) cta_totals[i] = run + ex;
        run += total;
    }
    if (threadIdx.x == 0) *grand_total = run;
}
__global__ void __launch_bounds__(1024) msm_scan_apply_kernel(const uint32_t* counts, uint32_t count, const uint32_t* cta_offsets, uint32_t* offsets,
                                                              uint32_t* cursors, int halve) {
    __shared__ uint32_t sm[32];
    const uint32_t base = blockIdx.x * SCAN_PER_CTA + threadIdx.x * 4;
    uint32_t v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = base + k < count ? scan_count_at(counts, base + k, halve) : 0; s += v[k]; }
    uint32_t total;
    uint32_t run = cta_offsets[blockIdx.x] + block_excl_scan_1024(s, sm, total);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < count) { offsets[base + k] = run; if (cursors) cursors[base + k] = run; }
        run += v[k];
    }
}

// K5: chunked segmented accumulation.  Thread `chunk` owns entries [chunk*Lc, (chunk+1)*Lc).
// DIRECT: the stream is the output of the pair rounds (K5a) — entry `pos` is the point bases[pos] itself, no sign,
// and may be the point at infinity (P + (-P) in a round).
template <class F, bool DIRECT>
__global__ void __launch_bounds__(128) __maxnreg__(sizeof(F) == sizeof(Fq) ? 512 / ZA_ACC_G1_BLOCKS : ZA_G2_MAXNREG) msm_accumulate_kernel(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ entries,
                                                             const uint32_t* __restrict__ offsets, uint32_t nkeys, uint32_t Lc,
                                                             XYZZ<F>* bucket_sums, XYZZ<F>* part_head, XYZZ<F>* part_tail,
                                                             uint32_t* tail_owner_key) {
    const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t E = offsets[nkeys];
    const uint64_t start64 = (uint64_t)chunk * Lc;
    if (start64 >= E) return;
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = (uint64_t)start + Lc < E ? start + Lc : E;
    // last key with offsets[key] <= start  (non-empty by construction)
    uint32_t lo = 0, hi = nkeys;
    while (lo + 1 < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= start) lo = mid; else hi = mid;
    }
    uint32_t key = lo;
    bool head_open = offsets[key] < start;
    uint32_t bend = offsets[key + 1];
    XYZZ<F> acc = XYZZ<F>::inf();
    uint32_t e = DIRECT ? start : entries[start];
    Affine<F> P = ldg_vec(bases + (e & 0x7fffffffu));
    // One flat loop: every lane performs one mixed addition per iteration (convergent); the bucket
    // bookkeeping at run boundaries is the only divergent part and is a few instructions plus one store.
    for (uint32_t pos = start; pos < end; pos++) {
        const uint32_t e_cur = e;
        const Affine<F> P_cur = P;
        if (pos + 1 < end) {                       // prefetch the next point while this one is added
            e = DIRECT ? pos + 1 : entries[pos + 1];
            P = ldg_vec(bases + (e & 0x7fffffffu));
        }
        if (!DIRECT || !P_cur.is_inf()) xyzz_madd_hot(acc, P_cur.x, P_cur.y, !DIRECT && (e_cur >> 31) != 0);
        const uint32_t nxt = pos + 1;
        if (nxt == bend || nxt == end) {
            const bool closes = nxt == bend;
            if (!head_open && closes) st_vec(bucket_sums + key, acc);
            else if (head_open) st_vec(part_head + chunk, acc);
            else { st_vec(part_tail + chunk, acc); tail_owner_key[chunk] = key; }
            head_open = false;
            acc = XYZZ<F>::inf();
            if (nxt < end) {
                do { key++; } while (offsets[key + 1] <= nxt);
                bend = offsets[key + 1];
            }
        }
    }
}


// K5 (G1): the same chunked accumulation with the thread's working set — the XYZZ accumulator, the two point
// buffers and five temporaries — kept in SHARED memory, and a field product that takes three shared-memory
// addresses.  Measured reason (profiles/r01_session3.md): with the product as a by-value call, every call costs
// ~32 IMAD.MOV register moves (16 argument words in, 8 result words out, plus rotation), ptxas issues those moves on
// the integer-multiply pipe (fmaheavy, 2 cycles each), and that pipe is the kernel's bound: 82 % busy for 70 % of
// algorithmic IMAD.  Operands that live in shared memory are loaded straight into the registers the product reads
// (LDS.128 on the otherwise idle LSU pipe) and the result is stored from the registers it is produced in.
// Variable v of thread t: chunk c (16 bytes) at ((2 v + c) * 128 + t) * 16 — a warp's access is 512 contiguous bytes.
// The next point travels into the other point buffer with cp.async while the current one is added.
#define ACC_SM_NT 128
enum { AV_X = 0, AV_Y, AV_ZZ, AV_ZZZ, AV_PX0, AV_PY0, AV_PX1, AV_PY1, AV_P, AV_R, AV_PP, AV_PPP, AV_Q, AV_COUNT };
#define ACC_SM_BYTES (AV_COUNT * 32 * ACC_SM_NT)
static __device__ __forceinline__ Fq sm_ld(uint32_t a) {
    Fq r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "r"(a) : "memory");
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "r"(a + ACC_SM_NT * 16) : "memory");
    return r;
}
static __device__ __forceinline__ void sm_st(uint32_t a, const Fq& r) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]) : "memory");
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a + ACC_SM_NT * 16), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]) : "memory");
}
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ void fq_mul_sm(uint32_t d, uint32_t a, uint32_t b) { sm_st(d, fp_mul<FqParams>(sm_ld(a), sm_ld(b))); }
static __device__ __noinline__ void fq_sqr_sm(uint32_t d, uint32_t a) { sm_st(d, fp_sqr<FqParams>(sm_ld(a))); }
#else
static inline void fq_mul_sm(uint32_t, uint32_t, uint32_t) {}
static inline void fq_sqr_sm(uint32_t, uint32_t) {}
#endif
static __device__ __forceinline__ void cp_async16_sm(uint32_t smem_addr, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
static __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
static __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// K4' — several queries in ONE sort / accumulation / reduction: part p owns the keys [p << nb, (p + 1) << nb) (its own
// bucket space) and its entries index its own fixed-base table.  The witness multiexps of a proof (B in G1, L, A) run
// this way: one digit sort instead of three, one accumulation launch, one reduction chain.
struct AccTabs {
    const Affine<Fq>* tab[4];
    uint32_t nb;          // log2 of the keys per part (31 when there is a single part)
    uint32_t nparts;
};
template <bool DIRECT>
__global__ void __launch_bounds__(ACC_SM_NT, 4) msm_accumulate_g1_sm_kernel(const AccTabs tabs, const uint32_t* __restrict__ entries,
                                                                             const uint32_t* __restrict__ offsets, uint32_t nkeys, uint32_t Lc,
                                                                             XYZZ<Fq>* bucket_sums, XYZZ<Fq>* part_head, XYZZ<Fq>* part_tail,
                                                                             uint32_t* tail_owner_key) {
    extern __shared__ uint4 acc_sm[];
    const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t E = offsets[nkeys];
    const uint64_t start64 = (uint64_t)chunk * Lc;
    if (start64 >= E) return;
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = (uint64_t)start + Lc < E ? start + Lc : E;
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(acc_sm) + threadIdx.x * 16;
    auto var = [&](int v) { return sm0 + (uint32_t)v * (2 * ACC_SM_NT * 16); };
    uint32_t lo = 0, hi = nkeys;                    // last key with offsets[key] <= start (non-empty by construction)
    while (lo + 1 < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= start) lo = mid; else hi = mid;
    }
    uint32_t key = lo;
    bool head_open = offsets[key] < start;
    uint32_t bend = offsets[key + 1];
    bool acc_inf = true;
    const uint32_t nb = tabs.nb;
    uint32_t part = tabs.nparts > 1 ? key >> nb : 0u;
    const Affine<Fq>* bases = tabs.tab[part];
    uint32_t part_end = tabs.nparts > 1 ? offsets[(part + 1) << nb] : 0xffffffffu;      // first entry of the next part
    auto fetch = [&](uint32_t e, uint32_t buf, const Affine<Fq>* from) {     // point e -> point buffer `buf` (x: 2 chunks, y: 2 chunks)
        const uint8_t* src = reinterpret_cast<const uint8_t*>(from + (e & 0x7fffffffu));
        const uint32_t px = var(AV_PX0 + 2 * (int)buf), py = var(AV_PY0 + 2 * (int)buf);
        cp_async16_sm(px, src); cp_async16_sm(px + ACC_SM_NT * 16, src + 16);
        cp_async16_sm(py, src + 32); cp_async16_sm(py + ACC_SM_NT * 16, src + 48);
    };
    uint32_t e = DIRECT ? start : entries[start];
    fetch(e, start & 1u, bases);
    cp_async_commit();
    for (uint32_t pos = start; pos < end; pos++) {
        const uint32_t e_cur = e;
        const uint32_t buf = pos & 1u;
        if (pos + 1 < end) {                         // the next point travels while this one is added
            e = DIRECT ? pos + 1 : entries[pos + 1];
            const Affine<Fq>* from = bases;
            if (pos + 1 >= part_end) {               // it belongs to a later part (rare: twice per launch)
                uint32_t p2 = part + 1;
                while (p2 + 1 < tabs.nparts && offsets[(p2 + 1) << nb] <= pos + 1) p2++;
                from = tabs.tab[p2];
            }
            fetch(e, buf ^ 1u, from);
        }
        cp_async_commit();
        cp_async_wait<1>();
        const uint32_t vx = var(AV_PX0 + 2 * (int)buf), vy = var(AV_PY0 + 2 * (int)buf);
        const bool negate = !DIRECT && (e_cur >> 31) != 0;
        bool skip = false;
        if (DIRECT) { const Fq x = sm_ld(vx), y = sm_ld(vy); skip = x.is_zero() && y.is_zero(); }
        if (!skip) {
            if (acc_inf) {
                Fq y = sm_ld(vy);
                if (negate) y = -y;
                sm_st(var(AV_X), sm_ld(vx)); sm_st(var(AV_Y), y);
                sm_st(var(AV_ZZ), Fq::one()); sm_st(var(AV_ZZZ), Fq::one());
                acc_inf = false;
            } else {
                fq_mul_sm(var(AV_P), vx, var(AV_ZZ));                 // U2
                fq_mul_sm(var(AV_R), vy, var(AV_ZZZ));                // S2
                const Fq P = sm_ld(var(AV_P)) - sm_ld(var(AV_X));
                Fq S2 = sm_ld(var(AV_R));
                if (negate) S2 = -S2;
                const Fq R = S2 - sm_ld(var(AV_Y));
                if (P.is_zero()) {                                     // rare: the same point again, or its negative
                    if (R.is_zero()) {
                        Fq y = sm_ld(vy);
                        if (negate) y = -y;
                        const XYZZ<Fq> d = xyzz_dbl_affine<Fq>(sm_ld(vx), y);
                        sm_st(var(AV_X), d.X); sm_st(var(AV_Y), d.Y); sm_st(var(AV_ZZ), d.ZZ); sm_st(var(AV_ZZZ), d.ZZZ);
                    } else {
                        acc_inf = true;
                    }
                } else {
                    sm_st(var(AV_P), P); sm_st(var(AV_R), R);
                    fq_sqr_sm(var(AV_PP), var(AV_P));
                    fq_mul_sm(var(AV_PPP), var(AV_P), var(AV_PP));
                    fq_mul_sm(var(AV_Q), var(AV_X), var(AV_PP));
                    fq_sqr_sm(var(AV_P), var(AV_R));                  // R^2 (P is dead)
                    const Fq Q = sm_ld(var(AV_Q));
                    const Fq X3 = sm_ld(var(AV_P)) - sm_ld(var(AV_PPP)) - dbl(Q);
                    sm_st(var(AV_X), X3);
                    sm_st(var(AV_P), Q - X3);
                    fq_mul_sm(var(AV_P), var(AV_R), var(AV_P));       // R (Q - X3)
                    fq_mul_sm(var(AV_Q), var(AV_Y), var(AV_PPP));     // Y1 PPP
                    sm_st(var(AV_Y), sm_ld(var(AV_P)) - sm_ld(var(AV_Q)));
                    fq_mul_sm(var(AV_ZZ), var(AV_ZZ), var(AV_PP));
                    fq_mul_sm(var(AV_ZZZ), var(AV_ZZZ), var(AV_PPP));
                }
            }
        }
        const uint32_t nxt = pos + 1;
        if (nxt == bend || nxt == end) {
            const bool closes = nxt == bend;
            XYZZ<Fq> acc = XYZZ<Fq>::inf();
            if (!acc_inf) { acc.X = sm_ld(var(AV_X)); acc.Y = sm_ld(var(AV_Y)); acc.ZZ = sm_ld(var(AV_ZZ)); acc.ZZZ = sm_ld(var(AV_ZZZ)); }
            if (!head_open && closes) st_vec(bucket_sums + key, acc);
            else if (head_open) st_vec(part_head + chunk, acc);
            else { st_vec(part_tail + chunk, acc); tail_owner_key[chunk] = key; }
            head_open = false;
            acc_inf = true;
            if (nxt < end) {
                do { key++; } while (offsets[key + 1] <= nxt);
                bend = offsets[key + 1];
                if (tabs.nparts > 1 && (key >> nb) != part) {
                    part = key >> nb;
                    bases = tabs.tab[part];
                    part_end = offsets[(part + 1) << nb];
                }
            }
        }
    }
    cp_async_wait<0>();
}


// K5a: batched-affine pair rounds.  One round halves every bucket: the k points of a bucket become ceil(k/2)
// by adding neighbours pairwise in AFFINE coordinates — lambda = (y1-y0)/(x1-x0), 6 field products per addition
// instead of the 10 of the XYZZ mixed addition — with all the divisions of a WARP (32 x LP) sharing ONE field
// inversion (Montgomery's simultaneous-inversion trick).  A thread owns LP consecutive output points:
//   forward   d_j = x1 - x0 (or 2 y0 when the pair is a doubling), prefix products of the d_j in local memory
//   warp      prefix / suffix product scans of the lane totals over the lanes (shuffles) + one fp_inv_kaliski,
//             which runs on the ALU pipe; no block-level synchronisation
//   backward  1/d_j = I * prefix_j,  I *= d_j,  lambda, x3 = lambda^2 - x0 - x1, y3 = lambda (x0 - x3) - y0
// Exceptional pairs (P + P, P + (-P), infinity) are detected in the forward pass and keep the batch intact.
// After a few rounds the remaining points go through the XYZZ accumulation (DIRECT) as before.
#if !defined(__CUDA_ARCH__)
static inline Fq fq_mul_call(const Fq& a, const Fq& b) { return a * b; }      // host pass only parses the kernels
static inline Fq fq_sqr_call(const Fq& a) { return a * a; }
#endif
static __device__ __forceinline__ Fq fmul(const Fq& a, const Fq& b) { return fq_mul_call(a, b); }
static __device__ __forceinline__ Fq2 fmul(const Fq2& a, const Fq2& b) { return a * b; }
static __device__ __forceinline__ Fq fsqr(const Fq& a) { return fq_sqr_call(a); }
static __device__ __forceinline__ Fq2 fsqr(const Fq2& a) { return sqr(a); }
static __device__ __noinline__ Fq fq_inv_call(const Fq a) { return fp_inv_kaliski<FqParams>(a); }
static __device__ __forceinline__ Fq finv(const Fq& a) { return fq_inv_call(a); }
static __device__ __forceinline__ Fq2 finv(const Fq2& a) {
    const Fq n = fq_inv_call(fsqr(a.c0) + fsqr(a.c1));
    Fq2 r;
    r.c0 = fmul(a.c0, n);
    r.c1 = -fmul(a.c1, n);
    return r;
}
template <class T>
static __device__ __forceinline__ T shfl_up_vec(const T& v, int d) {
    T r;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (unsigned k = 0; k < sizeof(T) / 4; k++) o[k] = __shfl_up_sync(0xffffffffu, a[k], d);
    return r;
}
template <class T>
static __device__ __forceinline__ T shfl_down_vec(const T& v, int d) {
    T r;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (unsigned k = 0; k < sizeof(T) / 4; k++) o[k] = __shfl_down_sync(0xffffffffu, a[k], d);
    return r;
}
template <class T>
static __device__ __forceinline__ T shfl_idx_vec(const T& v, int src) {
    T r;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (unsigned k = 0; k < sizeof(T) / 4; k++) o[k] = __shfl_sync(0xffffffffu, a[k], src);
    return r;
}

enum { PR_COPY0 = 0, PR_COPY1 = 1, PR_INF = 2, PR_ADD = 3, PR_DBL = 4 };

// resolved source of output j: point index (bit 31 = negate y) of the first operand; the second one (pairs only)
// is the next entry / next point
template <class F, bool FIRST>
static __device__ __forceinline__ uint32_t pair_ref(const uint32_t* entries, uint32_t s) { return FIRST ? entries[s] : s; }
template <class F>
static __device__ __forceinline__ const Affine<F>* pair_ptr(const Affine<F>* pts, uint32_t ref) { return pts + (ref & 0x7fffffffu); }

template <class F, bool FIRST, int NT, int LP>
__global__ void __launch_bounds__(NT) __maxnreg__(sizeof(F) == sizeof(Fq) ? 128 : ZA_G2_MAXNREG) msm_pair_round_kernel(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ entries,
                                                                  const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out,
                                                                  uint32_t nkeys, Affine<F>* __restrict__ out) {
    constexpr int FB = sizeof(F) == sizeof(Fq) ? 4 : 2;      // forward pass: loads of FB outputs in flight per thread
    static_assert(LP % FB == 0, "LP must be a multiple of the forward batch");
    const uint32_t n_out = off_out[nkeys];
    const uint32_t tid = threadIdx.x;
    const uint64_t cta_first = (uint64_t)blockIdx.x * (NT * LP);
    if (cta_first + (uint64_t)(tid & ~31u) * LP >= n_out) return;       // the whole warp is past the end
    const uint64_t q0_64 = cta_first + (uint64_t)tid * LP;
    const uint32_t q0 = (uint32_t)q0_64;
    const uint32_t cnt = q0_64 < n_out ? (n_out - q0 < (uint32_t)LP ? n_out - q0 : (uint32_t)LP) : 0u;
    uint32_t ref0[LP], ref1[LP];                     // operands of output j (ref1 = 0xffffffff: no second operand)
    uint8_t flags[LP];
    F pre[LP];
    F run = F::one();
    if (cnt) {
        // ---- walk the buckets of this thread's outputs, resolve the operands
        uint32_t lo = 0, hi = nkeys;                  // last key with off_out[key] <= q0: the bucket of output q0
        while (lo + 1 < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (off_out[mid] <= q0) lo = mid; else hi = mid;
        }
        uint32_t key = lo;
        uint32_t obase = off_out[key], oend = off_out[key + 1], ibase = off_in[key], icnt = off_in[key + 1] - ibase;
#pragma unroll 4
        for (uint32_t j = 0; j < (uint32_t)LP; j++) {
            uint32_t r0 = 0, r1 = 0xffffffffu;
            if (j < cnt) {
                const uint32_t q = q0 + j;
                if (q == oend) {
                    do { key++; } while (off_out[key + 1] <= q);
                    obase = off_out[key]; oend = off_out[key + 1]; ibase = off_in[key]; icnt = off_in[key + 1] - ibase;
                }
                const uint32_t k = q - obase;
                const uint32_t s0 = ibase + 2 * k;
                r0 = pair_ref<F, FIRST>(entries, s0);
                if (2 * k + 1 < icnt) r1 = pair_ref<F, FIRST>(entries, s0 + 1);
            }
            ref0[j] = r0; ref1[j] = r1;
        }
        // ---- forward: x differences, FB outputs' loads in flight at a time
#pragma unroll 1
        for (uint32_t j0 = 0; j0 < cnt; j0 += FB) {
            F x0[FB], x1[FB];
#pragma unroll
            for (int u = 0; u < FB; u++) {
                const uint32_t j = j0 + u;
                const uint32_t r0 = ref0[j], r1 = ref1[j];     // j < LP always; entries beyond cnt are (0, none)
                x0[u] = ldg_vec(&pair_ptr<F>(pts, r0)->x);
                x1[u] = ldg_vec(&pair_ptr<F>(pts, r1 == 0xffffffffu ? r0 : r1)->x);
            }
            if (j0 + FB < cnt) {                       // the next batch's points on their way into L1 meanwhile
#pragma unroll
                for (int u = 0; u < FB; u++) {
                    const uint32_t r0 = ref0[j0 + FB + u], r1 = ref1[j0 + FB + u];
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pair_ptr<F>(pts, r0)));
                    if (r1 != 0xffffffffu) asm volatile("prefetch.global.L1 [%0];" ::"l"(pair_ptr<F>(pts, r1)));
                }
            }
#pragma unroll
            for (int u = 0; u < FB; u++) {
                const uint32_t j = j0 + u;
                if (j >= cnt) break;
                const uint32_t r0 = ref0[j], r1 = ref1[j];
                uint32_t fl = PR_COPY0;
                if (r1 != 0xffffffffu) {
                    F d = x1[u] - x0[u];
                    fl = PR_ADD;
                    if (d.is_zero() || x0[u].is_zero() || x1[u].is_zero()) {          // rare: look at the y coordinates
                        F y0 = ldg_vec(&pair_ptr<F>(pts, r0)->y), y1 = ldg_vec(&pair_ptr<F>(pts, r1)->y);
                        if (r0 >> 31) y0 = -y0;
                        if (r1 >> 31) y1 = -y1;
                        if (x0[u].is_zero() && y0.is_zero()) fl = PR_COPY1;
                        else if (x1[u].is_zero() && y1.is_zero()) fl = PR_COPY0;
                        else if (d.is_zero()) {
                            if (y0 == y1 && !y0.is_zero()) { fl = PR_DBL; d = dbl(y0); }
                            else fl = PR_INF;
                        }
                    }
                    if (fl >= PR_ADD) { pre[j] = run; run = fmul(run, d); }
                }
                flags[j] = (uint8_t)fl;
            }
        }
    }
    // ---- one inversion per warp: prefix and suffix products of the lane totals over the lanes, the warp product
    // inverted once (same value in every lane, so the branchy binary-Euclid inverse stays convergent), and
    // 1 / total_lane = ginv * (product of the lanes before) * (product of the lanes after).  No block-level
    // synchronisation: the warps of a CTA drift freely and hide each other's memory latency.
    F I;
    {
        const unsigned lane = tid & 31;
        F pf = run, sf = run;
#pragma unroll 1
        for (int d = 1; d < 32; d <<= 1) {
            F up = shfl_up_vec(pf, d), dn = shfl_down_vec(sf, d);
            if (lane < (unsigned)d) up = F::one();
            if (lane + d >= 32) dn = F::one();
            pf = fmul(pf, up);
            sf = fmul(sf, dn);
        }
        const F ginv = finv(shfl_idx_vec(pf, 31));
        F e = shfl_up_vec(pf, 1), sfx = shfl_down_vec(sf, 1);
        if (lane == 0) e = F::one();
        if (lane == 31) sfx = F::one();
        I = fmul(fmul(ginv, e), sfx);
    }
    // ---- backward, software pipelined by one output: the operand indices and the prefix product of output j-1
    // are loaded (local memory) and its two points prefetched into L1 while output j is computed
    if (!cnt) return;
    uint32_t nr0 = ref0[cnt - 1], nr1 = ref1[cnt - 1], nfl = flags[cnt - 1];
    F npre = pre[cnt - 1];
#pragma unroll 1
    for (uint32_t j = cnt; j-- > 0;) {
        const uint32_t r0 = nr0, r1 = nr1, fl = nfl;
        const F pj = npre;
        if (j) {
            nr0 = ref0[j - 1]; nr1 = ref1[j - 1]; nfl = flags[j - 1];
            npre = pre[j - 1];
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pair_ptr<F>(pts, nr0)));
            if (nr1 != 0xffffffffu) asm volatile("prefetch.global.L1 [%0];" ::"l"(pair_ptr<F>(pts, nr1)));
        }
        Affine<F> R;
        if (fl >= PR_ADD) {
            Affine<F> A = ldg_vec(pair_ptr<F>(pts, r0)), B = ldg_vec(pair_ptr<F>(pts, r1));
            if (r0 >> 31) A.y = -A.y;
            if (r1 >> 31) B.y = -B.y;
            F d = B.x - A.x, num = B.y - A.y, xs = A.x + B.x;
            if (fl == PR_DBL) {
                const F xx = fsqr(A.x);
                d = dbl(A.y); num = dbl(xx) + xx; xs = dbl(A.x);
            }
            const F id = fmul(I, pj);
            I = fmul(I, d);
            const F lam = fmul(num, id);
            R.x = fsqr(lam) - xs;
            R.y = fmul(lam, A.x - R.x) - A.y;
        } else if (fl == PR_INF) {
            R = Affine<F>::inf();
        } else {
            const uint32_t r = fl == PR_COPY1 ? r1 : r0;
            R = ldg_vec(pair_ptr<F>(pts, r));
            if (r >> 31) R.y = -R.y;
        }
        st_vec(out + q0 + j, R);
    }
}

// K5a' — the pair rounds over Fq2 with one Fq2 value split over a LANE PAIR (lane 2k holds c0, lane 2k + 1 holds c1).
// The single-thread version above needs ~190-224 registers (8 warps per SM: latency bound, multiply pipe ~60 % busy,
// 2.3 KB of local memory per thread).  Here every lane carries half of every value: the working set lives in shared
// memory (12 slots of 32 bytes per lane, the layout of the G1 accumulation), the kernel fits 128 registers and four
// CTAs per SM (16 warps), and the Fq2 product is convergent over the pair:
//     lane 0:  c0 = a0 b0 + a1 (p - b1)        lane 1:  c1 = a0 b1 + a1 b0
// i.e. both lanes run  x0 y0 + x1 y1  (two 512-bit products, one addition, ONE Montgomery reduction) on operands they
// read from the pair's two slots; the squaring is one product per lane ((a0 + a1)(a0 - a1) and (2 a0) a1).  Per Fq2
// product that is 4 wide products + 2 reductions (the three-product Karatsuba cannot be balanced over two lanes, and
// the pipe is paid per warp instruction), the same count as round 1's three full products, but at twice the occupancy
// and without argument marshalling or spills.
#define LPK_NT 128
enum { LV_I = 0, LV_PJ, LV_AX, LV_AY, LV_BX, LV_BY, LV_D, LV_NUM, LV_XS, LV_T, LV_PF, LV_SF, LV_COUNT };
#define LPK_SM_BYTES (LV_COUNT * 32 * LPK_NT)
static_assert(LPK_NT == ACC_SM_NT, "sm_ld / sm_st address the second half of a value ACC_SM_NT * 16 bytes further");
static __device__ __forceinline__ void pair_sync() { __syncwarp(3u << (threadIdx.x & 30u)); }
static __device__ __forceinline__ bool pair_all(bool v) {
    const unsigned m = 3u << (threadIdx.x & 30u);
    return __all_sync(m, v);
}
#if defined(__CUDA_ARCH__)
// slot d = this lane's component of (slot a) * (slot b); a, b, d are the CALLING lane's slot addresses
static __device__ __noinline__ void lp_mul_sm(uint32_t d, uint32_t a, uint32_t b) {
    const uint32_t h = threadIdx.x & 1u;
    const uint32_t a_c0 = a - h * 16u;
    const Fq x0 = sm_ld(a_c0), x1 = sm_ld(a_c0 + 16u);
    const Fq y0 = sm_ld(b);
    Fq yo = sm_ld(h ? b - 16u : b + 16u);
    {   // lane 0 multiplies a1 by -b1 = p - b1 (p itself when b1 = 0: the product is then a multiple of p)
        uint32_t n[8];
        n[0] = p_sub_cc(FqParams::mod(0), yo.v[0]);
#pragma unroll
        for (int i = 1; i < 7; i++) n[i] = p_subc_cc(FqParams::mod(i), yo.v[i]);
        n[7] = p_subc(FqParams::mod(7), yo.v[7]);
#pragma unroll
        for (int i = 0; i < 8; i++) yo.v[i] = h ? yo.v[i] : n[i];
    }
    pair_sync();                                  // both lanes have read before either overwrites (d may be a or b)
    uint32_t T[16], U[16];
    u256_mul_wide(T, x0.v, y0.v);
    u256_mul_wide(U, x1.v, yo.v);
    T[0] = p_add_cc(T[0], U[0]);
#pragma unroll
    for (int i = 1; i < 15; i++) T[i] = p_addc_cc(T[i], U[i]);
    T[15] = p_addc(T[15], U[15]);                 // x0 y0 + x1 y1 <= 2 p^2 < p 2^256
    sm_st(d, fp_redc_wide<FqParams>(T));
    pair_sync();
}
// slot d = this lane's component of (slot a)^2:  c0 = (a0 + a1)(a0 - a1),  c1 = (2 a0) a1
static __device__ __noinline__ void lp_sqr_sm(uint32_t d, uint32_t a) {
    const uint32_t h = threadIdx.x & 1u;
    const uint32_t a_c0 = a - h * 16u;
    const Fq a0 = sm_ld(a_c0), a1 = sm_ld(a_c0 + 16u);
    Fq t;
#pragma unroll
    for (int i = 0; i < 8; i++) t.v[i] = h ? a0.v[i] : a1.v[i];
    const Fq u = a0 + t;                          // lane 0: a0 + a1, lane 1: 2 a0
    const Fq w = a0 - a1;
    Fq v;
#pragma unroll
    for (int i = 0; i < 8; i++) v.v[i] = h ? a1.v[i] : w.v[i];
    pair_sync();
    sm_st(d, fp_mul<FqParams>(u, v));
    pair_sync();
}
#else
static inline void lp_mul_sm(uint32_t, uint32_t, uint32_t) {}
static inline void lp_sqr_sm(uint32_t, uint32_t) {}
#endif
static __device__ __forceinline__ Fq shfl_fq(const Fq& v, int src_lane) {
    Fq r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = __shfl_sync(0xffffffffu, v.v[k], src_lane);
    return r;
}
// component h of the point's x (which = 0) or y (which = 1) coordinate
static __device__ __forceinline__ const Fq* lp_coord(const Affine<Fq2>* p, uint32_t ref, uint32_t which, uint32_t h) {
    return reinterpret_cast<const Fq*>(p + (ref & 0x7fffffffu)) + 2 * which + h;
}

template <bool FIRST, int LP>
__global__ void __launch_bounds__(LPK_NT, 4) msm_pair_round_g2lp_kernel(const Affine<Fq2>* __restrict__ pts, const uint32_t* __restrict__ entries,
                                                                        const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out,
                                                                        uint32_t nkeys, Affine<Fq2>* __restrict__ out) {
    extern __shared__ uint4 lpk_sm[];
    constexpr int FB = 2;
    static_assert(LP % FB == 0, "LP must be a multiple of the forward batch");
    const uint32_t n_out = off_out[nkeys];
    const uint32_t tid = threadIdx.x, h = tid & 1u, lane = tid & 31u;
    const uint64_t cta_first = (uint64_t)blockIdx.x * ((LPK_NT / 2) * LP);
    if (cta_first + (uint64_t)((tid & ~31u) >> 1) * LP >= n_out) return;       // the whole warp is past the end
    const uint64_t q0_64 = cta_first + (uint64_t)(tid >> 1) * LP;
    const uint32_t q0 = (uint32_t)q0_64;
    const uint32_t cnt = q0_64 < n_out ? (n_out - q0 < (uint32_t)LP ? n_out - q0 : (uint32_t)LP) : 0u;
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(lpk_sm) + tid * 16u;
    auto var = [&](int v) { return sm0 + (uint32_t)v * (2 * LPK_NT * 16); };
    const Fq one_h = h ? Fq::zero() : Fq::one();                               // this lane's component of 1
    uint32_t ref0[LP], ref1[LP];
    uint8_t flags[LP];
    Fq pre[LP];
    sm_st(var(LV_I), one_h);                                                    // running product of the differences
    if (cnt) {
        uint32_t lo = 0, hi = nkeys;
        while (lo + 1 < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (off_out[mid] <= q0) lo = mid; else hi = mid;
        }
        uint32_t key = lo;
        uint32_t obase = off_out[key], oend = off_out[key + 1], ibase = off_in[key], icnt = off_in[key + 1] - ibase;
#pragma unroll 4
        for (uint32_t j = 0; j < (uint32_t)LP; j++) {
            uint32_t r0 = 0, r1 = 0xffffffffu;
            if (j < cnt) {
                const uint32_t q = q0 + j;
                if (q == oend) {
                    do { key++; } while (off_out[key + 1] <= q);
                    obase = off_out[key]; oend = off_out[key + 1]; ibase = off_in[key]; icnt = off_in[key + 1] - ibase;
                }
                const uint32_t k = q - obase;
                const uint32_t s0 = ibase + 2 * k;
                r0 = FIRST ? entries[s0] : s0;
                if (2 * k + 1 < icnt) r1 = FIRST ? entries[s0 + 1] : s0 + 1;
            }
            ref0[j] = r0; ref1[j] = r1;
        }
        // ---- forward: x differences, FB outputs' loads in flight at a time
#pragma unroll 1
        for (uint32_t j0 = 0; j0 < cnt; j0 += FB) {
            Fq x0[FB], x1[FB];
#pragma unroll
            for (int u = 0; u < FB; u++) {
                const uint32_t r0 = ref0[j0 + u], r1 = ref1[j0 + u];
                x0[u] = ldg_vec(lp_coord(pts, r0, 0, h));
                x1[u] = ldg_vec(lp_coord(pts, r1 == 0xffffffffu ? r0 : r1, 0, h));
            }
            if (j0 + FB < cnt) {
#pragma unroll
                for (int u = 0; u < FB; u++) {
                    const uint32_t r0 = ref0[j0 + FB + u], r1 = ref1[j0 + FB + u];
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(lp_coord(pts, r0, 0, 0)));
                    if (r1 != 0xffffffffu) asm volatile("prefetch.global.L1 [%0];" ::"l"(lp_coord(pts, r1, 0, 0)));
                }
            }
#pragma unroll
            for (int u = 0; u < FB; u++) {
                const uint32_t j = j0 + u;
                if (j >= cnt) break;
                const uint32_t r0 = ref0[j], r1 = ref1[j];
                uint32_t fl = PR_COPY0;
                if (r1 != 0xffffffffu) {
                    Fq d = x1[u] - x0[u];
                    fl = PR_ADD;
                    const bool dz = pair_all(d.is_zero()), x0z = pair_all(x0[u].is_zero()), x1z = pair_all(x1[u].is_zero());
                    if (dz || x0z || x1z) {                                     // rare: look at the y coordinates
                        Fq y0 = ldg_vec(lp_coord(pts, r0, 1, h)), y1 = ldg_vec(lp_coord(pts, r1, 1, h));
                        if (r0 >> 31) y0 = -y0;
                        if (r1 >> 31) y1 = -y1;
                        const bool y0z = pair_all(y0.is_zero()), y1z = pair_all(y1.is_zero()), yeq = pair_all(y0 == y1);
                        if (x0z && y0z) fl = PR_COPY1;
                        else if (x1z && y1z) fl = PR_COPY0;
                        else if (dz) {
                            if (yeq && !y0z) { fl = PR_DBL; d = dbl(y0); }
                            else fl = PR_INF;
                        }
                    }
                    if (fl >= PR_ADD) {
                        pre[j] = sm_ld(var(LV_I));
                        sm_st(var(LV_D), d);
                        pair_sync();
                        lp_mul_sm(var(LV_I), var(LV_I), var(LV_D));
                    }
                }
                flags[j] = (uint8_t)fl;
            }
        }
    }
    __syncwarp();
    // ---- one inversion per warp over its 16 pairs: prefix and suffix products of the pair totals (shuffles by an even
    // number of lanes keep the component), the warp product inverted once (every lane computes the same norm, so the
    // binary-Euclid inverse stays convergent), 1 / total_pair = ginv * (product of the pairs before) * (... after)
    {
        const uint32_t pidx = lane >> 1;
        sm_st(var(LV_PF), sm_ld(var(LV_I)));
        sm_st(var(LV_SF), sm_ld(var(LV_I)));
        __syncwarp();
#pragma unroll 1
        for (int d = 1; d < 16; d <<= 1) {
            Fq up = shfl_up_vec(sm_ld(var(LV_PF)), 2 * d), dn = shfl_down_vec(sm_ld(var(LV_SF)), 2 * d);
            if (pidx < (uint32_t)d) up = one_h;
            if (pidx + d >= 16) dn = one_h;
            sm_st(var(LV_T), up);
            sm_st(var(LV_D), dn);
            __syncwarp();
            lp_mul_sm(var(LV_PF), var(LV_PF), var(LV_T));
            lp_mul_sm(var(LV_SF), var(LV_SF), var(LV_D));
        }
        const Fq pf = sm_ld(var(LV_PF)), sf = sm_ld(var(LV_SF));
        const Fq tot = shfl_fq(pf, 30 + (int)h);                               // this lane's component of the warp product
        const Fq sq = fq_sqr_call(tot);
        Fq oth;
#pragma unroll
        for (int k = 0; k < 8; k++) oth.v[k] = __shfl_xor_sync(0xffffffffu, sq.v[k], 1);
        const Fq ninv = fq_inv_call(sq + oth);                                 // 1 / (t0^2 + t1^2)
        Fq g = fq_mul_call(tot, ninv);
        if (h) g = -g;                                                         // (t0 - t1 u) / norm
        Fq e = shfl_up_vec(pf, 2), sfx = shfl_down_vec(sf, 2);
        if (pidx == 0) e = one_h;
        if (pidx == 15) sfx = one_h;
        sm_st(var(LV_I), g); sm_st(var(LV_T), e); sm_st(var(LV_D), sfx);
        __syncwarp();
        lp_mul_sm(var(LV_I), var(LV_I), var(LV_T));
        lp_mul_sm(var(LV_I), var(LV_I), var(LV_D));
    }
    // ---- backward
    if (!cnt) return;
    uint32_t nr0 = ref0[cnt - 1], nr1 = ref1[cnt - 1], nfl = flags[cnt - 1];
    Fq npre = pre[cnt - 1];
#pragma unroll 1
    for (uint32_t j = cnt; j-- > 0;) {
        const uint32_t r0 = nr0, r1 = nr1, fl = nfl;
        const Fq pj = npre;
        if (j) {                                   // the next output's operands and prefix product one iteration ahead
            nr0 = ref0[j - 1]; nr1 = ref1[j - 1]; nfl = flags[j - 1];
            npre = pre[j - 1];
            asm volatile("prefetch.global.L1 [%0];" ::"l"(lp_coord(pts, nr0, 0, 0)));
            if (nr1 != 0xffffffffu) asm volatile("prefetch.global.L1 [%0];" ::"l"(lp_coord(pts, nr1, 0, 0)));
        }
        Fq rx, ry;
        if (fl >= PR_ADD) {
            Fq ax = ldg_vec(lp_coord(pts, r0, 0, h)), ay = ldg_vec(lp_coord(pts, r0, 1, h));
            Fq bx = ldg_vec(lp_coord(pts, r1, 0, h)), by = ldg_vec(lp_coord(pts, r1, 1, h));
            if (r0 >> 31) ay = -ay;
            if (r1 >> 31) by = -by;
            sm_st(var(LV_AX), ax); sm_st(var(LV_AY), ay);
            sm_st(var(LV_PJ), pj);
            if (fl == PR_DBL) {
                pair_sync();
                lp_sqr_sm(var(LV_T), var(LV_AX));
                const Fq xx = sm_ld(var(LV_T));
                sm_st(var(LV_D), dbl(ay)); sm_st(var(LV_NUM), dbl(xx) + xx); sm_st(var(LV_XS), dbl(ax));
            } else {
                sm_st(var(LV_D), bx - ax); sm_st(var(LV_NUM), by - ay); sm_st(var(LV_XS), ax + bx);
            }
            pair_sync();
            lp_mul_sm(var(LV_PJ), var(LV_I), var(LV_PJ));          // 1 / d
            lp_mul_sm(var(LV_I), var(LV_I), var(LV_D));
            lp_mul_sm(var(LV_NUM), var(LV_NUM), var(LV_PJ));       // lambda
            lp_sqr_sm(var(LV_T), var(LV_NUM));
            rx = sm_ld(var(LV_T)) - sm_ld(var(LV_XS));
            sm_st(var(LV_T), sm_ld(var(LV_AX)) - rx);
            pair_sync();
            lp_mul_sm(var(LV_T), var(LV_NUM), var(LV_T));
            ry = sm_ld(var(LV_T)) - sm_ld(var(LV_AY));
        } else if (fl == PR_INF) {
            rx = Fq::zero(); ry = Fq::zero();
        } else {
            const uint32_t r = fl == PR_COPY1 ? r1 : r0;
            rx = ldg_vec(lp_coord(pts, r, 0, h)); ry = ldg_vec(lp_coord(pts, r, 1, h));
            if (r >> 31) ry = -ry;
        }
        Fq* o = reinterpret_cast<Fq*>(out + q0 + j);
        st_vec(o + h, rx);
        st_vec(o + 2 + h, ry);
    }
}

// K5b: buckets that straddle chunk boundaries.  Element 0 of a chain is the tail partial of the chunk where
// the bucket starts, element k the head partial of chunk + k.  Short chains (the common case) are folded
// by one thread; long chains (one bucket holding a large share of all entries — witness vectors full of
// ones) are queued and folded warp-cooperatively by msm_fixup_long_kernel.
#define FIXUP_SHORT 6
// K5a'' — the single-thread pair rounds over Fq2 with the WORKING SET IN SHARED MEMORY.
// msm_pair_round_kernel<Fq2> needs 224 registers: two CTAs = 8 warps per SM, `wait` stalls with nothing else to issue and a
// fifth of the time in long-scoreboard stalls behind the gathers (profiles/r02_g2_lanepair.md).  The lane-pair kernel above
// bought 16 warps with 4 wide products per Fq2 product and an inversion per 512 outputs, and lost.  This kernel keeps one
// thread per output stream (3 wide products per Fq2 product, one inversion per 1024 outputs) and moves the six Fq2 values a
// thread works on into shared memory, the scheme of msm_accumulate_g1_sm_kernel: the Fq2 product, the squaring and the
// inversion are non-inlined routines that take shared-memory addresses (LDS.128 into the registers they read, STS.128 from
// the registers they produce, no argument marshalling), so the kernel body holds addresses and counters only: <= 128
// registers, 48 KB per CTA, four CTAs = 16 warps per SM.  Prefix products and operand indices stay in local memory as before.
#define G2SM_NT 128
// operands of a later output on their way while the current one is computed.  With four CTAs of 48 KB the L1 that is left
// (~64 KB for 512 threads x 2 points x 128 B) does not hold them: ZA_G2SM_PF = 1 / 2 prefetches into L2, one / two outputs ahead
#ifndef ZA_G2SM_PF
#define ZA_G2SM_PF 0
#endif
#define G2SM_PF_AHEAD (ZA_G2SM_PF == 2 ? 2u : 1u)
static __device__ __forceinline__ void g2sm_prefetch(const void* p) {
#if ZA_G2SM_PF == 0
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
enum { GV_I = 0, GV_P, GV_D, GV_X, GV_N, GV_A, GV_COUNT };      // Fq2 slots, see the kernel
#define G2SM_BYTES (GV_COUNT * 64 * G2SM_NT)
static_assert(G2SM_NT == ACC_SM_NT, "sm_ld / sm_st address the second half of a value ACC_SM_NT * 16 bytes further");
#define G2SM_C1 (2u * G2SM_NT * 16u)                            /* from the c0 component of a slot to its c1 component */
static __device__ __forceinline__ Fq2 g2sm_ld(uint32_t a) { Fq2 r; r.c0 = sm_ld(a); r.c1 = sm_ld(a + G2SM_C1); return r; }
static __device__ __forceinline__ void g2sm_st(uint32_t a, const Fq2& v) { sm_st(a, v.c0); sm_st(a + G2SM_C1, v.c1); }
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ void fq2_mul_sm(uint32_t d, uint32_t a, uint32_t b) { g2sm_st(d, fq2_mul_lazy(g2sm_ld(a), g2sm_ld(b))); }
static __device__ __noinline__ void fq2_sqr_sm(uint32_t d, uint32_t a) {
    const Fq2 v = g2sm_ld(a);
    Fq2 r;
    r.c0 = fp_mul<FqParams>(v.c0 + v.c1, v.c0 - v.c1);
    r.c1 = dbl(fp_mul<FqParams>(v.c0, v.c1));
    g2sm_st(d, r);
}
static __device__ __noinline__ void fq2_inv_sm(uint32_t d, uint32_t a) {          // once per thread and round
    const Fq2 v = g2sm_ld(a);
    const Fq n = fp_inv_kaliski<FqParams>(fp_sqr<FqParams>(v.c0) + fp_sqr<FqParams>(v.c1));
    Fq2 r;
    r.c0 = fp_mul<FqParams>(v.c0, n);
    r.c1 = -fp_mul<FqParams>(v.c1, n);
    g2sm_st(d, r);
}
// the rare operands of the forward pass (equal x, or a coordinate that is zero): classify, and for a doubling put 2 y into slot d
static __device__ __noinline__ uint32_t g2sm_classify(const Affine<Fq2>* pts, uint32_t r0, uint32_t r1, uint32_t d_slot) {
    const Fq2 x0 = ldg_vec(&pair_ptr<Fq2>(pts, r0)->x), x1 = ldg_vec(&pair_ptr<Fq2>(pts, r1)->x);
    Fq2 y0 = ldg_vec(&pair_ptr<Fq2>(pts, r0)->y), y1 = ldg_vec(&pair_ptr<Fq2>(pts, r1)->y);
    if (r0 >> 31) y0 = -y0;
    if (r1 >> 31) y1 = -y1;
    const Fq2 d = x1 - x0;
    if (x0.is_zero() && y0.is_zero()) return PR_COPY1;
    if (x1.is_zero() && y1.is_zero()) return PR_COPY0;
    if (d.is_zero()) {
        if (y0 == y1 && !y0.is_zero()) { g2sm_st(d_slot, dbl(y0)); return PR_DBL; }
        return PR_INF;
    }
    g2sm_st(d_slot, d);
    return PR_ADD;
}
// backward pass, doubling: d = 2 y, num = 3 x^2, xs = 2 x  (slots D, N, X; A = x)
static __device__ __noinline__ void g2sm_dbl_setup(const Affine<Fq2>* pts, uint32_t r0, uint32_t sD, uint32_t sN, uint32_t sX, uint32_t sA) {
    const Fq2 ax = ldg_vec(&pair_ptr<Fq2>(pts, r0)->x);
    Fq2 ay = ldg_vec(&pair_ptr<Fq2>(pts, r0)->y);
    if (r0 >> 31) ay = -ay;
    g2sm_st(sA, ax);
    g2sm_st(sD, dbl(ay));
    g2sm_st(sX, dbl(ax));
    fq2_sqr_sm(sN, sA);
    const Fq2 xx = g2sm_ld(sN);
    g2sm_st(sN, dbl(xx) + xx);
}
#else
static inline void fq2_mul_sm(uint32_t, uint32_t, uint32_t) {}
static inline void fq2_sqr_sm(uint32_t, uint32_t) {}
static inline void fq2_inv_sm(uint32_t, uint32_t) {}
static inline uint32_t g2sm_classify(const Affine<Fq2>*, uint32_t, uint32_t, uint32_t) { return 0; }
static inline void g2sm_dbl_setup(const Affine<Fq2>*, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t) {}
#endif

template <bool FIRST, int LP>
__global__ void __launch_bounds__(G2SM_NT, 4) msm_pair_round_g2sm_kernel(const Affine<Fq2>* __restrict__ pts, const uint32_t* __restrict__ entries,
                                                                        const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out,
                                                                        uint32_t nkeys, Affine<Fq2>* __restrict__ out) {
    extern __shared__ uint4 g2sm_raw[];
    const uint32_t n_out = off_out[nkeys];
    const uint32_t tid = threadIdx.x;
    const uint64_t cta_first = (uint64_t)blockIdx.x * (G2SM_NT * LP);
    if (cta_first + (uint64_t)(tid & ~31u) * LP >= n_out) return;       // the whole warp is past the end
    const uint64_t q0_64 = cta_first + (uint64_t)tid * LP;
    const uint32_t q0 = (uint32_t)q0_64;
    const uint32_t cnt = q0_64 < n_out ? (n_out - q0 < (uint32_t)LP ? n_out - q0 : (uint32_t)LP) : 0u;
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(g2sm_raw) + tid * 16u;
    // slot V (an Fq2 value): c0 at sl(V), c1 at sl(V) + G2SM_C1.  I: running product, later the running inverse; P: prefix
    // product of the output, later 1 / d; D: x difference, later lambda^2, later lambda (ax - rx); X: x sum, later the result's
    // x; N: numerator, later lambda; A: the first operand's x
    auto sl = [&](int v) { return sm0 + (uint32_t)v * (2u * G2SM_C1); };
    uint32_t ref0[LP], ref1[LP];
    uint8_t flags[LP];
    Fq2 pre[LP];
    g2sm_st(sl(GV_I), Fq2::one());
    if (cnt) {
        uint32_t lo = 0, hi = nkeys;
        while (lo + 1 < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (off_out[mid] <= q0) lo = mid; else hi = mid;
        }
        uint32_t key = lo;
        uint32_t obase = off_out[key], oend = off_out[key + 1], ibase = off_in[key], icnt = off_in[key + 1] - ibase;
#pragma unroll 4
        for (uint32_t j = 0; j < (uint32_t)LP; j++) {
            uint32_t r0 = 0, r1 = 0xffffffffu;
            if (j < cnt) {
                const uint32_t q = q0 + j;
                if (q == oend) {
                    do { key++; } while (off_out[key + 1] <= q);
                    obase = off_out[key]; oend = off_out[key + 1]; ibase = off_in[key]; icnt = off_in[key + 1] - ibase;
                }
                const uint32_t k = q - obase;
                const uint32_t s0 = ibase + 2 * k;
                r0 = FIRST ? entries[s0] : s0;
                if (2 * k + 1 < icnt) r1 = FIRST ? entries[s0 + 1] : s0 + 1;
            }
            ref0[j] = r0; ref1[j] = r1;
        }
        // ---- forward: running product of the x differences.  Software pipelined: the operand indices (local memory) are
        // read two outputs ahead and the points of the next output are prefetched while the current product runs
        uint32_t c0 = ref0[0], c1 = ref1[0];
        uint32_t n0 = cnt > 1 ? ref0[1] : 0u, n1 = cnt > 1 ? ref1[1] : 0xffffffffu;
#pragma unroll 1
        for (uint32_t j = 0; j < cnt; j++) {
            const uint32_t r0 = c0, r1 = c1;
            c0 = n0; c1 = n1;
            if (j + 2 < cnt) { n0 = ref0[j + 2]; n1 = ref1[j + 2]; }
            if (j + 1 < cnt) {
                g2sm_prefetch(pair_ptr<Fq2>(pts, c0));
                if (c1 != 0xffffffffu) g2sm_prefetch(pair_ptr<Fq2>(pts, c1));
            }
            uint32_t fl = PR_COPY0;
            if (r1 != 0xffffffffu) {
                const Fq2 x0 = ldg_vec(&pair_ptr<Fq2>(pts, r0)->x), x1 = ldg_vec(&pair_ptr<Fq2>(pts, r1)->x);
                const Fq2 d = x1 - x0;
                fl = PR_ADD;
                if (d.is_zero() || x0.is_zero() || x1.is_zero()) fl = g2sm_classify(pts, r0, r1, sl(GV_D));
                else g2sm_st(sl(GV_D), d);
                if (fl >= PR_ADD) {
                    pre[j] = g2sm_ld(sl(GV_I));
                    fq2_mul_sm(sl(GV_I), sl(GV_I), sl(GV_D));
                }
            }
            flags[j] = (uint8_t)fl;
        }
    }
    // ---- one inversion per warp (as in msm_pair_round_kernel): prefix products in slot P, suffix products in slot X
    {
        const unsigned lane = tid & 31;
        const Fq2 run = g2sm_ld(sl(GV_I));
        g2sm_st(sl(GV_P), run);
        g2sm_st(sl(GV_X), run);
#pragma unroll 1
        for (int d = 1; d < 32; d <<= 1) {
            Fq2 up = shfl_up_vec(g2sm_ld(sl(GV_P)), d);
            if (lane < (unsigned)d) up = Fq2::one();
            g2sm_st(sl(GV_D), up);
            Fq2 dn = shfl_down_vec(g2sm_ld(sl(GV_X)), d);
            if (lane + d >= 32) dn = Fq2::one();
            g2sm_st(sl(GV_N), dn);
            fq2_mul_sm(sl(GV_P), sl(GV_P), sl(GV_D));
            fq2_mul_sm(sl(GV_X), sl(GV_X), sl(GV_N));
        }
        g2sm_st(sl(GV_N), shfl_idx_vec(g2sm_ld(sl(GV_P)), 31));
        fq2_inv_sm(sl(GV_N), sl(GV_N));                                   // the same value in every lane: convergent
        Fq2 e = shfl_up_vec(g2sm_ld(sl(GV_P)), 1), sfx = shfl_down_vec(g2sm_ld(sl(GV_X)), 1);
        if (lane == 0) e = Fq2::one();
        if (lane == 31) sfx = Fq2::one();
        g2sm_st(sl(GV_D), e);
        g2sm_st(sl(GV_A), sfx);
        fq2_mul_sm(sl(GV_I), sl(GV_N), sl(GV_D));
        fq2_mul_sm(sl(GV_I), sl(GV_I), sl(GV_A));                         // I = 1 / (this lane's product)
    }
    if (!cnt) return;
    // ---- backward, software pipelined the same way (indices and flag two outputs ahead, points and the prefix product of
    // the next output prefetched)
    uint32_t b0 = ref0[cnt - 1], b1 = ref1[cnt - 1], bf = flags[cnt - 1];
    uint32_t m0 = cnt > 1 ? ref0[cnt - 2] : 0u, m1 = cnt > 1 ? ref1[cnt - 2] : 0xffffffffu, mf = cnt > 1 ? flags[cnt - 2] : 0u;
#pragma unroll 1
    for (uint32_t j = cnt; j-- > 0;) {
        const uint32_t r0 = b0, r1 = b1, fl = bf;
        b0 = m0; b1 = m1; bf = mf;
        if (j >= 2) { m0 = ref0[j - 2]; m1 = ref1[j - 2]; mf = flags[j - 2]; }
        if (j >= 1) {
            g2sm_prefetch(pair_ptr<Fq2>(pts, b0));
            if (b1 != 0xffffffffu) g2sm_prefetch(pair_ptr<Fq2>(pts, b1));
            asm volatile("prefetch.L1 [%0];" ::"l"(&pre[j - 1]));      // generic address of the local array
        }
        Affine<Fq2> R;
        if (fl >= PR_ADD) {
            g2sm_st(sl(GV_P), pre[j]);
            if (fl == PR_DBL) {
                g2sm_dbl_setup(pts, r0, sl(GV_D), sl(GV_N), sl(GV_X), sl(GV_A));
            } else {
                {
                    const Fq2 ax = ldg_vec(&pair_ptr<Fq2>(pts, r0)->x), bx = ldg_vec(&pair_ptr<Fq2>(pts, r1)->x);
                    g2sm_st(sl(GV_D), bx - ax);
                    g2sm_st(sl(GV_X), ax + bx);
                    g2sm_st(sl(GV_A), ax);
                }
                Fq2 ay = ldg_vec(&pair_ptr<Fq2>(pts, r0)->y), by = ldg_vec(&pair_ptr<Fq2>(pts, r1)->y);
                if (r0 >> 31) ay = -ay;
                if (r1 >> 31) by = -by;
                g2sm_st(sl(GV_N), by - ay);
            }
            fq2_mul_sm(sl(GV_P), sl(GV_I), sl(GV_P));                     // 1 / d = I * prefix
            fq2_mul_sm(sl(GV_I), sl(GV_I), sl(GV_D));                     // I *= d
            fq2_mul_sm(sl(GV_N), sl(GV_N), sl(GV_P));                     // lambda
            fq2_sqr_sm(sl(GV_D), sl(GV_N));                               // lambda^2
            R.x = g2sm_ld(sl(GV_D)) - g2sm_ld(sl(GV_X));
            g2sm_st(sl(GV_D), g2sm_ld(sl(GV_A)) - R.x);
            fq2_mul_sm(sl(GV_D), sl(GV_N), sl(GV_D));                     // lambda (ax - rx)
            Fq2 ay = ldg_vec(&pair_ptr<Fq2>(pts, r0)->y);                 // again (L1): it was not kept across the products
            if (r0 >> 31) ay = -ay;
            R.y = g2sm_ld(sl(GV_D)) - ay;
        } else if (fl == PR_INF) {
            R = Affine<Fq2>::inf();
        } else {
            const uint32_t r = fl == PR_COPY1 ? r1 : r0;
            R = ldg_vec(pair_ptr<Fq2>(pts, r));
            if (r >> 31) R.y = -R.y;
        }
        st_vec(out + q0 + j, R);
    }
}

template <class F>
__global__ void __launch_bounds__(128) msm_fixup_kernel(const uint32_t* __restrict__ offsets, uint32_t nkeys, uint32_t Lc, uint32_t nchunks,
                                                        const XYZZ<F>* part_head, const XYZZ<F>* part_tail, const uint32_t* tail_owner_key,
                                                        XYZZ<F>* bucket_sums, uint32_t* long_list, uint32_t* long_count) {
    const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk >= nchunks) return;
    const uint32_t key = tail_owner_key[chunk];
    if (key == 0xffffffffu) return;
    const uint32_t bend = offsets[key + 1];
    const uint32_t t_end = (bend - 1) / Lc;        // last chunk holding entries of this bucket
    const uint32_t count = t_end - chunk + 1;
    if (count > FIXUP_SHORT) { long_list[atomicAdd(long_count, 1u)] = chunk; return; }
    XYZZ<F> acc = ld_vec(part_tail + chunk);
    for (uint32_t k = 1; k < count; k++) {
        XYZZ<F> v = ld_vec(part_head + chunk + k);
        xyzz_add<F>(acc, v);
    }
    st_vec(bucket_sums + key, acc);
}

template <class F>
static __device__ __forceinline__ XYZZ<F> warp_reduce_xyzz(XYZZ<F> acc, unsigned lane) {
    constexpr int WORDS = sizeof(XYZZ<F>) / 4;
#pragma unroll 1
    for (int step = 16; step >= 1; step >>= 1) {
        XYZZ<F> other;
        uint32_t* o = reinterpret_cast<uint32_t*>(&other);
        const uint32_t* a = reinterpret_cast<const uint32_t*>(&acc);
#pragma unroll
        for (int k = 0; k < WORDS; k++) o[k] = __shfl_down_sync(0xffffffffu, a[k], step);
        if (lane < (unsigned)step) xyzz_add<F>(acc, other);
    }
    return acc;
}

template <class F>
__global__ void __launch_bounds__(128) msm_fixup_long_kernel(const uint32_t* __restrict__ offsets, uint32_t Lc, const XYZZ<F>* part_head,
                                                             const XYZZ<F>* part_tail, const uint32_t* tail_owner_key, XYZZ<F>* bucket_sums,
                                                             const uint32_t* long_list, const uint32_t* long_count) {
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    const uint32_t total = *long_count;
    for (uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < total; item += nwarps) {
        const uint32_t chunk = long_list[item];
        const uint32_t key = tail_owner_key[chunk];
        const uint32_t bend = offsets[key + 1];
        const uint32_t count = (bend - 1) / Lc - chunk + 1;
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t k = lane; k < count; k += 32) {
            XYZZ<F> v = k == 0 ? ld_vec(part_tail + chunk) : ld_vec(part_head + chunk + k);
            xyzz_add<F>(acc, v);
        }
        acc = warp_reduce_xyzz<F>(acc, lane);
        if (lane == 0) st_vec(bucket_sums + key, acc);
    }
}

// K6a: one level of the weighted bucket reduction.  For window w and segment j of s items X_i:
//   T_j = sum X_i,  acc_j = sum (i - lo_j) X_i   (running sums, 2(s-1) additions, no scalar multiplication)
// sum_i i X_i = sum_j acc_j + s * sum_j j T_j, so the T_j are the items of the next level; acc_j, scaled by
// s^level (scale_dbl doublings), goes into a per-window pool that is summed at the end:
//   sum_b (b+1) S_b = sum(pool) + T_final.
template <class F>
__global__ void __launch_bounds__(128) msm_weighted_level_kernel(const XYZZ<F>* __restrict__ in, uint32_t n_in, uint32_t s, uint32_t n_out,
                                                                 uint32_t total, int scale_dbl, XYZZ<F>* t_out, XYZZ<F>* pool,
                                                                 uint32_t pool_stride, uint32_t pool_off) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    const uint32_t w = id / n_out, j = id % n_out;
    const XYZZ<F>* base = in + (size_t)w * n_in + (size_t)j * s;
    XYZZ<F> running = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    for (uint32_t i = s; i-- > 1;) {
        XYZZ<F> v = ld_vec(base + i);
        xyzz_add<F>(running, v);
        xyzz_add<F>(acc, running);
    }
    {
        XYZZ<F> v = ld_vec(base);
        xyzz_add<F>(running, v);
    }
    for (int k = 0; k < scale_dbl; k++) acc = xyzz_dbl<F>(acc);
    st_vec(t_out + (size_t)w * n_out + j, running);
    st_vec(pool + (size_t)w * pool_stride + pool_off + j, acc);
}

// K6b: one level of the pool sum: warp g of window w adds the (up to) `per` points [g*per, (g+1)*per) of the
// window's `count` points and writes out[w*groups + g]: per/32 sequential additions per lane, then a five-step
// shuffle tree.
template <class F>
__global__ void __launch_bounds__(128) msm_group_reduce_kernel(const XYZZ<F>* __restrict__ in, uint32_t stride_in, uint32_t count, uint32_t per,
                                                               uint32_t groups, uint32_t total_groups, XYZZ<F>* out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (warp >= total_groups) return;
    const uint32_t w = warp / groups, g = warp % groups;
    const uint32_t lo = g * per, hi = lo + per < count ? lo + per : count;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t k = lo + lane; k < hi; k += 32) {
        XYZZ<F> v = ld_vec(in + (size_t)w * stride_in + k);
        xyzz_add<F>(acc, v);
    }
    acc = warp_reduce_xyzz<F>(acc, lane);
    if (lane == 0) st_vec(out + warp, acc);
}

// K6' — row / column bucket reduction.  The hierarchical running-sum reduction above is a chain of ~100 dependent
// group additions and ~90 doublings (1.4 ms for 2^19 buckets with the GPU otherwise idle behind the last multiexp, and
// the dominant cost of a multi-GPU proof, where every rank still reduces all its buckets).  With the bucket index
// split as i = hi * 2^k + lo,
//     sum_i (i + 1) S_i = 2^k * sum_hi hi R_hi + sum_lo (lo + 1) C_lo,   R_hi = sum_lo S_i (row sums), C_lo = sum_hi S_i (column sums):
// two PLAIN sums over all buckets (trees of eight-way stages: the same ~2 additions per bucket, every lane busy,
// depth ~22) and two weighted sums over at most 1024 points each, which one CTA computes with a suffix scan
// (sum_i i X_i = sum_{j>=1} suffix_j).  Measured at 2^19 buckets: 1.37 -> 0.90 ms (G1), 3.71 -> 2.61 ms (G2).
// K6'a: one stage of a strided sum: out[w][g][lo] = sum_{t<f} in[w][g*f + t][lo].  Column sums: w = bucket space,
// rows of `width` = 2^k points.  Row sums: w = row, width = 1.
template <class F>
__global__ void __launch_bounds__(128) msm_colsum_kernel(const XYZZ<F>* __restrict__ in, uint32_t rows_in, uint32_t width, uint32_t f, uint32_t rows_out,
                                                         uint32_t total, XYZZ<F>* out) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    const uint32_t per_w = rows_out * width;
    const uint32_t w = id / per_w, rem = id % per_w, g = rem / width, lo = rem % width;
    const XYZZ<F>* base = in + ((size_t)w * rows_in + (size_t)g * f) * width + lo;
    XYZZ<F> acc = ld_vec(base);
    for (uint32_t t = 1; t < f; t++) {
        XYZZ<F> v = ld_vec(base + (size_t)t * width);
        xyzz_add<F>(acc, v);
    }
    st_vec(out + ((size_t)w * rows_out + g) * width + lo, acc);
}

template <class F>
static __device__ __forceinline__ XYZZ<F> shfl_down_xyzz(const XYZZ<F>& v, int d, unsigned lane, unsigned width) {
    XYZZ<F> o;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* b = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (unsigned k = 0; k < sizeof(XYZZ<F>) / 4; k++) b[k] = __shfl_down_sync(0xffffffffu, a[k], d);
    if (lane + d >= width) o = XYZZ<F>::inf();
    return o;
}
// sum over the 256 threads of a CTA (result in thread 0); sm: 8 slots
template <class F>
static __device__ __forceinline__ XYZZ<F> block_sum_xyzz_256(XYZZ<F> v, XYZZ<F>* sm, unsigned lane, unsigned warp) {
    v = warp_reduce_xyzz<F>(v, lane);
    __syncthreads();
    if (lane == 0) st_vec(sm + warp, v);
    __syncthreads();
    XYZZ<F> r = XYZZ<F>::inf();
    if (warp == 0) {
        if (lane < 8) r = ld_vec(sm + lane);
        r = warp_reduce_xyzz<F>(r, lane);
    }
    return r;
}
// K6'b: weighted sums of short arrays.  CTA a < Wr handles the row sums of bucket space a (n = nR points, may be 0),
// CTA Wr + w the column sums of space w (nC points); n is a power of two <= 1024.  Thread t owns ipt = max(1, n / 256)
// consecutive points.  With L_t = the thread's own sum_j j X_j and S_t = the sum of all points of the threads >= t
// (suffix scan), sum_i i X_i = sum_t (L_t + ipt [t >= 1] S_t) = out[0]; out[1] = U = the sum of all points, so that
// sum_i (i + 1) X_i = out[0] + out[1] (host); out[2] is unused (infinity).
template <class F>
__global__ void __launch_bounds__(256, 1) msm_small_weighted_kernel(const XYZZ<F>* __restrict__ inR, uint32_t nR, const XYZZ<F>* __restrict__ inC, uint32_t nC,
                                                                     uint32_t Wr, XYZZ<F>* out3) {
    __shared__ XYZZ<F> sm[8];
    const uint32_t a = blockIdx.x;
    const bool is_c = a >= Wr;
    const uint32_t w = is_c ? a - Wr : a;
    const uint32_t n = is_c ? nC : nR;
    const XYZZ<F>* in = is_c ? inC + (size_t)w * nC : inR + (size_t)w * nR;
    const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t ipt = n > 256 ? n / 256 : 1;
    XYZZ<F> T = XYZZ<F>::inf(), L = XYZZ<F>::inf();
    for (uint32_t j = ipt; j-- > 0;) {
        const uint32_t idx = t * ipt + j;
        if (idx < n) {
            XYZZ<F> x = ld_vec(in + idx);
            xyzz_add<F>(T, x);
        }
        if (j > 0) xyzz_add<F>(L, T);
    }
    // suffix sums of the T_t over the CTA: inside the warp by shuffles, across the warps through shared memory
    XYZZ<F> S = T;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        XYZZ<F> o = shfl_down_xyzz<F>(S, d, lane, 32);
        xyzz_add<F>(S, o);
    }
    if (lane == 0) st_vec(sm + warp, S);
    __syncthreads();
    if (warp == 0) {
        XYZZ<F> v = lane < 8 ? ld_vec(sm + lane) : XYZZ<F>::inf();
#pragma unroll 1
        for (int d = 1; d < 8; d <<= 1) {
            XYZZ<F> o = shfl_down_xyzz<F>(v, d, lane, 8);
            xyzz_add<F>(v, o);
        }
        XYZZ<F> ex = shfl_down_xyzz<F>(v, 1, lane, 8);      // sum of the warps after this one
        __syncwarp();
        if (lane < 8) st_vec(sm + lane, ex);
    }
    __syncthreads();
    {
        XYZZ<F> after = ld_vec(sm + warp);
        xyzz_add<F>(S, after);                              // S = sum_{t' >= t} T_t'
    }
    const XYZZ<F> U = S;                                    // thread 0: the sum of all points
    // sum_i i X_i = sum_t (L_t + ipt * [t >= 1] S_t): one block sum
    if (t != 0) {
        for (uint32_t m = ipt; m > 1; m >>= 1) S = xyzz_dbl<F>(S);
        xyzz_add<F>(L, S);
    }
    const XYZZ<F> Wsum = block_sum_xyzz_256<F>(L, sm, lane, warp);
    if (t == 0) {
        st_vec(out3 + (size_t)a * 3 + 0, Wsum);
        st_vec(out3 + (size_t)a * 3 + 1, U);
        st_vec(out3 + (size_t)a * 3 + 2, XYZZ<F>::inf());
    }
}

// K6'c: the same weighted sums with a shorter chain.  The array of n = a x b points (b = min(32, n) the low index bits) is
// itself reduced by rows and columns: sum_i i X_i = b * sum_hi hi R_hi + sum_lo lo C_lo with R_hi = sum_lo X, C_lo = sum_hi X.
// First kernel: one WARP per row sum and per column sum (a + b <= 64 warps per array, five shuffle steps each).  Second
// kernel: one warp per weighted sum of <= 32 points (suffix scan + reduction, ten steps).  Fifteen dependent additions
// instead of the ~27 of msm_small_weighted_kernel (which took 1.0 ms over Fq2 — the largest part of the tail behind a G2
// multiexp).  Results per array: out3[0] = sum_lo lo C_lo, out3[1] = U = sum of all points, out3[2] = sum_hi hi R_hi; the
// host adds 2^log2(b) * out3[2] (msm_finish).
template <class F>
__global__ void __launch_bounds__(128) msm_w2d_sums_kernel(const XYZZ<F>* __restrict__ inR, uint32_t nR, const XYZZ<F>* __restrict__ inC, uint32_t nC,
                                                           uint32_t Wr, XYZZ<F>* tmp /* 64 per array */) {
    const unsigned lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // global warp
    const uint32_t arr = gw >> 6, o = gw & 63;
    if (arr >= 2 * Wr) return;
    const bool is_c = arr >= Wr;
    const uint32_t w = is_c ? arr - Wr : arr;
    const uint32_t n = is_c ? nC : nR;
    if (n == 0) return;
    const XYZZ<F>* in = is_c ? inC + (size_t)w * nC : inR + (size_t)w * nR;
    const uint32_t b = n < 32 ? n : 32, a = n / b;
    if (o >= a + b) return;
    XYZZ<F> v = XYZZ<F>::inf();
    if (o < a) { if (lane < b) v = ld_vec(in + (size_t)o * b + lane); }             // row sum R_o
    else { const uint32_t lo = o - a; if (lane < a) v = ld_vec(in + (size_t)lane * b + lo); }      // column sum C_lo
    v = warp_reduce_xyzz<F>(v, lane);
    if (lane == 0) st_vec(tmp + (size_t)arr * 64 + o, v);
}
template <class F>
__global__ void __launch_bounds__(64) msm_w2d_final_kernel(const XYZZ<F>* __restrict__ tmp, uint32_t nR, uint32_t nC, uint32_t Wr, XYZZ<F>* out3) {
    const uint32_t arr = blockIdx.x;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;      // warp 0: rows (and U), warp 1: columns
    const bool is_c = arr >= Wr;
    const uint32_t n = is_c ? nC : nR;
    if (n == 0) {
        if (threadIdx.x < 3) st_vec(out3 + (size_t)arr * 3 + threadIdx.x, XYZZ<F>::inf());
        return;
    }
    const uint32_t b = n < 32 ? n : 32, a = n / b;
    const uint32_t cnt = warp == 0 ? a : b;
    const XYZZ<F>* src = tmp + (size_t)arr * 64 + (warp == 0 ? 0 : a);
    XYZZ<F> S = lane < cnt ? ld_vec(src + lane) : XYZZ<F>::inf();
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {                       // suffix sums S_j = sum_{j' >= j} Y_j'
        XYZZ<F> o = shfl_down_xyzz<F>(S, d, lane, 32);
        xyzz_add<F>(S, o);
    }
    const XYZZ<F> U = S;                                     // lane 0: the sum of all
    XYZZ<F> t = lane >= 1 ? S : XYZZ<F>::inf();              // sum_j j Y_j = sum_{j >= 1} S_j
    t = warp_reduce_xyzz<F>(t, lane);
    if (lane == 0) {
        st_vec(out3 + (size_t)arr * 3 + (warp == 0 ? 2 : 0), t);
        if (warp == 0) st_vec(out3 + (size_t)arr * 3 + 1, U);
    }
}

// Small inputs (public-input queries have a handful of points): thread per point, double-and-add,
// then a single-warp tree.  Also an independent on-device check of the bucket pipeline (tests).
template <class F>
__global__ void __launch_bounds__(128) msm_naive_kernel(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ scalars, uint32_t n,
                                                        XYZZ<F>* out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (warp * 32 >= n) return;                      // whole warps past the end (the CTA has four): `out` has ceil(n/32) slots
    const uint32_t i = warp * 32 + lane;
    XYZZ<F> acc = XYZZ<F>::inf();
    if (i < n) {
        uint32_t k[8];
#pragma unroll
        for (int j = 0; j < 8; j++) k[j] = scalars[8 * (size_t)i + j];
        Affine<F> P = ldg_vec(bases + i);
        if (!P.is_inf()) {
            int top = 255;
            while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1u)) top--;
            for (int b = top; b >= 0; b--) {
                acc = xyzz_dbl<F>(acc);
                if ((k[b >> 5] >> (b & 31)) & 1u) xyzz_madd<F>(acc, P.x, P.y, false);
            }
        }
    }
    acc = warp_reduce_xyzz<F>(acc, lane);
    if (lane == 0) st_vec(out + warp, acc);
}

// any base at infinity paired with a non-zero scalar?  (bellman: SynthesisError::UnexpectedIdentity)
template <class F>
__global__ void msm_identity_check_kernel(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ scalars, size_t n, uint32_t* flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) o |= scalars[8 * i + j];
    if (o == 0) return;
    Affine<F> P = ldg_vec(bases + i);
    if (P.is_inf()) atomicOr(flag, 1u);
}

// canonical LE affine (host interchange) -> Montgomery affine; infinity (all zero) stays all zero.
// flags[0] |= 1 if a coordinate is not canonical, |= 2 if a point is off the curve, |= 4 if any infinity
template <class F, int NF>
__global__ void bases_import_kernel(Affine<F>* pts, size_t n, const F b_coeff, uint32_t* flags) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = ld_vec(pts + i);
    Fq* c = reinterpret_cast<Fq*>(&p);
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < NF; k++) {
#pragma unroll
        for (int j = 0; j < 8; j++) any |= c[k].v[j];
    }
    if (any == 0) { atomicOr(flags, 4u); return; }
#pragma unroll
    for (int k = 0; k < NF; k++) {
        if (!fp_is_canonical<FqParams>(c[k].v)) atomicOr(flags, 1u);
        c[k] = fp_to_mont<FqParams>(c[k]);
    }
    if (!affine_on_curve<F>(p, b_coeff)) atomicOr(flags, 2u);
    st_vec(pts + i, p);
}

// Montgomery -> canonical for a handful of XYZZ points
template <class F, int NF>
__global__ void xyzz_export_kernel(XYZZ<F>* pts, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ<F> p = ld_vec(pts + i);
    Fq* c = reinterpret_cast<Fq*>(&p);
#pragma unroll
    for (int k = 0; k < NF; k++) c[k] = fp_from_mont<FqParams>(c[k]);
    st_vec(pts + i, p);
}

// ---- synthetic bases: out[i] = (first + i) * G, distinct points with known
// discrete logarithms so that a full-size MSM can be checked with one scalar multiplication.
template <class F>
__global__ void __launch_bounds__(128) bases_gen_chain_kernel(XYZZ<F>* out, size_t n, uint64_t first, const Affine<F> G, int chunk) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t start = t * (size_t)chunk;
    if (start >= n) return;
    uint64_t k = first + start;
    XYZZ<F> P = XYZZ<F>::inf();
    for (int b = 63 - __clzll((long long)k); b >= 0; b--) {
        P = xyzz_dbl<F>(P);
        if ((k >> b) & 1ull) xyzz_madd<F>(P, G.x, G.y, false);
    }
    size_t end = start + chunk < n ? start + chunk : n;
    for (size_t i = start; i < end; i++) {
        st_vec(out + i, P);
        xyzz_madd<F>(P, G.x, G.y, false);
    }
}
#define GEN_CHUNK 32
template <class F>
__global__ void __launch_bounds__(128) bases_gen_normalise_kernel(const XYZZ<F>* in, Affine<F>* out, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t start = t * (size_t)GEN_CHUNK;
    if (start >= n) return;
    const int cnt = (int)(start + GEN_CHUNK < n ? GEN_CHUNK : n - start);
    F pre[GEN_CHUNK];                     // running products of ZZ*ZZZ (Montgomery's simultaneous inversion)
    F run = F::one();
    for (int j = 0; j < cnt; j++) {
        XYZZ<F> p = ld_vec(in + start + j);
        pre[j] = run;
        if (!p.is_inf()) run = run * (p.ZZ * p.ZZZ);
    }
    F iv = inv(run);
    for (int j = cnt - 1; j >= 0; j--) {
        XYZZ<F> p = ld_vec(in + start + j);
        if (p.is_inf()) { st_vec(out + start + j, Affine<F>::inf()); continue; }
        F tj = iv * pre[j];               // 1 / (ZZ_j ZZZ_j)
        iv = iv * (p.ZZ * p.ZZZ);
        Affine<F> a;
        a.x = p.X * (tj * p.ZZZ);
        a.y = p.Y * (tj * p.ZZ);
        st_vec(out + start + j, a);
    }
}

// Fixed-base table (Groth16 queries never change): out[i*W + w] = 2^(c w) * P_i, so that all windows of a
// multiexp share ONE bucket space and the window size can grow (c = 20: 13 windows instead of 16).
template <class F>
__global__ void __launch_bounds__(128) bases_table_kernel(const Affine<F>* __restrict__ pts, size_t n, int c, int W, XYZZ<F>* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> a = ldg_vec(pts + i);
    XYZZ<F> P = XYZZ<F>::from_affine(a);
    for (int w = 0; w < W; w++) {
        st_vec(out + i * (size_t)W + w, P);
        if (w + 1 < W) for (int k = 0; k < c; k++) P = xyzz_dbl<F>(P);
    }
}

// Integer-pipe peak probe: dependency-free mad.lo.u32 chains, 8 independent accumulators per thread.
__global__ void __launch_bounds__(256) imad_peak_kernel(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x0) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x1) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x2) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x3) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x4) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x5) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x6) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x7) : "r"(a), "r"(b));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

// ------------------------------------------------------------------------------------------ host
template <class F> struct GroupInfo;
template <> struct GroupInfo<Fq> { static constexpr int NF_AFF = 2, NF_XYZZ = 4; };
template <> struct GroupInfo<Fq2> { static constexpr int NF_AFF = 4, NF_XYZZ = 8; };

Fq host_g1_b() { return fp_from_u64<FqParams>(3); }
Fq2 host_g2_b() {
    // 3 / (9 + u)
    Fq2 xi; xi.c0 = fp_from_u64<FqParams>(9); xi.c1 = fp_from_u64<FqParams>(1);
    Fq2 three; three.c0 = fp_from_u64<FqParams>(3); three.c1 = Fq::zero();
    return three * inv(xi);
}

static inline unsigned nblk(size_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

int msm_window_bits(size_t n) {
    // ZA_MSM_C overrides the choice (tuning / tests of other window sizes)
    if (const char* e = getenv("ZA_MSM_C")) { int c = atoi(e); if (c >= 2 && c <= 16) return c; }
    // minimise n*W*10 + W*2^(c-1)*28 field multiplications, with W = ceil(255/c); c <= 16 keeps the
    // bucket array (W * 2^(c-1) XYZZ points) and the reduction latency small.
    int best = 1;
    double best_cost = 1e300;
    for (int c = 2; c <= 16; c++) {
        double W = (255 + c - 1) / c;
        double cost = (double)n * W * 10.0 + W * (double)(1u << (c - 1)) * 40.0;
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

template <class F>
static void import_bases(Ctx* ctx, Affine<F>* d_pts, size_t n, uint32_t* d_flags);
template <>
void import_bases<Fq>(Ctx* ctx, Affine<Fq>* d_pts, size_t n, uint32_t* d_flags) {
    bases_import_kernel<Fq, 2><<<nblk(n, 128), 128, 0, ctx->stream>>>(d_pts, n, host_g1_b(), d_flags);
    ctx->launches++;
}
template <>
void import_bases<Fq2>(Ctx* ctx, Affine<Fq2>* d_pts, size_t n, uint32_t* d_flags) {
    bases_import_kernel<Fq2, 4><<<nblk(n, 128), 128, 0, ctx->stream>>>(d_pts, n, host_g2_b(), d_flags);
    ctx->launches++;
}

// Upload canonical affine points and convert in place.  Returns flags (see bases_import_kernel).
template <class F>
uint32_t bases_import(Ctx* ctx, void* d_pts, size_t n) {
    if (!n) return 0;
    DevBuf flags(4);
    ZA_CUDA(cudaMemsetAsync(flags.p, 0, 4, ctx->stream));
    import_bases<F>(ctx, (Affine<F>*)d_pts, n, flags.as<uint32_t>());
    ZA_CUDA(cudaGetLastError());
    uint32_t h = 0;
    ZA_CUDA(cudaMemcpyAsync(&h, flags.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ZA_CUDA(cudaStreamSynchronize(ctx->stream));
    return h;
}
template uint32_t bases_import<Fq>(Ctx*, void*, size_t);
template uint32_t bases_import<Fq2>(Ctx*, void*, size_t);

// Enqueue one multiexp into `slot` (0..7).  Sort + accumulate go to the context's stream, the bucket reduction
// and the read-back of the W window sums to the side stream, so the next multiexp's accumulation overlaps it.
// share_sort >= 0: reuse the digit sort of that slot (same scalar vector, e.g. the G1 and G2 B queries).
// msm_finish waits for the slot and does the window combination on the host.
template <class F>
void msm_enqueue(Ctx* ctx, int slot_id, const Affine<F>* d_bases, const uint32_t* d_scalars, size_t n, bool has_infinity, int share_sort,
                 const Affine<F>* d_table, int tab_c, int tab_W, const MsmPart* more, int n_more) {
    MsmSlot& sl = ctx->slots[slot_id];
    if (n_more < 0 || n_more > 3 || (n_more && (!more || !d_table || sizeof(F) != sizeof(Fq) || n <= 64)))
        throw ZaError(ZA_ERR_INVALID, "internal: several queries in one multiexp need G1 fixed-base tables");
    const int nparts = 1 + n_more;
    if (!sl.side) {
        int lo_prio = 0, hi_prio = 0;
        ZA_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        ZA_CUDA(cudaStreamCreateWithPriority(&sl.side, cudaStreamNonBlocking, hi_prio));
    }
    cudaStream_t st = ctx->stream, side = sl.side;
    if (sl.busy) throw ZaError(ZA_ERR_INVALID, "msm slot enqueued twice without msm_finish");
    sl.kind = 0; sl.n = n; sl.nparts = nparts; sl.single = false;
    if (n == 0) { sl.busy = true; return; }
    // this slot's buffers (and its sort, if another slot borrowed it) may still be read by side-stream work
    if (sl.done_valid) ZA_CUDA(cudaStreamWaitEvent(st, sl.done, 0));
    for (int j = 0; j < 8; j++)
        if ((sl.sort_users >> j) & 1u) { if (ctx->slots[j].done_valid) ZA_CUDA(cudaStreamWaitEvent(st, ctx->slots[j].done, 0)); }
    sl.sort_users = 0;
    if (n >= ((size_t)1 << 27)) throw ZaError(ZA_ERR_INVALID, "multiexp of 2^27 or more points is not supported");
    if (!sl.done) {
        ZA_CUDA(cudaEventCreateWithFlags(&sl.acc_done, cudaEventDisableTiming));
        ZA_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    static const bool timeline = getenv("ZA_DEBUG_TIMELINE") != nullptr;
    if (timeline) {
        if (!sl.dbg_start) { cudaEventCreate(&sl.dbg_start); cudaEventCreate(&sl.dbg_sort); cudaEventCreate(&sl.dbg_acc); cudaEventCreate(&sl.dbg_done); }
        cudaEventRecord(sl.dbg_start, st);
    }
    const size_t want_host = 64 + 6 * 128 * sizeof(XYZZ<Fq2>); // header (entry count) + up to 128 bucket spaces x 6 partial results
    if (sl.host_win_bytes < want_host) {
        if (sl.host_win) cudaFreeHost(sl.host_win);
        ZA_CUDA(cudaHostAlloc(&sl.host_win, want_host, cudaHostAllocDefault));
        sl.host_win_bytes = want_host;
    }
    if (has_infinity) {
        DevBuf& flag = ctx->scratch[11];
        flag.ensure(256);
        ZA_CUDA(cudaMemsetAsync(flag.p, 0, 4, st));
        msm_identity_check_kernel<F><<<nblk(n, 256), 256, 0, st>>>(d_bases, d_scalars, n, flag.as<uint32_t>());
        ctx->launches++;
        uint32_t h = 0;
        ZA_CUDA(cudaMemcpyAsync(&h, flag.p, 4, cudaMemcpyDeviceToHost, st));
        ZA_CUDA(cudaStreamSynchronize(st));
        if (h) throw ZaError(ZA_ERR_UNEXPECTED_IDENTITY, "multiexp: a base at infinity has a non-zero exponent");
    }
    if (n <= 64) {
        sl.kind = 1;
        sl.warps = (int)((n + 31) / 32);
        sl.segs.ensure(2 * sizeof(XYZZ<F>));
        msm_naive_kernel<F><<<nblk((size_t)sl.warps * 32, 128), 128, 0, st>>>(d_bases, d_scalars, (uint32_t)n, sl.segs.as<XYZZ<F>>());
        ctx->launches++;
        ZA_CUDA(cudaGetLastError());
        ZA_CUDA(cudaMemcpyAsync((uint8_t*)sl.host_win + 64, sl.segs.p, (size_t)sl.warps * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, st));
        ZA_CUDA(cudaEventRecord(sl.done, st));
        sl.done_valid = true;
        sl.busy = true;
        return;
    }
    // d_table != nullptr: fixed-base table of this query range (entry i*W + w = 2^(c w) P_i): one bucket space
    const bool single = d_table != nullptr;
    const int c = single ? tab_c : msm_window_bits(n);
    const int W = single ? tab_W : (255 + c - 1) / c;
    const uint32_t B = 1u << (c - 1);
    const int Wr = single ? nparts : W;             // bucket spaces to reduce
    const uint32_t nkeys = (uint32_t)Wr * B;
    uint64_t Emax = (uint64_t)n * W;
    for (int p = 0; p < n_more; p++) {
        if (more[p].n >= ((size_t)1 << 27)) throw ZaError(ZA_ERR_INVALID, "multiexp of 2^27 or more points is not supported");
        Emax += (uint64_t)more[p].n * W;
    }
    sl.single = single;
    if (single) d_bases = d_table;
    if (Emax >= 0xffffffffull) throw ZaError(ZA_ERR_INVALID, "multiexp too large for 32-bit entry offsets");
    // batched-affine pair rounds before the XYZZ accumulation: worth it while buckets hold several points
    // (load = entries per bucket).  ZA_MSM_ROUNDS overrides (0 = XYZZ accumulation only).
    constexpr int PAIR_NT = 128, PAIR_LP = 32, PAIR_MAX_ROUNDS = 6;
    int rounds = 0;
    {
        // Measured on B200 (profiles/r01_pair_rounds.md): over Fq2 the rounds win (17 instead of 28 Fq products
        // per addition, -19 % at 2^20 points); over Fq the first round has to gather every 64-byte point twice
        // from a table far larger than L2 (the memory system fetches 128 bytes per miss) and is DRAM bound,
        // so G1 keeps the single-pass XYZZ accumulation.
        const double load = (double)Emax / (double)nkeys;
        if (sizeof(F) == sizeof(Fq2)) rounds = load >= 12 ? 4 : load >= 6 ? 3 : load >= 3 ? 1 : 0;
        // a round is one full pass of launches (rescan + round kernel); below ~1.5 waves of its CTAs it costs more than
        // the products it saves.  One rank's share of a 2^20 proof, stage 2 alone on a B200 (profiles/r02_shards.md):
        // 1/8: 4 rounds 5.35 ms, 3 rounds 4.98, 2 rounds 5.1, none 5.4; 1/4: 7.97 / 7.55 / 7.62; 1/2: 13.47 / 13.17;
        // the whole multiexp (13.6 M entries) takes the same time with 3 or 4 rounds.
        if (rounds > 3) rounds = 3;
        if (Emax < (1u << 16)) rounds = 0;
        if (const char* e = getenv("ZA_MSM_ROUNDS")) { int v = atoi(e); if (v >= 0 && v <= PAIR_MAX_ROUNDS) rounds = v; }
        if (nparts > 1) rounds = 0;
    }
    sl.rounds = rounds;
    int pair_lp = PAIR_LP;
    if (const char* e = getenv("ZA_MSM_PAIR_LP")) { if (atoi(e) == 16) pair_lp = 16; }
    uint64_t pair_cap[PAIR_MAX_ROUNDS + 1];               // upper bounds of the point count after each round
    {
        uint64_t cap = Emax;
        for (int r = 0; r < rounds; r++) { cap = (cap + (cap < nkeys ? cap : nkeys) + 1) / 2; pair_cap[r] = cap; }
        pair_cap[rounds] = cap;
    }
    const uint64_t Efinal = rounds ? pair_cap[rounds - 1] : Emax;
    // chunk length: ~2 chunks per resident thread slot, between 16 and 2048 entries (measured best of 2/4/8 at 2^20)
    uint64_t per_slot = 2;
    if (const char* e = getenv("ZA_MSM_CHUNKS_PER_SLOT")) { int v = atoi(e); if (v >= 1 && v <= 64) per_slot = (uint64_t)v; }
    uint64_t tslots = (uint64_t)ctx->sm_count * 512 * per_slot;
    uint32_t Lc = (uint32_t)((Efinal + tslots - 1) / tslots);
    if (Lc < 16) Lc = 16;
    if (Lc > 2048) Lc = 2048;
    const uint32_t nchunks = (uint32_t)((Efinal + Lc - 1) / Lc);
    sl.kind = 2; sl.c = c; sl.W = Wr; sl.nkeys = nkeys; sl.Lc = Lc; sl.nchunks = nchunks;
    sl.acc_cat = sizeof(F) == sizeof(Fq) ? PROF_ACC_G1 : PROF_ACC_G2;

    const uint32_t nctas = (nkeys + SCAN_PER_CTA - 1) / SCAN_PER_CTA;
    uint32_t *d_counts = nullptr, *d_offsets, *d_cursors = nullptr, *d_cta = nullptr;
    uint32_t* d_entries;
    if (share_sort >= 0) {
        const MsmSlot& src = ctx->slots[share_sort];
        // the borrowed sort may cover several parts: this multiexp runs over its FIRST part (keys [0, B), entries from 0)
        if (src.kind != 2 || src.n != n || src.c != c || src.single != single || (!single && src.W != Wr) || nparts != 1)
            throw ZaError(ZA_ERR_INVALID, "msm sort sharing needs an identical scalar vector and window layout");
        d_offsets = src.d_offsets;
        d_entries = src.d_entries;
        ctx->slots[share_sort].sort_users |= 1u << slot_id;
    } else {
        sl.counts.ensure(((size_t)nkeys * 3 + 4 + nctas) * 4);   // counts | offsets (+total) | cursors | cta totals
        d_counts = sl.counts.as<uint32_t>();
        d_offsets = d_counts + nkeys;
        d_cursors = d_offsets + nkeys + 1;
        d_cta = d_cursors + nkeys;
        sl.entries.ensure((size_t)Emax * 4);
        d_entries = sl.entries.as<uint32_t>();
    }
    sl.d_offsets = d_offsets; sl.d_entries = d_entries;
    uint32_t *d_pair_off = nullptr, *d_pair_cta = nullptr;
    const uint32_t* d_offsets_final = d_offsets;
    if (rounds) {
        sl.pair_offs.ensure(((size_t)rounds * (nkeys + 1) + nctas) * 4);
        d_pair_off = sl.pair_offs.as<uint32_t>();
        d_pair_cta = d_pair_off + (size_t)rounds * (nkeys + 1);
        sl.pair_pts[0].ensure((size_t)pair_cap[0] * sizeof(Affine<F>));
        if (rounds > 1) sl.pair_pts[1].ensure((size_t)pair_cap[1] * sizeof(Affine<F>));
    }
    sl.buckets.ensure((size_t)nkeys * sizeof(XYZZ<F>));
    sl.parts.ensure((size_t)nchunks * (2 * sizeof(XYZZ<F>) + 8) + 16);   // part_head | part_tail | owner | long_list | long_count
    XYZZ<F>* d_head = sl.parts.as<XYZZ<F>>();
    XYZZ<F>* d_tail = d_head + nchunks;
    uint32_t* d_owner = reinterpret_cast<uint32_t*>(d_tail + nchunks);
    uint32_t* d_long_list = d_owner + nchunks;
    uint32_t* d_long_count = d_long_list + nchunks;
    // weighted bucket reduction: levels of segment size 16 (B is a power of two)
    // segment size 2^SEG_LOG: the reduction is latency bound (each level is a chain of 2(s-1) dependent group
    // additions per thread), so small segments and more levels win: 4 measured best of 4/8/16
    int SEG_LOG = 2;
    if (const char* e = getenv("ZA_MSM_SEG_LOG")) { int v = atoi(e); if (v >= 1 && v <= 5) SEG_LOG = v; }
    const uint32_t SEG = 1u << SEG_LOG;
    struct Level { uint32_t n_in, s, n_out, pool_off; };
    std::vector<Level> levels;
    uint32_t pool_len = 0;
    for (uint32_t nin = B; nin > 1;) {
        uint32_t sgm = nin < SEG ? nin : SEG;
        Level lv{nin, sgm, nin / sgm, pool_len};
        pool_len += lv.n_out;
        levels.push_back(lv);
        nin = lv.n_out;
    }
    const uint32_t pool_used = pool_len + 1;                         // + the final plain sum T
    const uint32_t pool_stride = (pool_used + 63) / 64 * 64;         // padded with points at infinity
    // pool sum: a tree of warp reductions over GRP_PER points each (4 sequential additions per lane, then a
    // five-step shuffle tree), sizes pool_stride -> /128 -> ... -> 1 per window
    constexpr uint32_t GRP_PER = 128;
    std::vector<uint32_t> grp_sizes;
    size_t grp_total = 0;
    for (uint32_t cnt = pool_stride; cnt > 1;) { cnt = (cnt + GRP_PER - 1) / GRP_PER; grp_sizes.push_back(cnt); grp_total += cnt; }
    if (grp_sizes.empty()) { grp_sizes.push_back(1); grp_total = 1; }
    const size_t lvl_elems = (size_t)Wr * (B / (B < SEG ? B : SEG) + 1);
    sl.segs.ensure(((size_t)Wr * pool_stride + 2 * lvl_elems + (size_t)Wr * grp_total) * sizeof(XYZZ<F>));
    XYZZ<F>* d_pool = sl.segs.as<XYZZ<F>>();
    XYZZ<F>* d_lvl[2] = {d_pool + (size_t)Wr * pool_stride, d_pool + (size_t)Wr * pool_stride + lvl_elems};
    XYZZ<F>* d_grp = d_lvl[1] + lvl_elems;
    XYZZ<F>* d_win = d_grp + (size_t)Wr * (grp_total - 1);          // the last level has one point per window
    size_t win_points = (size_t)Wr;
    XYZZ<F>* d_buckets = sl.buckets.as<XYZZ<F>>();

    ZA_CUDA(cudaMemsetAsync(d_buckets, 0, (size_t)nkeys * sizeof(XYZZ<F>), st));
    ZA_CUDA(cudaMemsetAsync(d_owner, 0xff, (size_t)nchunks * 4, st));
    ZA_CUDA(cudaMemsetAsync(d_long_count, 0, 4, st));
    // bucket reduction: row / column scheme (K6') unless the bucket space is larger than 2^20 or ZA_MSM_REDUCE_OLD is set
    static const bool reduce_old = getenv("ZA_MSM_REDUCE_OLD") != nullptr;
    const int nb = c - 1;
    const bool use_rc = !reduce_old && nb <= 20;
    sl.red_mode = use_rc ? 1 : 0;
    if (!use_rc) ZA_CUDA(cudaMemsetAsync(d_pool, 0, (size_t)Wr * pool_stride * sizeof(XYZZ<F>), st));
    if (share_sort < 0) {
        ZA_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)nkeys * 4, st));
        ProfScope prof(ctx, PROF_MSM_SORT, (double)n);
        msm_digits_hist_kernel<<<nblk(n, 256), 256, 0, st>>>(d_scalars, n, c, W, B, d_counts, single ? 1 : 0, 0u);
        for (int p = 0; p < n_more; p++)
            if (more[p].n) msm_digits_hist_kernel<<<nblk(more[p].n, 256), 256, 0, st>>>(more[p].scalars, more[p].n, c, W, B, d_counts, 1, (uint32_t)(p + 1) * B);
        msm_scan_totals_kernel<<<nctas, 1024, 0, st>>>(d_counts, nkeys, d_cta, 0);
        msm_scan_ctas_kernel<<<1, 1024, 0, st>>>(d_cta, nctas, d_offsets + nkeys);
        msm_scan_apply_kernel<<<nctas, 1024, 0, st>>>(d_counts, nkeys, d_cta, d_offsets, d_cursors, 0);
        msm_digits_scatter_kernel<<<nblk(n, 256), 256, 0, st>>>(d_scalars, n, c, W, B, d_cursors, d_entries, single ? 1 : 0, 0u);
        for (int p = 0; p < n_more; p++)
            if (more[p].n) msm_digits_scatter_kernel<<<nblk(more[p].n, 256), 256, 0, st>>>(more[p].scalars, more[p].n, c, W, B, d_cursors, d_entries, 1, (uint32_t)(p + 1) * B);
        ctx->launches += 5 + 2 * n_more;
        if (!sl.sort_ev) ZA_CUDA(cudaEventCreateWithFlags(&sl.sort_ev, cudaEventDisableTiming));
        ZA_CUDA(cudaEventRecord(sl.sort_ev, st));          // a multiexp that shares this sort from another stream waits for it
    } else if (ctx->slots[share_sort].sort_ev) {
        ZA_CUDA(cudaStreamWaitEvent(st, ctx->slots[share_sort].sort_ev, 0));
    }
    if (timeline) cudaEventRecord(sl.dbg_sort, st);
    {
        ProfScope prof(ctx, sl.acc_cat, 0);
        const Affine<F>* cur_pts = d_bases;
        const uint32_t* cur_off = d_offsets;
        uint64_t cap = Emax;
        for (int r = 0; r < rounds; r++) {
            uint32_t* nxt_off = d_pair_off + (size_t)r * (nkeys + 1);
            Affine<F>* nxt_pts = sl.pair_pts[r & 1].as<Affine<F>>();
            msm_scan_totals_kernel<<<nctas, 1024, 0, st>>>(cur_off, nkeys, d_pair_cta, 1);
            msm_scan_ctas_kernel<<<1, 1024, 0, st>>>(d_pair_cta, nctas, nxt_off + nkeys);
            msm_scan_apply_kernel<<<nctas, 1024, 0, st>>>(cur_off, nkeys, d_pair_cta, nxt_off, nullptr, 1);
            cap = pair_cap[r];
            bool lp_done = false;
            if constexpr (sizeof(F) == sizeof(Fq2)) {
                // Fq2 split over lane pairs (K5a'): 64 pairs per CTA, four CTAs per SM; LP outputs per pair, shorter runs
                // when the round is small so that the grid still covers the SMs a few times
                // MEASURED (profiles/r02_g2_lanepair.md): 8.48 ms against 7.75 ms for the single-thread rounds at 2^20 — the
                // rounds are bound by their gathers, the per-warp inversion (19 % of the samples at 512 outputs per warp)
                // and bookkeeping, not by occupancy, and the pair pays 4 wide products per Fq2 product where one thread
                // pays 3.  Kept selectable (ZA_G2_LANEPAIR=1), off by default.
                static const bool lanepair = getenv("ZA_G2_LANEPAIR") && atoi(getenv("ZA_G2_LANEPAIR")) != 0;
                if (lanepair) {
                    const uint64_t wave32 = (uint64_t)ctx->sm_count * 4 * (LPK_NT / 2) * 32;
                    static const int lp_max = getenv("ZA_G2_LP_MAX") ? atoi(getenv("ZA_G2_LP_MAX")) : 32;
                    int lp = cap >= 6 * wave32 ? 64 : cap >= 3 * wave32 ? 32 : cap >= wave32 ? 16 : 8;
                    if (lp > lp_max) lp = lp_max;
                    const unsigned grid = nblk(cap, (LPK_NT / 2) * lp);
                    const Affine<Fq2>* in_pts = cur_pts;
                    Affine<Fq2>* out_pts = nxt_pts;
#define ZA_LPK_LAUNCH(FIRST_, LP_)                                                                                                         \
    do {                                                                                                                                   \
        ZA_CUDA(cudaFuncSetAttribute(msm_pair_round_g2lp_kernel<FIRST_, LP_>, cudaFuncAttributeMaxDynamicSharedMemorySize, LPK_SM_BYTES)); \
        msm_pair_round_g2lp_kernel<FIRST_, LP_><<<grid, LPK_NT, LPK_SM_BYTES, st>>>(in_pts, r == 0 ? d_entries : nullptr, cur_off, nxt_off, nkeys, out_pts); \
    } while (0)
                    if (r == 0) { if (lp == 64) ZA_LPK_LAUNCH(true, 64); else if (lp == 32) ZA_LPK_LAUNCH(true, 32); else if (lp == 16) ZA_LPK_LAUNCH(true, 16); else ZA_LPK_LAUNCH(true, 8); }
                    else { if (lp == 64) ZA_LPK_LAUNCH(false, 64); else if (lp == 32) ZA_LPK_LAUNCH(false, 32); else if (lp == 16) ZA_LPK_LAUNCH(false, 16); else ZA_LPK_LAUNCH(false, 8); }
#undef ZA_LPK_LAUNCH
                    lp_done = true;
                }
            }
            if constexpr (sizeof(F) == sizeof(Fq2)) {
                // single-thread rounds with the working set in shared memory (K5a''): 16 warps per SM
                // MEASURED (profiles/r02_g2_lanepair.md): G2 accumulation at 2^20 7.60 -> 7.04 ms, same results; ZA_G2_SM=0 restores
                // the register-resident rounds
                static const bool g2sm = !(getenv("ZA_G2_SM") && atoi(getenv("ZA_G2_SM")) == 0);
                if (g2sm && !lp_done) {
                    const bool small = pair_lp == 16 || cap < (uint64_t)ctx->sm_count * 4 * G2SM_NT * 32;
                    const unsigned grid = nblk(cap, G2SM_NT * (small ? 16 : 32));
                    const Affine<Fq2>* in_pts = cur_pts;
                    Affine<Fq2>* out_pts = nxt_pts;
                    if (r == 0) { if (small) msm_pair_round_g2sm_kernel<true, 16><<<grid, G2SM_NT, G2SM_BYTES, st>>>(in_pts, d_entries, cur_off, nxt_off, nkeys, out_pts);
                                  else msm_pair_round_g2sm_kernel<true, 32><<<grid, G2SM_NT, G2SM_BYTES, st>>>(in_pts, d_entries, cur_off, nxt_off, nkeys, out_pts); }
                    else { if (small) msm_pair_round_g2sm_kernel<false, 16><<<grid, G2SM_NT, G2SM_BYTES, st>>>(in_pts, nullptr, cur_off, nxt_off, nkeys, out_pts);
                           else msm_pair_round_g2sm_kernel<false, 32><<<grid, G2SM_NT, G2SM_BYTES, st>>>(in_pts, nullptr, cur_off, nxt_off, nkeys, out_pts); }
                    lp_done = true;
                }
            }
            // small rounds: shorter per-thread runs so that the grid still fills the SMs
            if (lp_done) {
            } else if (pair_lp == 16 || cap < (uint64_t)ctx->sm_count * 4 * PAIR_NT * 32) {
                const unsigned grid = nblk(cap, PAIR_NT * 16);
                if (r == 0) msm_pair_round_kernel<F, true, PAIR_NT, 16><<<grid, PAIR_NT, 0, st>>>(cur_pts, d_entries, cur_off, nxt_off, nkeys, nxt_pts);
                else msm_pair_round_kernel<F, false, PAIR_NT, 16><<<grid, PAIR_NT, 0, st>>>(cur_pts, nullptr, cur_off, nxt_off, nkeys, nxt_pts);
            } else {
                const unsigned grid = nblk(cap, PAIR_NT * PAIR_LP);
                if (r == 0) msm_pair_round_kernel<F, true, PAIR_NT, PAIR_LP><<<grid, PAIR_NT, 0, st>>>(cur_pts, d_entries, cur_off, nxt_off, nkeys, nxt_pts);
                else msm_pair_round_kernel<F, false, PAIR_NT, PAIR_LP><<<grid, PAIR_NT, 0, st>>>(cur_pts, nullptr, cur_off, nxt_off, nkeys, nxt_pts);
            }
            ctx->launches += 4;
            if (ctx->profile) ZA_CUDA(cudaMemcpyAsync((uint32_t*)sl.host_win + 1 + r, nxt_off + nkeys, 4, cudaMemcpyDeviceToHost, st));
            cur_pts = nxt_pts;
            cur_off = nxt_off;
        }
        // G1: working set in shared memory (msm_accumulate_g1_sm_kernel); ZA_MSM_ACC_SM=0 selects the register version
        static const bool acc_sm = !(getenv("ZA_MSM_ACC_SM") && atoi(getenv("ZA_MSM_ACC_SM")) == 0);
        bool launched = false;
        if constexpr (sizeof(F) == sizeof(Fq)) {
            if (acc_sm || nparts > 1) {
                AccTabs tabs;
                tabs.tab[0] = rounds ? cur_pts : d_bases;
                for (int p = 0; p < 3; p++) tabs.tab[p + 1] = p < n_more ? reinterpret_cast<const Affine<Fq>*>(more[p].table) : nullptr;
                tabs.nparts = (uint32_t)nparts;
                tabs.nb = nparts > 1 ? (uint32_t)(c - 1) : 31u;
                if (rounds) {
                    ZA_CUDA(cudaFuncSetAttribute(msm_accumulate_g1_sm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACC_SM_BYTES));
                    msm_accumulate_g1_sm_kernel<true><<<nblk(nchunks, ACC_SM_NT), ACC_SM_NT, ACC_SM_BYTES, st>>>(tabs, nullptr, cur_off, nkeys, Lc, d_buckets, d_head, d_tail, d_owner);
                } else {
                    ZA_CUDA(cudaFuncSetAttribute(msm_accumulate_g1_sm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACC_SM_BYTES));
                    msm_accumulate_g1_sm_kernel<false><<<nblk(nchunks, ACC_SM_NT), ACC_SM_NT, ACC_SM_BYTES, st>>>(tabs, d_entries, d_offsets, nkeys, Lc, d_buckets, d_head, d_tail, d_owner);
                }
                launched = true;
            }
        }
        if (!launched) {
            if (rounds) msm_accumulate_kernel<F, true><<<nblk(nchunks, 128), 128, 0, st>>>(cur_pts, nullptr, cur_off, nkeys, Lc, d_buckets, d_head, d_tail, d_owner);
            else msm_accumulate_kernel<F, false><<<nblk(nchunks, 128), 128, 0, st>>>(d_bases, d_entries, d_offsets, nkeys, Lc, d_buckets, d_head, d_tail, d_owner);
        }
        ctx->launches++;
        d_offsets_final = cur_off;
    }
    ZA_CUDA(cudaEventRecord(sl.acc_done, st));
    if (timeline) cudaEventRecord(sl.dbg_acc, st);
    ZA_CUDA(cudaStreamWaitEvent(side, sl.acc_done, 0));
    {
        ProfScope prof(ctx, PROF_MSM_REDUCE, (double)nkeys, side, true);
        msm_fixup_kernel<F><<<nblk(nchunks, 128), 128, 0, side>>>(d_offsets_final, nkeys, Lc, nchunks, d_head, d_tail, d_owner, d_buckets, d_long_list,
                                                                  d_long_count);
        msm_fixup_long_kernel<F><<<ctx->sm_count * 2, 128, 0, side>>>(d_offsets_final, Lc, d_head, d_tail, d_owner, d_buckets, d_long_list, d_long_count);
        ctx->launches += 2;
        if (use_rc) {
            const uint32_t k = nb < 10 ? (uint32_t)nb : 10u, Lw = 1u << k, H = B >> k;      // B = H rows x Lw columns
            // fan-in of a tree stage: a thread adds `fan` points one after the other (fan - 1 dependent additions, ~9 us each
            // over Fq, ~25 us over Fq2 when nothing hides the latency); a smaller fan-in is a shorter chain and more launches
            static const uint32_t fan_env = getenv("ZA_MSM_RED_FANIN") ? (uint32_t)atoi(getenv("ZA_MSM_RED_FANIN")) : 0u;
            const uint32_t fan = (fan_env == 2 || fan_env == 4 || fan_env == 8) ? fan_env : 4u;      // measured 8 / 4 / 2: G2 reduction 1.98 / 1.89 / 1.85 ms, G1 0.77 / 0.71 / 0.71
            const size_t stage_elems = (size_t)Wr * (B / fan > Lw ? B / fan : Lw);
            const size_t r_elems = H > 1 ? (size_t)Wr * H : 0;
            sl.red.ensure((4 * stage_elems + r_elems + (size_t)6 * Wr + (size_t)128 * Wr) * sizeof(XYZZ<F>));      // ... + 64 partial sums per array (K6'c)
            XYZZ<F>* d_stage[2] = {sl.red.as<XYZZ<F>>(), sl.red.as<XYZZ<F>>() + stage_elems};
            XYZZ<F>* d_rstage[2] = {d_stage[1] + stage_elems, d_stage[1] + 2 * stage_elems};
            XYZZ<F>* d_r = d_rstage[1] + stage_elems;
            XYZZ<F>* d_out3 = d_r + r_elems;
            const XYZZ<F>* d_c = d_buckets;
            if (H > 1) {
                // row sums: every row (Lw contiguous buckets) is a batch of Lw one-point "rows" for the same kernel.  They
                // run on the slot's second side stream, next to the column sums (both chains are latency bound)
                if (!sl.side2) {
                    int lo_prio = 0, hi_prio = 0;
                    ZA_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
                    ZA_CUDA(cudaStreamCreateWithPriority(&sl.side2, cudaStreamNonBlocking, hi_prio));
                    ZA_CUDA(cudaEventCreateWithFlags(&sl.red_fork, cudaEventDisableTiming));
                    ZA_CUDA(cudaEventCreateWithFlags(&sl.red_join, cudaEventDisableTiming));
                }
                ZA_CUDA(cudaEventRecord(sl.red_fork, side));
                ZA_CUDA(cudaStreamWaitEvent(sl.side2, sl.red_fork, 0));
                {
                    cudaStream_t side = sl.side2;        // shadows the slot's first side stream inside this block
                    const uint32_t rows = (uint32_t)Wr * H;
                    const XYZZ<F>* cur = d_buckets;
                    int pp = 0;
                    for (uint32_t cnt = Lw; cnt > 1;) {
                        const uint32_t f = cnt >= fan ? fan : cnt, cnt_out = cnt / f;
                        const uint32_t total = rows * cnt_out;
                        XYZZ<F>* dst = cnt_out == 1 ? d_r : d_rstage[pp];
                        msm_colsum_kernel<F><<<nblk(total, 128), 128, 0, side>>>(cur, cnt, 1, f, cnt_out, total, dst);
                        ctx->launches++;
                        cur = dst;
                        pp ^= 1;
                        cnt = cnt_out;
                    }
                    ZA_CUDA(cudaEventRecord(sl.red_join, side));
                }
                // column sums: H rows of Lw points, eight rows per thread and stage
                const XYZZ<F>* cur = d_buckets;
                int pp = 0;
                for (uint32_t rows_in = H; rows_in > 1;) {
                    const uint32_t f = rows_in >= fan ? fan : rows_in, rows_out = rows_in / f;
                    const uint32_t total = (uint32_t)Wr * rows_out * Lw;
                    msm_colsum_kernel<F><<<nblk(total, 128), 128, 0, side>>>(cur, rows_in, Lw, f, rows_out, total, d_stage[pp]);
                    ctx->launches++;
                    cur = d_stage[pp];
                    pp ^= 1;
                    rows_in = rows_out;
                }
                d_c = cur;
                ZA_CUDA(cudaStreamWaitEvent(side, sl.red_join, 0));
            }
            static const bool w2d = !(getenv("ZA_MSM_W2D") && atoi(getenv("ZA_MSM_W2D")) == 0);
            sl.red_sR = sl.red_sC = 0;
            if (w2d) {
                XYZZ<F>* d_tmp = d_out3 + (size_t)6 * Wr;
                const uint32_t warps = 2u * (uint32_t)Wr * 64u;
                msm_w2d_sums_kernel<F><<<nblk((size_t)warps * 32, 128), 128, 0, side>>>(d_r, H > 1 ? H : 0, d_c, Lw, (uint32_t)Wr, d_tmp);
                msm_w2d_final_kernel<F><<<2 * Wr, 64, 0, side>>>(d_tmp, H > 1 ? H : 0, Lw, (uint32_t)Wr, d_out3);
                ctx->launches += 2;
                auto lg_b = [](uint32_t n) { int s = 0; const uint32_t b = n < 32 ? n : 32; while ((1u << s) < b) s++; return s; };
                sl.red_sR = H > 1 ? lg_b(H) : 0; sl.red_sC = lg_b(Lw);
            } else {
                msm_small_weighted_kernel<F><<<2 * Wr, 256, 0, side>>>(d_r, H > 1 ? H : 0, d_c, Lw, (uint32_t)Wr, d_out3);
                ctx->launches++;
            }
            sl.red_k = (int)k; sl.red_H = H; sl.red_Lw = Lw;
            d_win = d_out3;
            win_points = 6 * (size_t)Wr;
        } else {
            const XYZZ<F>* src = d_buckets;
            int li = 0;
            for (const Level& lv : levels) {
                XYZZ<F>* dst = d_lvl[li & 1];
                // acc of level l carries weight SEG^l: every earlier level had the full segment size SEG
                const uint32_t total = (uint32_t)Wr * lv.n_out;
                msm_weighted_level_kernel<F><<<nblk(total, 128), 128, 0, side>>>(src, lv.n_in, lv.s, lv.n_out, total, SEG_LOG * li, dst, d_pool, pool_stride,
                                                                                 lv.pool_off);
                ctx->launches++;
                src = dst;
                li++;
            }
            // the final T (one per window) is the plain sum of all buckets: pool slot pool_len
            ZA_CUDA(cudaMemcpy2DAsync(d_pool + pool_len, (size_t)pool_stride * sizeof(XYZZ<F>), src, sizeof(XYZZ<F>), sizeof(XYZZ<F>), Wr,
                                      cudaMemcpyDeviceToDevice, side));
            {
                const XYZZ<F>* gin = d_pool;
                uint32_t gstride = pool_stride, gcount = pool_stride;
                XYZZ<F>* gout = d_grp;
                for (uint32_t groups : grp_sizes) {
                    msm_group_reduce_kernel<F><<<nblk((size_t)Wr * groups * 32, 128), 128, 0, side>>>(gin, gstride, gcount, GRP_PER, groups, (uint32_t)Wr * groups, gout);
                    ctx->launches++;
                    gin = gout; gstride = groups; gcount = groups;
                    gout += (size_t)Wr * groups;
                }
            }
        }
    }
    ZA_CUDA(cudaGetLastError());
    ZA_CUDA(cudaMemcpyAsync((uint8_t*)sl.host_win + 64, d_win, win_points * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, side));
    if (ctx->profile) ZA_CUDA(cudaMemcpyAsync(sl.host_win, d_offsets + nkeys, 4, cudaMemcpyDeviceToHost, side));
    ZA_CUDA(cudaEventRecord(sl.done, side));
    if (timeline) cudaEventRecord(sl.dbg_done, side);
    sl.done_valid = true;
    sl.busy = true;
}
template void msm_enqueue<Fq>(Ctx*, int, const Affine<Fq>*, const uint32_t*, size_t, bool, int, const Affine<Fq>*, int, int, const MsmPart*, int);
template void msm_enqueue<Fq2>(Ctx*, int, const Affine<Fq2>*, const uint32_t*, size_t, bool, int, const Affine<Fq2>*, int, int, const MsmPart*, int);

// parts_out != nullptr: the multiexp covered several queries (K4'); their sums go to parts_out[0 .. nparts), the return
// value is the first one
template <class F>
XYZZ<F> msm_finish(Ctx* ctx, int slot_id, XYZZ<F>* parts_out) {
    MsmSlot& sl = ctx->slots[slot_id];
    XYZZ<F> result = XYZZ<F>::inf();
    if (!sl.busy) throw ZaError(ZA_ERR_INVALID, "msm_finish on an idle slot");
    sl.busy = false;
    if ((sl.nparts > 1) != (parts_out != nullptr)) throw ZaError(ZA_ERR_INVALID, "internal: msm_finish and msm_enqueue disagree about the number of queries");
    if (sl.kind == 0) return result;
    ZA_CUDA(cudaEventSynchronize(sl.done));
    const XYZZ<F>* win = reinterpret_cast<const XYZZ<F>*>((const uint8_t*)sl.host_win + 64);
    if (sl.kind == 1) {
        for (int i = 0; i < sl.warps; i++) xyzz_add<F>(result, win[i]);
        return result;
    }
    if (ctx->profile) {
        // algorithmic work of the accumulation in Fq products: a batched-affine addition is 5M + 1S (G2: 5 Fq2
        // products of 3 and one squaring of 2 = 17), an XYZZ mixed addition 8M + 2S (G2: 28); the point count
        // after every round came back with the window sums
        uint32_t E[8];
        memcpy(E, sl.host_win, 4 * (1 + sl.rounds));
        const bool g1 = sizeof(F) == sizeof(Fq);
        double work = 0;
        for (int r = 0; r < sl.rounds; r++) work += (double)(E[r] - E[r + 1]) * (g1 ? 6.0 : 17.0);
        work += (double)E[sl.rounds] * (g1 ? 10.0 : 28.0);
        ctx->prof_work[sl.acc_cat] += work;
    }
    // row / column reduction: the six partial results of bucket space w -> sum_b (b + 1) S_b  (K6')
    std::vector<XYZZ<F>> spaces;
    if (sl.red_mode == 1) {
        // o[0] + 2^shift o[2] = sum_i i X_i (K6'c splits the array once more; shift = 0 and o[2] = infinity otherwise), o[1] = sum_i X_i
        auto weighted = [](const XYZZ<F>* o, int shift, bool plus_one) {
            XYZZ<F> r = o[2];
            for (int k = 0; k < shift; k++) r = xyzz_dbl<F>(r);
            xyzz_add<F>(r, o[0]);
            if (plus_one) xyzz_add<F>(r, o[1]);
            return r;
        };
        spaces.resize(sl.W);
        for (int w = 0; w < sl.W; w++) {
            XYZZ<F> t = weighted(win + 3 * (size_t)(sl.W + w), sl.red_sC, true);
            if (sl.red_H > 1) {
                XYZZ<F> hi = weighted(win + 3 * (size_t)w, sl.red_sR, false);
                for (int k = 0; k < sl.red_k; k++) hi = xyzz_dbl<F>(hi);
                xyzz_add<F>(t, hi);
            }
            spaces[w] = t;
        }
        win = spaces.data();
    }
    if (parts_out) {
        if (sl.red_mode != 1) throw ZaError(ZA_ERR_INVALID, "internal: several queries need the row / column reduction");
        for (int p = 0; p < sl.nparts; p++) parts_out[p] = win[p];
        return win[0];
    }
    // window combination on the host: result = sum_w 2^(c w) S_w   (bellman: `higher.double()` x c, then add)
    for (int w = sl.W - 1; w >= 0; w--) {
        for (int k = 0; k < sl.c; k++) result = xyzz_dbl<F>(result);
        xyzz_add<F>(result, win[w]);
    }
    return result;
}
template XYZZ<Fq> msm_finish<Fq>(Ctx*, int, XYZZ<Fq>*);
template XYZZ<Fq2> msm_finish<Fq2>(Ctx*, int, XYZZ<Fq2>*);

// After a failed enqueue or collect: wait for whatever is in flight on this context and return every slot to idle, so
// that the next call on the context starts clean (a slot left busy would refuse every later multiexp).
void msm_abort(Ctx* ctx) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->g2_stream) cudaStreamSynchronize(ctx->g2_stream);
    if (ctx->h_stream) cudaStreamSynchronize(ctx->h_stream);
    for (MsmSlot& sl : ctx->slots) {
        if (sl.side) cudaStreamSynchronize(sl.side);
        if (sl.side2) cudaStreamSynchronize(sl.side2);
        sl.busy = false; sl.kind = 0; sl.sort_users = 0;
    }
    cudaGetLastError();
}

// One multiexp, synchronously.  d_scalars: n canonical scalars on device; has_infinity: bases may contain (0,0).
template <class F>
XYZZ<F> msm_run(Ctx* ctx, const Affine<F>* d_bases, const uint32_t* d_scalars, size_t n, bool has_infinity) {
    msm_enqueue<F>(ctx, 0, d_bases, d_scalars, n, has_infinity, -1, nullptr, 0, 0);
    return msm_finish<F>(ctx, 0);
}
template XYZZ<Fq> msm_run<Fq>(Ctx*, const Affine<Fq>*, const uint32_t*, size_t, bool);
template XYZZ<Fq2> msm_run<Fq2>(Ctx*, const Affine<Fq2>*, const uint32_t*, size_t, bool);

template <class F>
void bases_generate(Ctx* ctx, Affine<F>* d_out, size_t n, uint64_t first, const Affine<F>& G) {
    if (!n) return;
    DevBuf tmp(n * sizeof(XYZZ<F>));
    size_t threads = (n + GEN_CHUNK - 1) / GEN_CHUNK;
    bases_gen_chain_kernel<F><<<nblk(threads, 128), 128, 0, ctx->stream>>>(tmp.as<XYZZ<F>>(), n, first, G, GEN_CHUNK);
    bases_gen_normalise_kernel<F><<<nblk(threads, 128), 128, 0, ctx->stream>>>(tmp.as<XYZZ<F>>(), d_out, n);
    ctx->launches += 2;
    ZA_CUDA(cudaGetLastError());
    ZA_CUDA(cudaStreamSynchronize(ctx->stream));
}
template void bases_generate<Fq>(Ctx*, Affine<Fq>*, size_t, uint64_t, const Affine<Fq>&);
template void bases_generate<Fq2>(Ctx*, Affine<Fq2>*, size_t, uint64_t, const Affine<Fq2>&);

// XYZZ -> affine for n points with shared inversions (points at infinity become (0,0)).
template <class F>
void xyzz_normalise(Ctx* ctx, const XYZZ<F>* d_in, Affine<F>* d_out, size_t n) {
    if (!n) return;
    size_t threads = (n + GEN_CHUNK - 1) / GEN_CHUNK;
    bases_gen_normalise_kernel<F><<<nblk(threads, 128), 128, 0, ctx->stream>>>(d_in, d_out, n);
    ctx->launches++;
    ZA_CUDA(cudaGetLastError());
}
template void xyzz_normalise<Fq>(Ctx*, const XYZZ<Fq>*, Affine<Fq>*, size_t);
template void xyzz_normalise<Fq2>(Ctx*, const XYZZ<Fq2>*, Affine<Fq2>*, size_t);

// Build the fixed-base table of `n` points: d_table[i*W + w] = 2^(c w) * P_i (affine).
template <class F>
void bases_table_build(Ctx* ctx, const Affine<F>* d_pts, size_t n, int c, int W, Affine<F>* d_table) {
    if (!n) return;
    // in pieces of 2^21 points: the projective staging buffer is twice the size of the table part it becomes (a 2^26-point
    // table is 56 GB; staged in one piece it needed 168 GB)
    const size_t piece = (size_t)1 << 21;
    DevBuf tmp(std::min(n, piece) * (size_t)W * sizeof(XYZZ<F>));
    for (size_t i0 = 0; i0 < n; i0 += piece) {
        const size_t ni = std::min(piece, n - i0), total = ni * (size_t)W;
        bases_table_kernel<F><<<nblk(ni, 128), 128, 0, ctx->stream>>>(d_pts + i0, ni, c, W, tmp.as<XYZZ<F>>());
        size_t threads = (total + GEN_CHUNK - 1) / GEN_CHUNK;
        bases_gen_normalise_kernel<F><<<nblk(threads, 128), 128, 0, ctx->stream>>>(tmp.as<XYZZ<F>>(), d_table + i0 * (size_t)W, total);
        ctx->launches += 2;
    }
    ZA_CUDA(cudaGetLastError());
    ZA_CUDA(cudaStreamSynchronize(ctx->stream));
}
template void bases_table_build<Fq>(Ctx*, const Affine<Fq>*, size_t, int, int, Affine<Fq>*);
template void bases_table_build<Fq2>(Ctx*, const Affine<Fq2>*, size_t, int, int, Affine<Fq2>*);

// measured 32-bit multiply-add throughput of the whole chip, in IMAD/s
double imad_peak(Ctx* ctx) {
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 2000;
    DevBuf out((size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    ZA_CUDA(cudaEventCreate(&e0));
    ZA_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        ZA_CUDA(cudaEventRecord(e0, ctx->stream));
        imad_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out.as<uint32_t>(), 0x9e3779b1u, 0x7f4a7c15u, iters);
        ZA_CUDA(cudaEventRecord(e1, ctx->stream));
        ZA_CUDA(cudaEventSynchronize(e1));
        ctx->launches++;
        float ms = 0;
        ZA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double rate = (double)blocks * threads * iters * 128.0 / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

}  // namespace za
