// BN254 prime-field arithmetic on 8 x 32-bit limbs, Montgomery form (R = 2^256).
//
// Replaces (un-vendored) ff_ce's derived `Fr` / `Fq`
// (`PrimeField` impls used at /root/reference/prover/src/groth16/format.rs:35,51,204)
// for the device side of the Groth16 hot path.
//
// One source, two targets:
//   * device (sm_100a): every limb operation is a PTX `mad.lo.cc / madc.hi.cc / addc`
//     carry-chain instruction; ptxas pairs adjacent lo/hi into IMAD.WIDE.U32 + carry.
//   * host: the same algorithm runs on an emulated carry flag, so the exact
//     limb schedule that the GPU executes is unit-testable without a GPU
//     (tests/test_ff_host.py) and is reused by the host-side proof assembly.
//
// The Montgomery product keeps two half-accumulators ("even"/"odd" limb
// alignment) so that each 32x32->64 product lands in one aligned register pair
// and every carry chain is a single pass.
#pragma once
#include <stdint.h>
#include <stddef.h>

// Force inlining only in the device pass: force-inlining the host copies of the curve formulas
// (proof assembly) makes the host compiler spend minutes on a few giant functions.
#if defined(__CUDACC__)
#if defined(__CUDA_ARCH__)
#define ZA_HD __host__ __device__ __forceinline__
#else
#define ZA_HD __host__ __device__ inline
#endif
#define ZA_D __device__ __forceinline__
#else
#define ZA_HD inline
#endif

namespace za {

// ---------------------------------------------------------------------------
// carry-chain primitives
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)

ZA_D uint32_t p_mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZA_D uint32_t p_mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// full 32x32 -> 64 product as one instruction (IMAD.WIDE.U32); the halves then enter add.cc chains
ZA_D void p_mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    uint64_t p;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p));
}
ZA_D uint32_t p_add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZA_D uint32_t p_addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZA_D uint32_t p_addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZA_D uint32_t p_sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZA_D uint32_t p_subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZA_D uint32_t p_subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZA_D uint32_t p_mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZA_D uint32_t p_mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZA_D uint32_t p_madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZA_D uint32_t p_madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZA_D uint32_t p_madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

#else  // host emulation of the PTX carry flag

// Number of times an instruction that cannot report a carry-out would have
// produced one (must stay 0; checked by tests/test_ff_host.py).
struct HostCC {
    uint32_t cf = 0;
    uint64_t lost = 0;
};
inline HostCC& host_cc() { static thread_local HostCC s; return s; }

inline uint32_t p_mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t p_mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline void p_mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { uint64_t p = (uint64_t)a * b; lo = (uint32_t)p; hi = (uint32_t)(p >> 32); }
inline uint32_t emu_add(uint64_t a, uint64_t b, uint64_t cin, bool set_cc) {
    uint64_t s = a + b + cin;
    if (set_cc) host_cc().cf = (uint32_t)(s >> 32);
    else if (s >> 32) host_cc().lost++;
    return (uint32_t)s;
}
inline uint32_t emu_sub(uint64_t a, uint64_t b, uint64_t bin, bool set_cc) {
    uint64_t s = a - b - bin;
    if (set_cc) host_cc().cf = (uint32_t)((s >> 32) & 1);
    return (uint32_t)s;
}
inline uint32_t p_add_cc(uint32_t a, uint32_t b) { return emu_add(a, b, 0, true); }
inline uint32_t p_addc_cc(uint32_t a, uint32_t b) { return emu_add(a, b, host_cc().cf, true); }
inline uint32_t p_addc(uint32_t a, uint32_t b) { return emu_add(a, b, host_cc().cf, false); }
inline uint32_t p_sub_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, 0, true); }
inline uint32_t p_subc_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, host_cc().cf, true); }
inline uint32_t p_subc(uint32_t a, uint32_t b) { return emu_sub(a, b, host_cc().cf, false); }
inline uint32_t p_mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(p_mul_lo(a, b), c, 0, true); }
inline uint32_t p_mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(p_mul_hi(a, b), c, 0, true); }
inline uint32_t p_madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(p_mul_lo(a, b), c, host_cc().cf, true); }
inline uint32_t p_madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(p_mul_hi(a, b), c, host_cc().cf, true); }
inline uint32_t p_madc_hi(uint32_t a, uint32_t b, uint32_t c) { return emu_add(p_mul_hi(a, b), c, host_cc().cf, false); }

#endif

// ---------------------------------------------------------------------------
// field parameters.  Moduli pinned in-tree by the reference:
//   r: /root/reference/compiler/src/algebra/fs.rs:15-16, prover/src/groth16/ethereum.rs:173
//   q: /root/reference/prover/src/groth16/ethereum.rs:37
// R, R^2 and -p^-1 mod 2^32 are derived and re-derived by
// tests/test_ff_host.py from the moduli with python integers.
// ---------------------------------------------------------------------------
struct FrParams {
    static ZA_HD constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static ZA_HD constexpr uint32_t one(int i) {  // R mod r
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static ZA_HD constexpr uint32_t r2(int i) {  // R^2 mod r
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
    static constexpr uint32_t inv = 0xefffffffu;  // -r^-1 mod 2^32
};

struct FqParams {
    static ZA_HD constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static ZA_HD constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static ZA_HD constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
    static constexpr uint32_t inv = 0xe4866389u;  // -q^-1 mod 2^32
};

// ---------------------------------------------------------------------------
// Fp<P>: value in [0, p), Montgomery form, little-endian 32-bit limbs.
// Memory image == ff_ce's 4 x u64 little-endian Montgomery limbs.
// ---------------------------------------------------------------------------
template <class P>
struct alignas(16) Fp {
    uint32_t v[8];

    static ZA_HD Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }
    static ZA_HD Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
        return r;
    }
    static ZA_HD Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::r2(i);
        return r;
    }
    ZA_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i];
        return o == 0;
    }
    ZA_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    ZA_HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// r = (x >= p) ? x - p : x      (x < 2p < 2^256)
template <class P>
ZA_HD void fp_final_sub(uint32_t* x) {
    uint32_t t[8];
    t[0] = p_sub_cc(x[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = p_subc_cc(x[i], P::mod(i));
    uint32_t borrow = p_subc(0u, 0u);  // 0 or 0xffffffff
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = borrow ? x[i] : t[i];
}

template <class P>
ZA_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    r.v[0] = p_add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = p_addc_cc(a.v[i], b.v[i]);
    r.v[7] = p_addc(a.v[7], b.v[7]);
    fp_final_sub<P>(r.v);
    return r;
}

template <class P>
ZA_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    r.v[0] = p_sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = p_subc_cc(a.v[i], b.v[i]);
    uint32_t borrow = p_subc(0u, 0u);
    r.v[0] = p_add_cc(r.v[0], P::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = p_addc_cc(r.v[i], P::mod(i) & borrow);
    // the carry out of the top limb cancels the earlier borrow and is discarded on purpose
    r.v[7] = p_addc_cc(r.v[7], P::mod(7) & borrow);
    return r;
}

template <class P>
ZA_HD Fp<P> fp_neg(const Fp<P>& a) {
    if (a.is_zero()) return a;
    Fp<P> r;
    r.v[0] = p_sub_cc(P::mod(0), a.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = p_subc_cc(P::mod(i), a.v[i]);
    r.v[7] = p_subc(P::mod(7), a.v[7]);
    return r;
}

template <class P>
ZA_HD Fp<P> fp_dbl(const Fp<P>& a) { return fp_add<P>(a, a); }

// One CIOS round: (E,O) += a*bi ; (E,O) += m*q ; implicit >>32 by role swap.
// Value represented: sum E[j] 2^(32j) + sum O[j] 2^(32(j+1)).
// The a*bi products do not depend on the running reduction, so they are issued as eight independent
// IMAD.WIDE (p_mul_wide) and folded in with add.cc chains; the m*q products do depend on it and go
// through mad.lo.cc/madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32.X (multiply-add with
// carry in and out).  ~136 integer-pipe multiply instructions per field product.
template <class P, bool FIRST>
ZA_HD void fp_mad_redc(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {
    uint32_t pl[8], ph[8];
#pragma unroll
    for (int j = 0; j < 8; j++) p_mul_wide(pl[j], ph[j], a[j], bi);
    if (FIRST) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            O[j] = pl[j + 1];
            O[j + 1] = ph[j + 1];
            E[j] = pl[j];
            E[j + 1] = ph[j];
        }
    } else {
        // O still holds the previous round's low word (now weight 2^0) in O[1].
        E[0] = p_add_cc(E[0], O[1]);
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
            O[j] = p_addc_cc(pl[j + 1], O[j + 2]);
            O[j + 1] = p_addc_cc(ph[j + 1], O[j + 3]);
        }
        O[6] = p_addc_cc(pl[7], 0u);
        O[7] = p_addc(ph[7], 0u);
        E[0] = p_add_cc(E[0], pl[0]);
        E[1] = p_addc_cc(E[1], ph[0]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            E[j] = p_addc_cc(E[j], pl[j]);
            E[j + 1] = p_addc_cc(E[j + 1], ph[j]);
        }
        O[7] = p_addc(O[7], 0u);
    }
    uint32_t q = p_mul_lo(E[0], P::inv);
#if defined(ZA_FF_REDC_WIDE)
    // variant: the m*q products as plain IMAD.WIDE, carries in the ALU.  Measured slower on B200 (11.5 vs
    // 15.0 T IMAD-class/s in scratch/mb/modmul_bench: the ALU pipe becomes the limit); kept for reference.
    uint32_t ql[8], qh[8];
#pragma unroll
    for (int j = 0; j < 8; j++) p_mul_wide(ql[j], qh[j], P::mod(j), q);
    O[0] = p_add_cc(O[0], ql[1]);
    O[1] = p_addc_cc(O[1], qh[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        O[j] = p_addc_cc(O[j], ql[j + 1]);
        O[j + 1] = (j == 6) ? p_addc(O[j + 1], qh[j + 1]) : p_addc_cc(O[j + 1], qh[j + 1]);
    }
    E[0] = p_add_cc(E[0], ql[0]);
    E[1] = p_addc_cc(E[1], qh[0]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        E[j] = p_addc_cc(E[j], ql[j]);
        E[j + 1] = p_addc_cc(E[j + 1], qh[j]);
    }
    O[7] = p_addc(O[7], 0u);
#else
    O[0] = p_mad_lo_cc(P::mod(1), q, O[0]);
    O[1] = p_madc_hi_cc(P::mod(1), q, O[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        O[j] = p_madc_lo_cc(P::mod(j + 1), q, O[j]);
        O[j + 1] = (j == 6) ? p_madc_hi(P::mod(j + 1), q, O[j + 1]) : p_madc_hi_cc(P::mod(j + 1), q, O[j + 1]);
    }
    E[0] = p_mad_lo_cc(P::mod(0), q, E[0]);
    E[1] = p_madc_hi_cc(P::mod(0), q, E[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        E[j] = p_madc_lo_cc(P::mod(j), q, E[j]);
        E[j + 1] = p_madc_hi_cc(P::mod(j), q, E[j + 1]);
    }
    O[7] = p_addc(O[7], 0u);
#endif
}

template <class P>
ZA_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
#if !defined(__CUDA_ARCH__) && !defined(ZA_FF_EMULATE_PTX)
    // Host fast path (proof assembly, window combination): 4 x 64-bit CIOS.  Define
    // ZA_FF_EMULATE_PTX to run the device limb schedule on the host instead (unit tests do).
    typedef unsigned __int128 u128;
    uint64_t A[4], B[4], M[4], t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        A[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
        B[i] = (uint64_t)b.v[2 * i] | ((uint64_t)b.v[2 * i + 1] << 32);
        M[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
    }
    // -p^-1 mod 2^64 from the 32-bit constant by one Newton step: x' = x (2 + p x)  for x = -p^-1
    uint64_t ninv = (uint64_t)P::inv;
    ninv = ninv * (2 + M[0] * ninv);
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)A[j] * B[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * ninv;
        c = (u128)m * M[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * M[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    bool ge = t[4] != 0;
    if (!ge) {
        ge = true;
        for (int i = 3; i >= 0; i--) { if (t[i] > M[i]) break; if (t[i] < M[i]) { ge = false; break; } }
    }
    if (ge) {
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - M[i] - borrow; t[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1; }
    }
    Fp<P> r;
    for (int i = 0; i < 4; i++) { r.v[2 * i] = (uint32_t)t[i]; r.v[2 * i + 1] = (uint32_t)(t[i] >> 32); }
    return r;
#else
    uint32_t ev[8], od[8];
    fp_mad_redc<P, true>(ev, od, a.v, b.v[0]);
    fp_mad_redc<P, false>(od, ev, a.v, b.v[1]);
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
        fp_mad_redc<P, false>(ev, od, a.v, b.v[i]);
        fp_mad_redc<P, false>(od, ev, a.v, b.v[i + 1]);
    }
    Fp<P> r;
    r.v[0] = p_add_cc(ev[0], od[1]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = p_addc_cc(ev[i], od[i + 1]);
    r.v[7] = p_addc(ev[7], 0u);
    fp_final_sub<P>(r.v);
    return r;
#endif
}

// ---------------------------------------------------------------------------
// 512-bit products, the stand-alone Montgomery reduction, and what is built on them: a dedicated squaring
// (36 limb products instead of 64) and the lazily reduced Fq2 product (three wide products, two reductions).
// Same instruction discipline as fp_mad_redc: independent mul.wide products folded in with add.cc chains, the
// q*p rows as mad.lo.cc / madc.hi.cc pairs.
// ---------------------------------------------------------------------------
// T[0..15] = a * b for any 256-bit a, b.  X collects the products a_j b_i with j even (limb i + j), Y those with j odd
// (limb i + j, kept one limb lower); each half-row tiles eight consecutive limbs, so a row is two carry chains.
ZA_HD void u256_mul_wide(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    uint32_t Y[16];
    uint32_t* X = T;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        p_mul_wide(X[j], X[j + 1], a[j], b[0]);
        p_mul_wide(Y[j], Y[j + 1], a[j + 1], b[0]);
    }
    X[8] = 0; Y[8] = 0;
#pragma unroll
    for (int i = 1; i < 8; i++) {
        uint32_t pl[8], ph[8];
#pragma unroll
        for (int j = 0; j < 8; j++) p_mul_wide(pl[j], ph[j], a[j], b[i]);
        X[i] = p_add_cc(X[i], pl[0]);
        X[i + 1] = p_addc_cc(X[i + 1], ph[0]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            X[i + j] = p_addc_cc(X[i + j], pl[j]);
            X[i + j + 1] = p_addc_cc(X[i + j + 1], ph[j]);
        }
        X[i + 8] = p_addc(0u, 0u);
        Y[i] = p_add_cc(Y[i], pl[1]);
        Y[i + 1] = p_addc_cc(Y[i + 1], ph[1]);
#pragma unroll
        for (int j = 2; j < 6; j += 2) {
            Y[i + j] = p_addc_cc(Y[i + j], pl[j + 1]);
            Y[i + j + 1] = p_addc_cc(Y[i + j + 1], ph[j + 1]);
        }
        Y[i + 6] = p_addc_cc(Y[i + 6], pl[7]);
        if (i < 7) { Y[i + 7] = p_addc_cc(Y[i + 7], ph[7]); Y[i + 8] = p_addc(0u, 0u); }
        else Y[i + 7] = p_addc(Y[i + 7], ph[7]);       // 2^32 Y <= a b: limb 15 of Y stays empty
    }
    // T = X + 2^32 Y
    T[1] = p_add_cc(X[1], Y[0]);
#pragma unroll
    for (int k = 2; k < 15; k++) T[k] = p_addc_cc(X[k], Y[k - 1]);
    T[15] = p_addc(X[15], Y[14]);
}

// T[0..15] = a^2.  The 28 cross products a_i a_j (i < j) by diagonals d = j - i — diagonal d tiles the limbs
// [d, 16 - d) — doubled, plus the eight squares, which tile all 16 limbs.
ZA_HD void u256_sqr_wide(uint32_t* T, const uint32_t* a) {
    uint32_t S[16], c[8];
    S[0] = 0; S[15] = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) p_mul_wide(S[2 * i + 1], S[2 * i + 2], a[i], a[i + 1]);
#pragma unroll
    for (int d = 2; d < 8; d++) {
        uint32_t pl[6], ph[6];
#pragma unroll
        for (int i = 0; i + d < 8; i++) p_mul_wide(pl[i], ph[i], a[i], a[i + d]);
        S[d] = p_add_cc(S[d], pl[0]);
        S[d + 1] = p_addc_cc(S[d + 1], ph[0]);
#pragma unroll
        for (int i = 1; i + d < 8; i++) {
            S[2 * i + d] = p_addc_cc(S[2 * i + d], pl[i]);
            S[2 * i + d + 1] = p_addc_cc(S[2 * i + d + 1], ph[i]);
        }
        c[d] = p_addc(0u, 0u);                          // carry out of the diagonal: one unit of limb 16 - d
    }
    S[9] = p_add_cc(S[9], c[7]);
#pragma unroll
    for (int k = 10; k < 15; k++) S[k] = p_addc_cc(S[k], c[16 - k]);
    S[15] = p_addc(S[15], 0u);
    // T = 2 S + sum a_i^2 2^(64 i)
    uint32_t ql[8], qh[8];
#pragma unroll
    for (int i = 0; i < 8; i++) p_mul_wide(ql[i], qh[i], a[i], a[i]);
    S[1] = p_add_cc(S[1], S[1]);
#pragma unroll
    for (int k = 2; k < 15; k++) S[k] = p_addc_cc(S[k], S[k]);
    S[15] = p_addc(S[15], S[15]);
    T[0] = ql[0];
    T[1] = p_add_cc(S[1], qh[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) {
        T[2 * i] = p_addc_cc(S[2 * i], ql[i]);
        T[2 * i + 1] = p_addc_cc(S[2 * i + 1], qh[i]);
    }
    T[14] = p_addc_cc(S[14], ql[7]);
    T[15] = p_addc(S[15], qh[7]);
}

// One round of the stand-alone reduction: (E, O) += q p with q = -E[0] / p mod 2^32, then the implicit >> 32 by
// role swap.  The previous round's odd array is shifted down by two limbs inside the multiply-adds (its limb 1,
// now of weight 2^0, moves over to E[0] first).
template <class P, bool FIRST>
ZA_HD void fp_redc_round(uint32_t* E, uint32_t* O) {
    if (!FIRST) E[0] = p_add_cc(E[0], O[1]);
    const uint32_t q = p_mul_lo(E[0], P::inv);
    if (FIRST) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) p_mul_wide(O[j], O[j + 1], P::mod(j + 1), q);
    } else {
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
            O[j] = p_madc_lo_cc(P::mod(j + 1), q, O[j + 2]);
            O[j + 1] = p_madc_hi_cc(P::mod(j + 1), q, O[j + 3]);
        }
        O[6] = p_madc_lo_cc(P::mod(7), q, 0u);
        O[7] = p_madc_hi(P::mod(7), q, 0u);
    }
    E[0] = p_mad_lo_cc(P::mod(0), q, E[0]);
    E[1] = p_madc_hi_cc(P::mod(0), q, E[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        E[j] = p_madc_lo_cc(P::mod(j), q, E[j]);
        E[j + 1] = p_madc_hi_cc(P::mod(j), q, E[j + 1]);
    }
    O[7] = p_addc(O[7], 0u);
}

// T / 2^256 mod p for T < p 2^256, result in [0, p).  Only the low half takes part in the eight rounds (a Montgomery
// product of T_lo by 1, at most p); the high half is added at the end: (T_lo + sum q_i p 2^(32 i)) / 2^256 + T_hi < 2 p.
template <class P>
ZA_HD Fp<P> fp_redc_wide(const uint32_t* T) {
    uint32_t ev[8], od[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ev[i] = T[i];
    fp_redc_round<P, true>(ev, od);
    fp_redc_round<P, false>(od, ev);
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
        fp_redc_round<P, false>(ev, od);
        fp_redc_round<P, false>(od, ev);
    }
    Fp<P> r;
    r.v[0] = p_add_cc(ev[0], od[1]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = p_addc_cc(ev[i], od[i + 1]);
    r.v[7] = p_addc(ev[7], 0u);
    r.v[0] = p_add_cc(r.v[0], T[8]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = p_addc_cc(r.v[i], T[8 + i]);
    r.v[7] = p_addc(r.v[7], T[15]);
    fp_final_sub<P>(r.v);
    return r;
}

template <class P>
ZA_HD Fp<P> fp_sqr(const Fp<P>& a) {
#if !defined(__CUDA_ARCH__) && !defined(ZA_FF_EMULATE_PTX)
    return fp_mul<P>(a, a);
#else
    uint32_t T[16];
    u256_sqr_wide(T, a.v);
    return fp_redc_wide<P>(T);
#endif
}

// canonical (non-Montgomery) little-endian limbs  <->  Montgomery
template <class P>
ZA_HD Fp<P> fp_to_mont(const Fp<P>& canonical) { return fp_mul<P>(canonical, Fp<P>::r2()); }
template <class P>
ZA_HD Fp<P> fp_from_mont(const Fp<P>& a) {
    Fp<P> o = Fp<P>::zero();
    o.v[0] = 1;
    return fp_mul<P>(a, o);
}

// a^e, e given as 8 little-endian 32-bit words (not secret; variable time)
template <class P>
ZA_HD Fp<P> fp_pow(const Fp<P>& a, const uint32_t* e) {
    Fp<P> r = Fp<P>::one();
    bool started = false;
    for (int i = 255; i >= 0; i--) {
        if (started) r = fp_sqr<P>(r);
        if ((e[i >> 5] >> (i & 31)) & 1u) {
            r = started ? fp_mul<P>(r, a) : a;
            started = true;
        }
    }
    return r;
}

// a^-1 = a^(p-2)   (0 -> 0)
template <class P>
ZA_HD Fp<P> fp_inv(const Fp<P>& a) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = P::mod(i);
    e[0] -= 2;  // both moduli end in ...01 / ...47: no borrow
    return fp_pow<P>(a, e);
}

// a^-1 by Kaliski's almost-Montgomery inverse (binary extended Euclid: shifts, additions and subtractions
// only — no field products, so on the device it runs on the ALU pipe instead of the integer-multiply pipe, and
// takes ~1/10 of the time of the a^(p-2) ladder).  Phase 1 yields x = a^-1 2^k (254 <= k <= 508) as a plain
// residue; two Montgomery products rescale it: for a = A*R the result is A^-1 * R.  (0 -> 0.)
// The batched-affine bucket rounds (msm.cu) call this once per CTA per round.
ZA_HD int za_ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
ZA_HD void u256_shr(uint32_t* x, int z) {   // 1 <= z <= 31
#pragma unroll
    for (int i = 0; i < 7; i++) x[i] = (x[i] >> z) | (x[i + 1] << (32 - z));
    x[7] >>= z;
}
ZA_HD void u256_shl(uint32_t* x, int z) {   // 1 <= z <= 31
#pragma unroll
    for (int i = 7; i > 0; i--) x[i] = (x[i] << z) | (x[i - 1] >> (32 - z));
    x[0] <<= z;
}
template <class P>
ZA_HD Fp<P> fp_inv_kaliski(const Fp<P>& a) {
    if (a.is_zero()) return a;
    uint32_t u[8], v[8], r[8], s[8], t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { u[i] = P::mod(i); v[i] = a.v[i]; r[i] = 0; s[i] = 0; }
    s[0] = 1;
    int k = 0;
    for (;;) {
        while (!(v[0] & 1u)) {                       // v != 0 here
            const int z = v[0] ? za_ctz32(v[0]) : 31;
            u256_shr(v, z); u256_shl(r, z); k += z;
        }
        // u and v odd
        t[0] = p_sub_cc(v[0], u[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) t[i] = p_subc_cc(v[i], u[i]);
        const uint32_t borrow = p_subc(0u, 0u);
        if (!borrow) {                               // v >= u:  v = v - u,  s += r
            s[0] = p_add_cc(s[0], r[0]);
#pragma unroll
            for (int i = 1; i < 7; i++) s[i] = p_addc_cc(s[i], r[i]);
            s[7] = p_addc(s[7], r[7]);
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) { v[i] = t[i]; any |= t[i]; }
            if (!any) { u256_shl(r, 1); k++; break; }
        } else {                                     // u > v:  u = u - v,  r += s, then make u odd again
            u[0] = p_sub_cc(u[0], v[0]);
#pragma unroll
            for (int i = 1; i < 7; i++) u[i] = p_subc_cc(u[i], v[i]);
            u[7] = p_subc(u[7], v[7]);
            r[0] = p_add_cc(r[0], s[0]);
#pragma unroll
            for (int i = 1; i < 7; i++) r[i] = p_addc_cc(r[i], s[i]);
            r[7] = p_addc(r[7], s[7]);
            while (!(u[0] & 1u)) {
                const int z = u[0] ? za_ctz32(u[0]) : 31;
                u256_shr(u, z); u256_shl(s, z); k += z;
            }
        }
    }
    // r < 2p;  x = p - (r mod p) = a^-1 2^k
    fp_final_sub<P>(r);
    Fp<P> x;
    x.v[0] = p_sub_cc(P::mod(0), r[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) x.v[i] = p_subc_cc(P::mod(i), r[i]);
    x.v[7] = p_subc(P::mod(7), r[7]);
    // y = x * 2^(512-k):  montmul(montmul(x, R^2), 2^e) = x * 2^e,  e = 512 - k in [4, 258]
    x = fp_mul<P>(x, Fp<P>::r2());
    const int e = 512 - k;
    const int e1 = e > 253 ? 253 : e;
    Fp<P> pw = Fp<P>::zero();
#pragma unroll
    for (int i = 0; i < 8; i++) pw.v[i] = (i == (e1 >> 5)) ? (1u << (e1 & 31)) : 0u;
    x = fp_mul<P>(x, pw);
    for (int i = e1; i < e; i++) x = fp_dbl<P>(x);
    return x;
}

template <class P>
ZA_HD Fp<P> fp_from_u64(uint64_t x) {
    Fp<P> c = Fp<P>::zero();
    c.v[0] = (uint32_t)x;
    c.v[1] = (uint32_t)(x >> 32);
    return fp_to_mont<P>(c);
}

// true iff canonical limbs < p
template <class P>
ZA_HD bool fp_is_canonical(const uint32_t* x) {
    for (int i = 7; i >= 0; i--) {
        if (x[i] < P::mod(i)) return true;
        if (x[i] > P::mod(i)) return false;
    }
    return false;
}

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

// operator sugar shared by Fq and Fq2 so the curve templates read naturally
template <class P> ZA_HD Fp<P> operator+(const Fp<P>& a, const Fp<P>& b) { return fp_add<P>(a, b); }
template <class P> ZA_HD Fp<P> operator-(const Fp<P>& a, const Fp<P>& b) { return fp_sub<P>(a, b); }
template <class P> ZA_HD Fp<P> operator*(const Fp<P>& a, const Fp<P>& b) { return fp_mul<P>(a, b); }
// One shared, non-inlined copy of the Fq product for the places that would otherwise inline dozens of them
// (Fq2 arithmetic, the G1 bucket-accumulation loop): less code than the instruction cache, seconds of ptxas.
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ Fp<FqParams> fq_mul_call(const Fp<FqParams> a, const Fp<FqParams> b) { return fp_mul<FqParams>(a, b); }
static __device__ __noinline__ Fp<FqParams> fq_sqr_call(const Fp<FqParams> a) { return fp_sqr<FqParams>(a); }
#endif
template <class P> ZA_HD Fp<P> operator-(const Fp<P>& a) { return fp_neg<P>(a); }
template <class P> ZA_HD Fp<P> sqr(const Fp<P>& a) { return fp_sqr<P>(a); }
template <class P> ZA_HD Fp<P> dbl(const Fp<P>& a) { return fp_dbl<P>(a); }
template <class P> ZA_HD Fp<P> inv(const Fp<P>& a) { return fp_inv<P>(a); }

// ---------------------------------------------------------------------------
// Fq2 = Fq[u]/(u^2+1)   (pairing_ce bn256::Fq2, used at format.rs:57-64)
// ---------------------------------------------------------------------------
struct Fq2 {
    Fq c0, c1;
    static ZA_HD Fq2 zero() { Fq2 r; r.c0 = Fq::zero(); r.c1 = Fq::zero(); return r; }
    static ZA_HD Fq2 one() { Fq2 r; r.c0 = Fq::one(); r.c1 = Fq::zero(); return r; }
    ZA_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    ZA_HD bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
    ZA_HD bool operator!=(const Fq2& b) const { return !(*this == b); }
};
ZA_HD Fq2 operator+(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
ZA_HD Fq2 operator-(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
ZA_HD Fq2 operator-(const Fq2& a) { Fq2 r; r.c0 = -a.c0; r.c1 = -a.c1; return r; }
ZA_HD Fq2 dbl(const Fq2& a) { Fq2 r; r.c0 = dbl(a.c0); r.c1 = dbl(a.c1); return r; }
// On the device the Fq product under Fq2 is a real call: G2 kernels would otherwise inline ~40 Montgomery
// products per group operation (minutes of ptxas time, code far beyond the instruction cache).
#if defined(__CUDA_ARCH__)
ZA_D Fq fq2_base_mul(const Fq& a, const Fq& b) { return fq_mul_call(a, b); }
#else
inline Fq fq2_base_mul(const Fq& a, const Fq& b) { return fp_mul<FqParams>(a, b); }
#endif
// Lazily reduced Karatsuba: three 512-bit products, two Montgomery reductions (instead of three full products):
//   c1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1   (in [0, 2 p^2), the sums unreduced: < 2^255)
//   c0 = a0 b0 - a1 b1                        (+ p 2^256 if negative: in [0, p 2^256))
ZA_HD Fq2 fq2_mul_lazy(const Fq2& a, const Fq2& b) {
    uint32_t T0[16], T1[16], T2[16], sa[8], sb[8];
    sa[0] = p_add_cc(a.c0.v[0], a.c1.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) sa[i] = p_addc_cc(a.c0.v[i], a.c1.v[i]);
    sa[7] = p_addc(a.c0.v[7], a.c1.v[7]);
    sb[0] = p_add_cc(b.c0.v[0], b.c1.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) sb[i] = p_addc_cc(b.c0.v[i], b.c1.v[i]);
    sb[7] = p_addc(b.c0.v[7], b.c1.v[7]);
    u256_mul_wide(T2, sa, sb);
    u256_mul_wide(T0, a.c0.v, b.c0.v);
    T2[0] = p_sub_cc(T2[0], T0[0]);
#pragma unroll
    for (int i = 1; i < 15; i++) T2[i] = p_subc_cc(T2[i], T0[i]);
    T2[15] = p_subc(T2[15], T0[15]);
    u256_mul_wide(T1, a.c1.v, b.c1.v);
    T2[0] = p_sub_cc(T2[0], T1[0]);
#pragma unroll
    for (int i = 1; i < 15; i++) T2[i] = p_subc_cc(T2[i], T1[i]);
    T2[15] = p_subc(T2[15], T1[15]);
    Fq2 r;
    r.c1 = fp_redc_wide<FqParams>(T2);
    T0[0] = p_sub_cc(T0[0], T1[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) T0[i] = p_subc_cc(T0[i], T1[i]);
    const uint32_t borrow = p_subc(0u, 0u);
    T0[8] = p_add_cc(T0[8], FqParams::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < 7; i++) T0[8 + i] = p_addc_cc(T0[8 + i], FqParams::mod(i) & borrow);
    // the carry out of the top limb cancels the earlier borrow and is discarded on purpose
    T0[15] = p_addc_cc(T0[15], FqParams::mod(7) & borrow);
    r.c0 = fp_redc_wide<FqParams>(T0);
    return r;
}
// Karatsuba: 3 Fq products
ZA_HD Fq2 fq2_mul_karatsuba(const Fq2& a, const Fq2& b) {
    Fq aa = fq2_base_mul(a.c0, b.c0);
    Fq bb = fq2_base_mul(a.c1, b.c1);
    Fq s = fq2_base_mul(a.c0 + a.c1, b.c0 + b.c1);
    Fq2 r;
    r.c0 = aa - bb;
    r.c1 = s - aa - bb;
    return r;
}
// The Fq2 product of every kernel: on the device ONE shared copy of the lazily reduced product (320 limb
// products instead of 384); ZA_FQ2_KARATSUBA restores the three full products.  The host uses the 64-bit path.
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ Fq2 fq2_mul_call(const Fq2 a, const Fq2 b) { return fq2_mul_lazy(a, b); }
#endif
ZA_HD Fq2 operator*(const Fq2& a, const Fq2& b) {
#if defined(__CUDA_ARCH__) && !defined(ZA_FQ2_KARATSUBA)
    return fq2_mul_call(a, b);
#elif defined(ZA_FF_EMULATE_PTX) && !defined(ZA_FQ2_KARATSUBA)
    return fq2_mul_lazy(a, b);
#else
    return fq2_mul_karatsuba(a, b);
#endif
}
// complex squaring: 2 Fq products
ZA_HD Fq2 sqr(const Fq2& a) {
    Fq t = fq2_base_mul(a.c0, a.c1);
    Fq2 r;
    r.c0 = fq2_base_mul(a.c0 + a.c1, a.c0 - a.c1);
    r.c1 = dbl(t);
    return r;
}
ZA_HD Fq2 inv(const Fq2& a) {
    Fq n = inv(sqr(a.c0) + sqr(a.c1));
    Fq2 r;
    r.c0 = a.c0 * n;
    r.c1 = -(a.c1 * n);
    return r;
}

}  // namespace za
