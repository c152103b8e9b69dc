// Short-Weierstrass arithmetic for BN254 G1 (over Fq) and G2 (over Fq2), a = 0.
//
// Replaces (un-vendored) pairing_ce bn256::{G1,G1Affine,G2,G2Affine}:
// `add_assign_mixed`, `add_assign`, `double`, `mul`, `into_affine`
// (driven from bellman's multiexp / create_proof, entered at
// /root/reference/prover/src/groth16/prover.rs:173).
//
// bellman accumulates in Jacobian (X,Y,Z); here buckets are XYZZ
// (X, Y, ZZ, ZZZ with x = X/ZZ, y = Y/ZZZ): mixed addition is 8M+2S with no
// inversion and negation is free.  Results are compared only after
// normalisation to affine, which is representation independent
// (SURVEY.md §0.5), so proofs stay bit-exact.
//
// Every formula handles the exceptional cases bellman handles
// (acc = inf, P = inf, P = acc -> double, P = -acc -> inf): duplicate and
// opposite points do occur in real proving keys.
//
// Inlining policy on the device: the G1 mixed addition (the bucket-accumulation hot loop) is inlined;
// every other group operation, and everything over Fq2, is a real call (`*_call` below) — otherwise
// a G2 kernel inlines ~100 Montgomery products per loop body and ptxas needs many minutes.
#pragma once
#include "ff.cuh"

namespace za {

// Affine point; the point at infinity is encoded as (0, 0), which is not on
// either curve (b != 0).
template <class F>
struct Affine {
    F x, y;
    ZA_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static ZA_HD Affine inf() { Affine p; p.x = F::zero(); p.y = F::zero(); return p; }
};

template <class F>
struct XYZZ {
    F X, Y, ZZ, ZZZ;
    ZA_HD bool is_inf() const { return ZZ.is_zero(); }
    static ZA_HD XYZZ inf() {
        XYZZ p;
        p.X = F::zero(); p.Y = F::zero(); p.ZZ = F::zero(); p.ZZZ = F::zero();
        return p;
    }
    static ZA_HD XYZZ from_affine(const Affine<F>& a) {
        if (a.is_inf()) return inf();
        XYZZ p;
        p.X = a.x; p.Y = a.y; p.ZZ = F::one(); p.ZZZ = F::one();
        return p;
    }
};

template <class F> ZA_HD XYZZ<F> xyzz_dbl_affine(const F& x, const F& y);
template <class F> ZA_HD XYZZ<F> xyzz_dbl(const XYZZ<F>& p);

// ------------------------------------------------------------------ formula bodies
// 2 * (x, y) -> XYZZ          (mdbl-2008-s-1, a = 0)
template <class F>
ZA_HD XYZZ<F> xyzz_dbl_affine_body(const F& x, const F& y) {
    XYZZ<F> r;
    F U = dbl(y);
    F V = sqr(U);
    F W = U * V;
    F S = x * V;
    F x2 = sqr(x);
    F M = dbl(x2) + x2;
    r.X = sqr(M) - dbl(S);
    r.Y = M * (S - r.X) - W * y;
    r.ZZ = V;
    r.ZZZ = W;
    return r;
}

// 2 * acc                      (dbl-2008-s-1, a = 0)
template <class F>
ZA_HD XYZZ<F> xyzz_dbl_body(const XYZZ<F>& p) {
    if (p.is_inf()) return p;
    XYZZ<F> r;
    F U = dbl(p.Y);
    F V = sqr(U);
    F W = U * V;
    F S = p.X * V;
    F x2 = sqr(p.X);
    F M = dbl(x2) + x2;
    r.X = sqr(M) - dbl(S);
    r.Y = M * (S - r.X) - W * p.Y;
    r.ZZ = V * p.ZZ;
    r.ZZZ = W * p.ZZZ;
    return r;
}

// acc += (x, negate ? -y : y)  (madd-2008-s)   — the bucket-accumulation step
template <class F>
ZA_HD void xyzz_madd_body(XYZZ<F>& acc, const F& x, const F& y_in, bool negate) {
    F y = negate ? -y_in : y_in;
    if (acc.is_inf()) {
        acc.X = x; acc.Y = y; acc.ZZ = F::one(); acc.ZZZ = F::one();
        return;
    }
    F U2 = x * acc.ZZ;
    F S2 = y * acc.ZZZ;
    F P = U2 - acc.X;
    F R = S2 - acc.Y;
    if (P.is_zero()) {
        if (R.is_zero()) acc = xyzz_dbl_affine<F>(x, y);
        else acc = XYZZ<F>::inf();
        return;
    }
    F PP = sqr(P);
    F PPP = P * PP;
    F Q = acc.X * PP;
    F X3 = sqr(R) - PPP - dbl(Q);
    acc.Y = R * (Q - X3) - acc.Y * PPP;
    acc.X = X3;
    acc.ZZ = acc.ZZ * PP;
    acc.ZZZ = acc.ZZZ * PPP;
}

// acc += b                     (add-2008-s)
template <class F>
ZA_HD void xyzz_add_body(XYZZ<F>& acc, const XYZZ<F>& b) {
    if (b.is_inf()) return;
    if (acc.is_inf()) { acc = b; return; }
    F U1 = acc.X * b.ZZ;
    F U2 = b.X * acc.ZZ;
    F S1 = acc.Y * b.ZZZ;
    F S2 = b.Y * acc.ZZZ;
    F P = U2 - U1;
    F R = S2 - S1;
    if (P.is_zero()) {
        if (R.is_zero()) acc = xyzz_dbl<F>(acc);
        else acc = XYZZ<F>::inf();
        return;
    }
    F PP = sqr(P);
    F PPP = P * PP;
    F Q = U1 * PP;
    F X3 = sqr(R) - PPP - dbl(Q);
    acc.Y = R * (Q - X3) - S1 * PPP;
    acc.X = X3;
    acc.ZZ = acc.ZZ * b.ZZ * PP;
    acc.ZZZ = acc.ZZZ * b.ZZZ * PPP;
}

// ------------------------------------------------------------------ dispatch
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ void g1_dbl_affine_call(XYZZ<Fq>* r, const Fq* x, const Fq* y) { *r = xyzz_dbl_affine_body<Fq>(*x, *y); }
static __device__ __noinline__ void g2_dbl_affine_call(XYZZ<Fq2>* r, const Fq2* x, const Fq2* y) { *r = xyzz_dbl_affine_body<Fq2>(*x, *y); }
static __device__ __noinline__ void g1_dbl_call(XYZZ<Fq>* r, const XYZZ<Fq>* p) { *r = xyzz_dbl_body<Fq>(*p); }
static __device__ __noinline__ void g2_dbl_call(XYZZ<Fq2>* r, const XYZZ<Fq2>* p) { *r = xyzz_dbl_body<Fq2>(*p); }
static __device__ __noinline__ void g2_madd_call(XYZZ<Fq2>* acc, const Fq2* x, const Fq2* y, bool negate) { xyzz_madd_body<Fq2>(*acc, *x, *y, negate); }
static __device__ __noinline__ void g1_add_call(XYZZ<Fq>* acc, const XYZZ<Fq>* b) { xyzz_add_body<Fq>(*acc, *b); }
static __device__ __noinline__ void g2_add_call(XYZZ<Fq2>* acc, const XYZZ<Fq2>* b) { xyzz_add_body<Fq2>(*acc, *b); }
#endif

template <class F>
ZA_HD XYZZ<F> xyzz_dbl_affine(const F& x, const F& y) {
#if defined(__CUDA_ARCH__)
    XYZZ<F> r;
    if constexpr (sizeof(F) == sizeof(Fq2)) g2_dbl_affine_call(&r, &x, &y);
    else g1_dbl_affine_call(&r, &x, &y);
    return r;
#else
    return xyzz_dbl_affine_body<F>(x, y);
#endif
}
template <class F>
ZA_HD XYZZ<F> xyzz_dbl(const XYZZ<F>& p) {
#if defined(__CUDA_ARCH__)
    XYZZ<F> r;
    if constexpr (sizeof(F) == sizeof(Fq2)) g2_dbl_call(&r, &p);
    else g1_dbl_call(&r, &p);
    return r;
#else
    return xyzz_dbl_body<F>(p);
#endif
}
template <class F>
ZA_HD void xyzz_madd(XYZZ<F>& acc, const F& x, const F& y, bool negate) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(F) == sizeof(Fq2)) { g2_madd_call(&acc, &x, &y, negate); return; }
#endif
    xyzz_madd_body<F>(acc, x, y, negate);
}

#if defined(__CUDA_ARCH__)
// The G1 mixed addition of the bucket-accumulation loop: same formula as xyzz_madd_body, inlined into the
// kernel, but its ten field products go through the one shared copy fq_mul_call (the fully inlined version is
// ~50 KB of code and stalls on instruction fetch: profiles/r01_ncu_summary.md, "no_instructions").
static __device__ __forceinline__ void g1_madd_hot(XYZZ<Fq>& acc, const Fq& x, const Fq& y_in, bool negate) {
    Fq y = negate ? -y_in : y_in;
    if (acc.is_inf()) {
        acc.X = x; acc.Y = y; acc.ZZ = Fq::one(); acc.ZZZ = Fq::one();
        return;
    }
    Fq U2 = fq_mul_call(x, acc.ZZ);
    Fq S2 = fq_mul_call(y, acc.ZZZ);
    Fq P = U2 - acc.X;
    Fq R = S2 - acc.Y;
    if (P.is_zero()) {
        if (R.is_zero()) acc = xyzz_dbl_affine<Fq>(x, y);
        else acc = XYZZ<Fq>::inf();
        return;
    }
    Fq PP = fq_sqr_call(P);
    Fq PPP = fq_mul_call(P, PP);
    Fq Q = fq_mul_call(acc.X, PP);
    Fq X3 = fq_sqr_call(R) - PPP - dbl(Q);
    acc.Y = fq_mul_call(R, Q - X3) - fq_mul_call(acc.Y, PPP);
    acc.X = X3;
    acc.ZZ = fq_mul_call(acc.ZZ, PP);
    acc.ZZZ = fq_mul_call(acc.ZZZ, PPP);
}
static __device__ __forceinline__ void xyzz_madd_hot(XYZZ<Fq>& acc, const Fq& x, const Fq& y, bool negate) { g1_madd_hot(acc, x, y, negate); }
static __device__ __forceinline__ void xyzz_madd_hot(XYZZ<Fq2>& acc, const Fq2& x, const Fq2& y, bool negate) { g2_madd_call(&acc, &x, &y, negate); }
#endif
template <class F>
ZA_HD void xyzz_madd(XYZZ<F>& acc, const Affine<F>& p, bool negate = false) {
    if (p.is_inf()) return;
    xyzz_madd<F>(acc, p.x, p.y, negate);
}
template <class F>
ZA_HD void xyzz_add(XYZZ<F>& acc, const XYZZ<F>& b) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(F) == sizeof(Fq2)) g2_add_call(&acc, &b);
    else g1_add_call(&acc, &b);
#else
    xyzz_add_body<F>(acc, b);
#endif
}

template <class F>
ZA_HD XYZZ<F> xyzz_neg(const XYZZ<F>& p) {
    XYZZ<F> r = p;
    r.Y = -p.Y;
    return r;
}

// scalar given as 8 canonical little-endian 32-bit words; MSB-first double-and-add
template <class F>
ZA_HD XYZZ<F> xyzz_mul(const XYZZ<F>& p, const uint32_t* k) {
    XYZZ<F> r = XYZZ<F>::inf();
    int top = 255;
    while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1u)) top--;
    for (int i = top; i >= 0; i--) {
        r = xyzz_dbl<F>(r);
        if ((k[i >> 5] >> (i & 31)) & 1u) xyzz_add<F>(r, p);
    }
    return r;
}

// x = X/ZZ, y = Y/ZZZ with one inversion: t = 1/(ZZ*ZZZ), 1/ZZ = t*ZZZ, 1/ZZZ = t*ZZ
template <class F>
ZA_HD Affine<F> xyzz_to_affine(const XYZZ<F>& p) {
    if (p.is_inf()) return Affine<F>::inf();
    F t = inv(p.ZZ * p.ZZZ);
    F iZZ = t * p.ZZZ;
    F iZZZ = t * p.ZZ;
    Affine<F> a;
    a.x = p.X * iZZ;
    a.y = p.Y * iZZZ;
    return a;
}

// y^2 == x^3 + b
template <class F>
ZA_HD bool affine_on_curve(const Affine<F>& p, const F& b) {
    return sqr(p.y) == sqr(p.x) * p.x + b;
}

typedef Affine<Fq> G1Affine;
typedef Affine<Fq2> G2Affine;
typedef XYZZ<Fq> G1XYZZ;
typedef XYZZ<Fq2> G2XYZZ;

}  // namespace za
