/* ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called
 * from the product (za_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * PARITY UNPINNED: the arithmetic restated here lives in the reference's
 * un-vendored dependencies (prover/Cargo.toml:19-25: bellman_ce @ git
 * adria0/bellman#test/affinecoords, pairing_ce @ git adria0/pairing#feature/
 * affinecoords — no pinned revision, Cargo.lock git-ignored — and ff_ce 0.7.1).
 * None of that source is under /root/reference and no Rust toolchain exists
 * here, so this file follows the published algorithms of those crates
 *  and is pinned against (a) the constants the reference
 * tree does hold and (b) independent python-integer golden vectors
 * (tests/golden/, tests/golden/make_golden.py).
 *
 * field.h: BN254 Fr / Fq, 4 x u64 little-endian limbs, Montgomery form with
 * R = 2^256 — the representation ff_ce's #[derive(PrimeField)] generates
 * (call sites: /root/reference/prover/src/groth16/format.rs:35,51,204).
 * Moduli: /root/reference/compiler/src/algebra/fs.rs:15-16 (r),
 *         /root/reference/prover/src/groth16/ethereum.rs:37 (q).
 */
#ifndef ZA_ORACLE_FIELD_H
#define ZA_ORACLE_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;           /* element of Fr or Fq (context decides) */

typedef struct {
    uint64_t mod[4];
    uint64_t r1[4];   /* R   mod p (Montgomery one) */
    uint64_t r2[4];   /* R^2 mod p */
    uint64_t inv;     /* -p^-1 mod 2^64 */
} fctx;

static const fctx FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL},
    {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL},
    0xc2e1f593efffffffULL};
static const fctx FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL},
    {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL},
    0x87d20782e4866389ULL};

static inline int fe_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe *a, const fe *b) { return memcmp(a, b, sizeof(fe)) == 0; }
static inline fe fe_zero(void) { fe z = {{0, 0, 0, 0}}; return z; }
static inline fe fe_one(const fctx *F) { fe o; memcpy(o.l, F->r1, 32); return o; }

static inline int limbs_geq(const uint64_t *a, const uint64_t *b) {
    for (int i = 3; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
    return 1;
}
static inline uint64_t limbs_sub(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)t; borrow = (uint64_t)(t >> 64) & 1;
    }
    return borrow;
}
static inline uint64_t limbs_add(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t carry = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] + b[i] + carry;
        r[i] = (uint64_t)t; carry = (uint64_t)(t >> 64);
    }
    return carry;
}

static inline fe fe_add(const fctx *F, const fe *a, const fe *b) {
    fe r; limbs_add(r.l, a->l, b->l);            /* p < 2^254: no carry out */
    if (limbs_geq(r.l, F->mod)) limbs_sub(r.l, r.l, F->mod);
    return r;
}
static inline fe fe_sub(const fctx *F, const fe *a, const fe *b) {
    fe r; if (limbs_sub(r.l, a->l, b->l)) limbs_add(r.l, r.l, F->mod);
    return r;
}
static inline fe fe_neg(const fctx *F, const fe *a) {
    fe r; if (fe_is_zero(a)) return *a; limbs_sub(r.l, F->mod, a->l); return r;
}
static inline fe fe_dbl(const fctx *F, const fe *a) { return fe_add(F, a, a); }

/* Montgomery product, operand-scanning CIOS on 64-bit limbs */
static inline fe fe_mul(const fctx *F, const fe *a, const fe *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c; c >>= 64;
        }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = (u128)m * F->mod[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * F->mod[j] + t[j];
            t[j - 1] = (uint64_t)c; c >>= 64;
        }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fe r; memcpy(r.l, t, 32);
    if (t[4] || limbs_geq(r.l, F->mod)) limbs_sub(r.l, r.l, F->mod);
    return r;
}
static inline fe fe_sqr(const fctx *F, const fe *a) { return fe_mul(F, a, a); }

/* a^e, e = nlimbs little-endian u64 */
static inline fe fe_pow(const fctx *F, const fe *a, const uint64_t *e, int nlimbs) {
    fe r = fe_one(F);
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        r = fe_sqr(F, &r);
        if ((e[i >> 6] >> (i & 63)) & 1) r = fe_mul(F, &r, a);
    }
    return r;
}
static inline fe fe_inv(const fctx *F, const fe *a) {       /* a^(p-2); 0 -> 0 */
    uint64_t e[4]; memcpy(e, F->mod, 32); e[0] -= 2;
    return fe_pow(F, a, e, 4);
}
/* canonical little-endian integer <-> Montgomery  (PrimeField::from_repr / into_repr) */
static inline fe fe_from_canon(const fctx *F, const uint64_t *x) {
    fe t, r2; memcpy(t.l, x, 32); memcpy(r2.l, F->r2, 32); return fe_mul(F, &t, &r2);
}
static inline void fe_to_canon(const fctx *F, const fe *a, uint64_t *out) {
    fe one = {{1, 0, 0, 0}}; fe r = fe_mul(F, a, &one); memcpy(out, r.l, 32);
}
static inline fe fe_from_u64(const fctx *F, uint64_t x) { uint64_t c[4] = {x, 0, 0, 0}; return fe_from_canon(F, c); }
static inline int canon_lt_mod(const fctx *F, const uint64_t *x) { return !limbs_geq(x, F->mod); }

/* ---- Fq2 = Fq[u]/(u^2+1)  (pairing_ce bn256::Fq2; format.rs:57-64) ---- */
typedef struct { fe c0, c1; } fe2;
static inline fe2 fe2_zero(void) { fe2 z; z.c0 = fe_zero(); z.c1 = fe_zero(); return z; }
static inline fe2 fe2_one(void) { fe2 z; z.c0 = fe_one(&FQ); z.c1 = fe_zero(); return z; }
static inline int fe2_is_zero(const fe2 *a) { return fe_is_zero(&a->c0) && fe_is_zero(&a->c1); }
static inline int fe2_eq(const fe2 *a, const fe2 *b) { return fe_eq(&a->c0, &b->c0) && fe_eq(&a->c1, &b->c1); }
static inline fe2 fe2_add(const fe2 *a, const fe2 *b) { fe2 r; r.c0 = fe_add(&FQ, &a->c0, &b->c0); r.c1 = fe_add(&FQ, &a->c1, &b->c1); return r; }
static inline fe2 fe2_sub(const fe2 *a, const fe2 *b) { fe2 r; r.c0 = fe_sub(&FQ, &a->c0, &b->c0); r.c1 = fe_sub(&FQ, &a->c1, &b->c1); return r; }
static inline fe2 fe2_neg(const fe2 *a) { fe2 r; r.c0 = fe_neg(&FQ, &a->c0); r.c1 = fe_neg(&FQ, &a->c1); return r; }
static inline fe2 fe2_dbl(const fe2 *a) { return fe2_add(a, a); }
static inline fe2 fe2_conj(const fe2 *a) { fe2 r; r.c0 = a->c0; r.c1 = fe_neg(&FQ, &a->c1); return r; }
static inline fe2 fe2_mul(const fe2 *a, const fe2 *b) {      /* schoolbook: (a0b0 - a1b1) + (a0b1 + a1b0)u */
    fe t0 = fe_mul(&FQ, &a->c0, &b->c0), t1 = fe_mul(&FQ, &a->c1, &b->c1);
    fe t2 = fe_mul(&FQ, &a->c0, &b->c1), t3 = fe_mul(&FQ, &a->c1, &b->c0);
    fe2 r; r.c0 = fe_sub(&FQ, &t0, &t1); r.c1 = fe_add(&FQ, &t2, &t3); return r;
}
static inline fe2 fe2_sqr(const fe2 *a) { return fe2_mul(a, a); }
static inline fe2 fe2_mul_fe(const fe2 *a, const fe *k) { fe2 r; r.c0 = fe_mul(&FQ, &a->c0, k); r.c1 = fe_mul(&FQ, &a->c1, k); return r; }
static inline fe2 fe2_inv(const fe2 *a) {
    fe t0 = fe_sqr(&FQ, &a->c0), t1 = fe_sqr(&FQ, &a->c1);
    fe n = fe_add(&FQ, &t0, &t1); n = fe_inv(&FQ, &n);
    fe2 r; r.c0 = fe_mul(&FQ, &a->c0, &n); fe m = fe_mul(&FQ, &a->c1, &n); r.c1 = fe_neg(&FQ, &m); return r;
}
static inline fe2 fe2_pow(const fe2 *a, const uint64_t *e, int nlimbs) {
    fe2 r = fe2_one();
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        r = fe2_sqr(&r);
        if ((e[i >> 6] >> (i & 63)) & 1) r = fe2_mul(&r, a);
    }
    return r;
}
#endif
