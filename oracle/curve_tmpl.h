/* ORACLE — TEST INFRASTRUCTURE ONLY (see field.h header).  PARITY UNPINNED.
 *
 * curve_tmpl.h: Jacobian short-Weierstrass arithmetic (a = 0), instantiated
 * twice by za_oracle.c: G1 over Fq and G2 over Fq2.  Restates pairing_ce
 * bn256 `CurveProjective::{double, add_assign, add_assign_mixed, mul_assign,
 * into_affine}`; reached from
 * /root/reference/prover/src/groth16/prover.rs:173 via bellman's multiexp.
 * Projective = (X,Y,Z), infinity <=> Z = 0; affine carries an explicit flag.
 *
 * Required macros: PT (prefix), FT (field type), F_ADD/F_SUB/F_MUL/F_SQR/F_NEG/
 * F_DBL/F_INV/F_ISZERO/F_EQ/F_ZERO/F_ONE.
 */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define AFF CAT(PT, _affine)
#define JAC CAT(PT, _jac)
#define FN(name) CAT(PT, name)

typedef struct { FT x, y; int inf; } AFF;
typedef struct { FT x, y, z; } JAC;

static inline JAC FN(_jac_zero)(void) { JAC r; r.x = F_ZERO(); r.y = F_ONE(); r.z = F_ZERO(); return r; }
static inline int FN(_jac_is_zero)(const JAC *p) { return F_ISZERO(&p->z); }
static inline AFF FN(_affine_zero)(void) { AFF r; r.x = F_ZERO(); r.y = F_ONE(); r.inf = 1; return r; }
static inline JAC FN(_from_affine)(const AFF *p) {
    if (p->inf) return FN(_jac_zero)();
    JAC r; r.x = p->x; r.y = p->y; r.z = F_ONE(); return r;
}

static inline void FN(_double)(JAC *p) {
    if (FN(_jac_is_zero)(p)) return;
    FT a = F_SQR(&p->x), b = F_SQR(&p->y), c = F_SQR(&b);
    FT d = F_ADD(&p->x, &b); d = F_SQR(&d); d = F_SUB(&d, &a); d = F_SUB(&d, &c); d = F_DBL(&d);
    FT e = F_DBL(&a); e = F_ADD(&e, &a);
    FT f = F_SQR(&e);
    FT z3 = F_MUL(&p->y, &p->z); z3 = F_DBL(&z3);
    FT x3 = F_SUB(&f, &d); x3 = F_SUB(&x3, &d);
    FT y3 = F_SUB(&d, &x3); y3 = F_MUL(&y3, &e);
    FT c8 = F_DBL(&c); c8 = F_DBL(&c8); c8 = F_DBL(&c8);
    y3 = F_SUB(&y3, &c8);
    p->x = x3; p->y = y3; p->z = z3;
}

static inline void FN(_add)(JAC *p, const JAC *o) {
    if (FN(_jac_is_zero)(p)) { *p = *o; return; }
    if (FN(_jac_is_zero)(o)) return;
    FT z1z1 = F_SQR(&p->z), z2z2 = F_SQR(&o->z);
    FT u1 = F_MUL(&p->x, &z2z2), u2 = F_MUL(&o->x, &z1z1);
    FT s1 = F_MUL(&p->y, &o->z); s1 = F_MUL(&s1, &z2z2);
    FT s2 = F_MUL(&o->y, &p->z); s2 = F_MUL(&s2, &z1z1);
    if (F_EQ(&u1, &u2) && F_EQ(&s1, &s2)) { FN(_double)(p); return; }
    FT h = F_SUB(&u2, &u1);
    FT i = F_DBL(&h); i = F_SQR(&i);
    FT j = F_MUL(&h, &i);
    FT r = F_SUB(&s2, &s1); r = F_DBL(&r);
    FT v = F_MUL(&u1, &i);
    FT x3 = F_SQR(&r); x3 = F_SUB(&x3, &j); x3 = F_SUB(&x3, &v); x3 = F_SUB(&x3, &v);
    FT y3 = F_SUB(&v, &x3); y3 = F_MUL(&y3, &r);
    FT s1j = F_MUL(&s1, &j); s1j = F_DBL(&s1j); y3 = F_SUB(&y3, &s1j);
    FT z3 = F_ADD(&p->z, &o->z); z3 = F_SQR(&z3); z3 = F_SUB(&z3, &z1z1); z3 = F_SUB(&z3, &z2z2); z3 = F_MUL(&z3, &h);
    p->x = x3; p->y = y3; p->z = z3;
}

static inline void FN(_add_mixed)(JAC *p, const AFF *o) {
    if (o->inf) return;
    if (FN(_jac_is_zero)(p)) { p->x = o->x; p->y = o->y; p->z = F_ONE(); return; }
    FT z1z1 = F_SQR(&p->z);
    FT u2 = F_MUL(&o->x, &z1z1);
    FT s2 = F_MUL(&o->y, &p->z); s2 = F_MUL(&s2, &z1z1);
    if (F_EQ(&p->x, &u2) && F_EQ(&p->y, &s2)) { FN(_double)(p); return; }
    FT h = F_SUB(&u2, &p->x);
    FT hh = F_SQR(&h);
    FT i = F_DBL(&hh); i = F_DBL(&i);
    FT j = F_MUL(&h, &i);
    FT r = F_SUB(&s2, &p->y); r = F_DBL(&r);
    FT v = F_MUL(&p->x, &i);
    FT x3 = F_SQR(&r); x3 = F_SUB(&x3, &j); x3 = F_SUB(&x3, &v); x3 = F_SUB(&x3, &v);
    FT yj = F_MUL(&p->y, &j); yj = F_DBL(&yj);
    FT y3 = F_SUB(&v, &x3); y3 = F_MUL(&y3, &r); y3 = F_SUB(&y3, &yj);
    FT z3 = F_ADD(&p->z, &h); z3 = F_SQR(&z3); z3 = F_SUB(&z3, &z1z1); z3 = F_SUB(&z3, &hh);
    p->x = x3; p->y = y3; p->z = z3;
}

static inline void FN(_negate)(JAC *p) { if (!FN(_jac_is_zero)(p)) p->y = F_NEG(&p->y); }

/* scalar = canonical little-endian u64[4]; MSB-first double-and-add */
static inline JAC FN(_mul)(const JAC *p, const uint64_t *k) {
    JAC r = FN(_jac_zero)();
    int started = 0;
    for (int i = 255; i >= 0; i--) {
        if (started) FN(_double)(&r);
        if ((k[i >> 6] >> (i & 63)) & 1) { FN(_add)(&r, p); started = 1; }
    }
    return r;
}

static inline AFF FN(_into_affine)(const JAC *p) {
    if (FN(_jac_is_zero)(p)) return FN(_affine_zero)();
    FT zi = F_INV(&p->z), zi2 = F_SQR(&zi), zi3 = F_MUL(&zi2, &zi);
    AFF r; r.x = F_MUL(&p->x, &zi2); r.y = F_MUL(&p->y, &zi3); r.inf = 0;
    return r;
}

static inline int FN(_on_curve)(const AFF *p, const FT *b) {
    if (p->inf) return 1;
    FT y2 = F_SQR(&p->y), x3 = F_SQR(&p->x); x3 = F_MUL(&x3, &p->x); x3 = F_ADD(&x3, b);
    return F_EQ(&y2, &x3);
}

#undef AFF
#undef JAC
#undef FN
#undef PT
#undef FT
#undef F_ADD
#undef F_SUB
#undef F_MUL
#undef F_SQR
#undef F_NEG
#undef F_DBL
#undef F_INV
#undef F_ISZERO
#undef F_EQ
#undef F_ZERO
#undef F_ONE
