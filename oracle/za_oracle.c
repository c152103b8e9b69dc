/* ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see field.h header for why).
 *
 * CPU restatement, in plain C, of the Groth16 proving path the reference enters at
 * /root/reference/prover/src/groth16/prover.rs:173 (`create_random_proof`) and of
 * the setup / verify calls beside it (prover.rs:122,191-200; helper.rs:153-158).
 * The algorithms are those of bellman_ce (groth16/{prover,generator,verifier}.rs,
 * domain.rs, multiexp.rs, multicore.rs) and pairing_ce bn256, restated from their
 * published behaviour because the crates are not
 * vendored in /root/reference.  Every function names the upstream routine it follows.
 *
 * It is deliberately a different implementation from the product: 4 x u64 limbs,
 * Jacobian coordinates, unsigned c = ceil(ln n) windows, bit-reversal radix-2 FFT.
 */
#include "za_oracle.h"
#include "field.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

/* ------------------------------------------------------------------ curves */
static inline fe fq_add_(const fe *a, const fe *b) { return fe_add(&FQ, a, b); }
static inline fe fq_sub_(const fe *a, const fe *b) { return fe_sub(&FQ, a, b); }
static inline fe fq_mul_(const fe *a, const fe *b) { return fe_mul(&FQ, a, b); }
static inline fe fq_sqr_(const fe *a) { return fe_sqr(&FQ, a); }
static inline fe fq_neg_(const fe *a) { return fe_neg(&FQ, a); }
static inline fe fq_dbl_(const fe *a) { return fe_dbl(&FQ, a); }
static inline fe fq_inv_(const fe *a) { return fe_inv(&FQ, a); }
static inline fe fq_one_(void) { return fe_one(&FQ); }

#define PT g1
#define FT fe
#define F_ADD fq_add_
#define F_SUB fq_sub_
#define F_MUL fq_mul_
#define F_SQR fq_sqr_
#define F_NEG fq_neg_
#define F_DBL fq_dbl_
#define F_INV fq_inv_
#define F_ISZERO fe_is_zero
#define F_EQ fe_eq
#define F_ZERO fe_zero
#define F_ONE fq_one_
#include "curve_tmpl.h"

#define PT g2
#define FT fe2
#define F_ADD fe2_add
#define F_SUB fe2_sub
#define F_MUL fe2_mul
#define F_SQR fe2_sqr
#define F_NEG fe2_neg
#define F_DBL fe2_dbl
#define F_INV fe2_inv
#define F_ISZERO fe2_is_zero
#define F_EQ fe2_eq
#define F_ZERO fe2_zero
#define F_ONE fe2_one
#include "curve_tmpl.h"

/* G2 generator: /root/reference/prover/src/groth16/ethereum.rs:28-31 (printed there as [c1, c0]) */
static const uint64_t G2_GEN_X_C0[4] = {0x46debd5cd992f6edULL, 0x674322d4f75edaddULL, 0x426a00665e5c4479ULL, 0x1800deef121f1e76ULL};
static const uint64_t G2_GEN_X_C1[4] = {0x97e485b7aef312c2ULL, 0xf1aa493335a9e712ULL, 0x7260bfb731fb5d25ULL, 0x198e9393920d483aULL};
static const uint64_t G2_GEN_Y_C0[4] = {0x4ce6cc0166fa7daaULL, 0xe3d1e7690c43d37bULL, 0x4aab71808dcb408fULL, 0x12c85ea5db8c6debULL};
static const uint64_t G2_GEN_Y_C1[4] = {0x55acdadcd122975bULL, 0xbc4b313370b38ef3ULL, 0xec9e99ad690c3395ULL, 0x090689d0585ff075ULL};
/* twist coefficient b' = 3/(9+u) */
static const uint64_t B2_C0[4] = {0x3267e6dc24a138e5ULL, 0xb5b4c5e559dbefa3ULL, 0x81be18991be06ac3ULL, 0x2b149d40ceb8aaaeULL};
static const uint64_t B2_C1[4] = {0xe4a2bd0685c315d2ULL, 0xa74fa084e52d1852ULL, 0xcd2cafadeed8fdf4ULL, 0x009713b03af0fed4ULL};
/* 2^28-th root of unity = 7^((r-1)/2^28): ff_ce derive with generator 7 */
static const uint64_t FR_ROOT_OF_UNITY[4] = {0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL};
#define FR_S 28
#define FR_NUM_BITS 254
#define FR_GENERATOR 7

static fe g1_b(void) { return fe_from_u64(&FQ, 3); }
static fe2 g2_b(void) { fe2 b; b.c0 = fe_from_canon(&FQ, B2_C0); b.c1 = fe_from_canon(&FQ, B2_C1); return b; }

/* ------------------------------------------------------- byte interchange */
static void rd_le(const uint8_t *p, uint64_t *out) { memcpy(out, p, 32); } /* host is little-endian */
static void wr_le(uint8_t *p, const uint64_t *in) { memcpy(p, in, 32); }
static fe fr_load(const uint8_t *p) { uint64_t c[4]; rd_le(p, c); return fe_from_canon(&FR, c); }
static void fr_store(uint8_t *p, const fe *a) { uint64_t c[4]; fe_to_canon(&FR, a, c); wr_le(p, c); }
static fe fq_load(const uint8_t *p) { uint64_t c[4]; rd_le(p, c); return fe_from_canon(&FQ, c); }
static void fq_store(uint8_t *p, const fe *a) { uint64_t c[4]; fe_to_canon(&FQ, a, c); wr_le(p, c); }
static int all_zero(const uint8_t *p, size_t n) { for (size_t i = 0; i < n; i++) if (p[i]) return 0; return 1; }

static g1_affine g1_load(const uint8_t *p) {
    g1_affine a;
    if (all_zero(p, 64)) return g1_affine_zero();
    a.x = fq_load(p); a.y = fq_load(p + 32); a.inf = 0; return a;
}
static void g1_store(uint8_t *p, const g1_affine *a) {
    if (a->inf) { memset(p, 0, 64); return; }
    fq_store(p, &a->x); fq_store(p + 32, &a->y);
}
static g2_affine g2_load(const uint8_t *p) {
    g2_affine a;
    if (all_zero(p, 128)) return g2_affine_zero();
    a.x.c0 = fq_load(p); a.x.c1 = fq_load(p + 32); a.y.c0 = fq_load(p + 64); a.y.c1 = fq_load(p + 96); a.inf = 0;
    return a;
}
static void g2_store(uint8_t *p, const g2_affine *a) {
    if (a->inf) { memset(p, 0, 128); return; }
    fq_store(p, &a->x.c0); fq_store(p + 32, &a->x.c1); fq_store(p + 64, &a->y.c0); fq_store(p + 96, &a->y.c1);
}

void ora_field_op(int field, int op, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    const fctx *F = field ? &FQ : &FR;
    uint64_t ca[4], cb[4] = {0, 0, 0, 0}, co[4];
    rd_le(a, ca); if (b) rd_le(b, cb);
    fe x = fe_from_canon(F, ca), y = fe_from_canon(F, cb), z;
    switch (op) {
    case 0: z = fe_add(F, &x, &y); break;
    case 1: z = fe_sub(F, &x, &y); break;
    case 2: z = fe_mul(F, &x, &y); break;
    case 3: z = fe_inv(F, &x); break;
    default: z = fe_neg(F, &x); break;
    }
    fe_to_canon(F, &z, co); wr_le(out, co);
}
void ora_fq2_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    fe2 x, y = fe2_zero(), z;
    x.c0 = fq_load(a); x.c1 = fq_load(a + 32);
    if (b) { y.c0 = fq_load(b); y.c1 = fq_load(b + 32); }
    switch (op) {
    case 0: z = fe2_add(&x, &y); break;
    case 1: z = fe2_sub(&x, &y); break;
    case 2: z = fe2_mul(&x, &y); break;
    case 3: z = fe2_inv(&x); break;
    default: z = fe2_neg(&x); break;
    }
    fq_store(out, &z.c0); fq_store(out + 32, &z.c1);
}

int ora_g1_mul(const uint8_t *p, const uint8_t *k, uint8_t *out) {
    g1_affine a = g1_load(p); g1_jac j = g1_from_affine(&a); uint64_t kk[4]; rd_le(k, kk);
    g1_jac r = g1_mul(&j, kk); g1_affine ra = g1_into_affine(&r); g1_store(out, &ra); return 0;
}
int ora_g1_add(const uint8_t *p, const uint8_t *q, uint8_t *out) {
    g1_affine a = g1_load(p), b = g1_load(q); g1_jac j = g1_from_affine(&a); g1_add_mixed(&j, &b);
    g1_affine ra = g1_into_affine(&j); g1_store(out, &ra); return 0;
}
int ora_g2_mul(const uint8_t *p, const uint8_t *k, uint8_t *out) {
    g2_affine a = g2_load(p); g2_jac j = g2_from_affine(&a); uint64_t kk[4]; rd_le(k, kk);
    g2_jac r = g2_mul(&j, kk); g2_affine ra = g2_into_affine(&r); g2_store(out, &ra); return 0;
}
int ora_g2_add(const uint8_t *p, const uint8_t *q, uint8_t *out) {
    g2_affine a = g2_load(p), b = g2_load(q); g2_jac j = g2_from_affine(&a); g2_add_mixed(&j, &b);
    g2_affine ra = g2_into_affine(&j); g2_store(out, &ra); return 0;
}
int ora_g1_on_curve(const uint8_t *p) { g1_affine a = g1_load(p); fe b = g1_b(); return g1_on_curve(&a, &b); }
int ora_g2_on_curve(const uint8_t *p) { g2_affine a = g2_load(p); fe2 b = g2_b(); return g2_on_curve(&a, &b); }
void ora_g1_generator(uint8_t *out) {
    g1_affine g; g.x = fe_from_u64(&FQ, 1); g.y = fe_from_u64(&FQ, 2); g.inf = 0; g1_store(out, &g);
}
void ora_g2_generator(uint8_t *out) {
    wr_le(out, G2_GEN_X_C0); wr_le(out + 32, G2_GEN_X_C1); wr_le(out + 64, G2_GEN_Y_C0); wr_le(out + 96, G2_GEN_Y_C1);
}

/* out[i] = (i+1) * base: batches of Jacobian running sums normalised with one shared inversion */
#define MULT_BATCH 1024
int ora_g1_multiples(const uint8_t *base, size_t n, uint8_t *out) {
    g1_affine b = g1_load(base); g1_jac acc = g1_jac_zero();
    g1_jac *buf = malloc(sizeof(g1_jac) * MULT_BATCH); fe *pre = malloc(sizeof(fe) * MULT_BATCH);
    for (size_t s = 0; s < n; s += MULT_BATCH) {
        size_t m = n - s < MULT_BATCH ? n - s : MULT_BATCH;
        for (size_t i = 0; i < m; i++) { g1_add_mixed(&acc, &b); buf[i] = acc; }
        fe run = fe_one(&FQ);
        for (size_t i = 0; i < m; i++) { pre[i] = run; if (!fe_is_zero(&buf[i].z)) run = fe_mul(&FQ, &run, &buf[i].z); }
        run = fe_inv(&FQ, &run);
        for (size_t i = m; i-- > 0;) {
            g1_affine a;
            if (fe_is_zero(&buf[i].z)) { a = g1_affine_zero(); }
            else {
                fe zi = fe_mul(&FQ, &run, &pre[i]); run = fe_mul(&FQ, &run, &buf[i].z);
                fe zi2 = fe_sqr(&FQ, &zi), zi3 = fe_mul(&FQ, &zi2, &zi);
                a.x = fe_mul(&FQ, &buf[i].x, &zi2); a.y = fe_mul(&FQ, &buf[i].y, &zi3); a.inf = 0;
            }
            g1_store(out + (s + i) * 64, &a);
        }
    }
    free(buf); free(pre); return 0;
}
int ora_g2_multiples(const uint8_t *base, size_t n, uint8_t *out) {
    g2_affine b = g2_load(base); g2_jac acc = g2_jac_zero();
    g2_jac *buf = malloc(sizeof(g2_jac) * MULT_BATCH); fe2 *pre = malloc(sizeof(fe2) * MULT_BATCH);
    for (size_t s = 0; s < n; s += MULT_BATCH) {
        size_t m = n - s < MULT_BATCH ? n - s : MULT_BATCH;
        for (size_t i = 0; i < m; i++) { g2_add_mixed(&acc, &b); buf[i] = acc; }
        fe2 run = fe2_one();
        for (size_t i = 0; i < m; i++) { pre[i] = run; if (!fe2_is_zero(&buf[i].z)) run = fe2_mul(&run, &buf[i].z); }
        run = fe2_inv(&run);
        for (size_t i = m; i-- > 0;) {
            g2_affine a;
            if (fe2_is_zero(&buf[i].z)) { a = g2_affine_zero(); }
            else {
                fe2 zi = fe2_mul(&run, &pre[i]); run = fe2_mul(&run, &buf[i].z);
                fe2 zi2 = fe2_sqr(&zi), zi3 = fe2_mul(&zi2, &zi);
                a.x = fe2_mul(&buf[i].x, &zi2); a.y = fe2_mul(&buf[i].y, &zi3); a.inf = 0;
            }
            g2_store(out + (s + i) * 128, &a);
        }
    }
    free(buf); free(pre); return 0;
}

/* ---------------------------------------------------------------- Worker
 * bellman_ce multicore.rs `Worker`: scope(n) splits n items into
 * ceil(n / cpus)-sized chunks, one thread each. */
typedef struct { void (*fn)(void *, size_t, size_t, int); void *arg; size_t lo, hi; int tid; } job_t;
static void *job_tramp(void *p) { job_t *j = p; j->fn(j->arg, j->lo, j->hi, j->tid); return NULL; }
static void worker_scope(int cpus, size_t n, void (*fn)(void *, size_t, size_t, int), void *arg) {
    if (cpus < 1) cpus = 1;
    size_t chunk = n < (size_t)cpus ? 1 : (n + cpus - 1) / cpus;
    size_t njobs = (n + chunk - 1) / chunk;
    if (njobs <= 1) { if (n) fn(arg, 0, n, 0); return; }
    pthread_t *th = malloc(sizeof(pthread_t) * njobs); job_t *jobs = malloc(sizeof(job_t) * njobs);
    for (size_t j = 0; j < njobs; j++) {
        jobs[j].fn = fn; jobs[j].arg = arg; jobs[j].lo = j * chunk; jobs[j].hi = (j + 1) * chunk < n ? (j + 1) * chunk : n; jobs[j].tid = (int)j;
        pthread_create(&th[j], NULL, job_tramp, &jobs[j]);
    }
    for (size_t j = 0; j < njobs; j++) pthread_join(th[j], NULL);
    free(th); free(jobs);
}
static int log2_floor(unsigned x) { int l = 0; while ((1u << (l + 1)) <= x) l++; return l; }

/* ------------------------------------------------------- EvaluationDomain
 * bellman_ce domain.rs. */
typedef struct { fe *coeffs; size_t m; int exp; fe omega, omegainv, geninv, minv; } domain_t;

static int domain_init(domain_t *d, fe *coeffs_owned, size_t len) {     /* EvaluationDomain::from_coeffs */
    size_t m = 1; int exp = 0;
    while (m < len) { m *= 2; exp++; if (exp >= FR_S) return ORA_ERR_POLY_DEGREE_TOO_LARGE; }
    fe omega = fe_from_canon(&FR, FR_ROOT_OF_UNITY);
    for (int i = exp; i < FR_S; i++) omega = fe_sqr(&FR, &omega);
    d->coeffs = realloc(coeffs_owned, sizeof(fe) * m);
    for (size_t i = len; i < m; i++) d->coeffs[i] = fe_zero();
    d->m = m; d->exp = exp; d->omega = omega; d->omegainv = fe_inv(&FR, &omega);
    fe g = fe_from_u64(&FR, FR_GENERATOR); d->geninv = fe_inv(&FR, &g);
    fe mm = fe_from_u64(&FR, (uint64_t)m); d->minv = fe_inv(&FR, &mm);
    return 0;
}
static uint32_t bitreverse(uint32_t n, int l) { uint32_t r = 0; for (int i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; } return r; }
static fe fr_pow_u64(const fe *a, uint64_t e) { return fe_pow(&FR, a, &e, 1); }

static void serial_fft(fe *a, size_t n, const fe *omega, int log_n) {   /* domain.rs serial_fft */
    for (size_t k = 0; k < n; k++) { size_t rk = bitreverse((uint32_t)k, log_n); if (k < rk) { fe t = a[k]; a[k] = a[rk]; a[rk] = t; } }
    size_t m = 1;
    for (int s = 0; s < log_n; s++) {
        fe w_m = fr_pow_u64(omega, (uint64_t)(n / (2 * m)));
        for (size_t k = 0; k < n; k += 2 * m) {
            fe w = fe_one(&FR);
            for (size_t j = 0; j < m; j++) {
                fe t = fe_mul(&FR, &a[k + j + m], &w);
                fe tmp = fe_sub(&FR, &a[k + j], &t);
                a[k + j + m] = tmp;
                a[k + j] = fe_add(&FR, &a[k + j], &t);
                w = fe_mul(&FR, &w, &w_m);
            }
        }
        m *= 2;
    }
}
typedef struct { fe *a; fe **tmp; const fe *omega; fe new_omega; int log_n, log_cpus, log_new_n; } pfft_t;
static void pfft_stage1(void *p, size_t lo, size_t hi, int tid) {
    (void)tid; pfft_t *c = p; size_t num_cpus = (size_t)1 << c->log_cpus, new_n = (size_t)1 << c->log_new_n, mask = ((size_t)1 << c->log_n) - 1;
    for (size_t j = lo; j < hi; j++) {
        fe *tmp = c->tmp[j];
        fe omega_j = fr_pow_u64(c->omega, (uint64_t)j);
        fe omega_step = fr_pow_u64(c->omega, (uint64_t)(j << c->log_new_n));
        fe elt = fe_one(&FR);
        for (size_t i = 0; i < new_n; i++) {
            fe acc = fe_zero();
            for (size_t s = 0; s < num_cpus; s++) {
                size_t idx = (i + (s << c->log_new_n)) & mask;
                fe t = fe_mul(&FR, &c->a[idx], &elt);
                acc = fe_add(&FR, &acc, &t);
                elt = fe_mul(&FR, &elt, &omega_step);
            }
            tmp[i] = acc;
            elt = fe_mul(&FR, &elt, &omega_j);
        }
        serial_fft(tmp, new_n, &c->new_omega, c->log_new_n);
    }
}
static void pfft_stage2(void *p, size_t lo, size_t hi, int tid) {
    (void)tid; pfft_t *c = p; size_t mask = ((size_t)1 << c->log_cpus) - 1;
    for (size_t idx = lo; idx < hi; idx++) c->a[idx] = c->tmp[idx & mask][idx >> c->log_cpus];
}
static void parallel_fft(fe *a, size_t n, const fe *omega, int log_n, int log_cpus, int cpus) {  /* domain.rs parallel_fft */
    pfft_t c; c.a = a; c.omega = omega; c.log_n = log_n; c.log_cpus = log_cpus; c.log_new_n = log_n - log_cpus;
    size_t num_cpus = (size_t)1 << log_cpus, new_n = (size_t)1 << c.log_new_n;
    c.tmp = malloc(sizeof(fe *) * num_cpus);
    for (size_t j = 0; j < num_cpus; j++) c.tmp[j] = malloc(sizeof(fe) * new_n);
    c.new_omega = fr_pow_u64(omega, (uint64_t)num_cpus);
    worker_scope((int)num_cpus, num_cpus, pfft_stage1, &c);
    worker_scope(cpus, n, pfft_stage2, &c);
    for (size_t j = 0; j < num_cpus; j++) free(c.tmp[j]);
    free(c.tmp);
}
static void best_fft(fe *a, size_t n, const fe *omega, int log_n, int cpus) {   /* domain.rs best_fft */
    int log_cpus = log2_floor((unsigned)(cpus < 1 ? 1 : cpus));
    if (log_n <= log_cpus) serial_fft(a, n, omega, log_n);
    else parallel_fft(a, n, omega, log_n, log_cpus, cpus);
}
typedef struct { fe *a; fe k; fe g; int mode; const fe *b; } pw_t;
static void pw_fn(void *p, size_t lo, size_t hi, int tid) {
    (void)tid; pw_t *c = p;
    switch (c->mode) {
    case 0: for (size_t i = lo; i < hi; i++) c->a[i] = fe_mul(&FR, &c->a[i], &c->k); break;               /* scale */
    case 1: { fe u = fr_pow_u64(&c->g, (uint64_t)lo);                                                     /* distribute_powers */
              for (size_t i = lo; i < hi; i++) { c->a[i] = fe_mul(&FR, &c->a[i], &u); u = fe_mul(&FR, &u, &c->g); } break; }
    case 2: for (size_t i = lo; i < hi; i++) c->a[i] = fe_mul(&FR, &c->a[i], &c->b[i]); break;            /* mul_assign */
    default: for (size_t i = lo; i < hi; i++) c->a[i] = fe_sub(&FR, &c->a[i], &c->b[i]); break;           /* sub_assign */
    }
}
static void dom_fft(domain_t *d, int cpus) { best_fft(d->coeffs, d->m, &d->omega, d->exp, cpus); }
static void dom_ifft(domain_t *d, int cpus) {
    best_fft(d->coeffs, d->m, &d->omegainv, d->exp, cpus);
    pw_t c = {d->coeffs, d->minv, d->minv, 0, NULL}; worker_scope(cpus, d->m, pw_fn, &c);
}
static void dom_distribute_powers(domain_t *d, const fe *g, int cpus) { pw_t c = {d->coeffs, *g, *g, 1, NULL}; worker_scope(cpus, d->m, pw_fn, &c); }
static void dom_coset_fft(domain_t *d, int cpus) { fe g = fe_from_u64(&FR, FR_GENERATOR); dom_distribute_powers(d, &g, cpus); dom_fft(d, cpus); }
static void dom_icoset_fft(domain_t *d, int cpus) { dom_ifft(d, cpus); dom_distribute_powers(d, &d->geninv, cpus); }
static fe dom_z(const domain_t *d, const fe *tau) { fe t = fr_pow_u64(tau, (uint64_t)d->m); fe o = fe_one(&FR); return fe_sub(&FR, &t, &o); }
static void dom_divide_by_z_on_coset(domain_t *d, int cpus) {
    fe g = fe_from_u64(&FR, FR_GENERATOR); fe i = dom_z(d, &g); i = fe_inv(&FR, &i);
    pw_t c = {d->coeffs, i, i, 0, NULL}; worker_scope(cpus, d->m, pw_fn, &c);
}
static void dom_mul_assign(domain_t *d, const domain_t *o, int cpus) { pw_t c = {d->coeffs, d->minv, d->minv, 2, o->coeffs}; worker_scope(cpus, d->m, pw_fn, &c); }
static void dom_sub_assign(domain_t *d, const domain_t *o, int cpus) { pw_t c = {d->coeffs, d->minv, d->minv, 3, o->coeffs}; worker_scope(cpus, d->m, pw_fn, &c); }

static fe *fr_load_vec(const uint8_t *p, size_t n) { fe *v = malloc(sizeof(fe) * (n ? n : 1)); for (size_t i = 0; i < n; i++) v[i] = fr_load(p + 32 * i); return v; }
static void fr_store_vec(uint8_t *p, const fe *v, size_t n) { for (size_t i = 0; i < n; i++) fr_store(p + 32 * i, &v[i]); }

void ora_domain_omega(int log_n, uint8_t *omega_out) {
    fe omega = fe_from_canon(&FR, FR_ROOT_OF_UNITY);
    for (int i = log_n; i < FR_S; i++) omega = fe_sqr(&FR, &omega);
    fr_store(omega_out, &omega);
}
int ora_fft(uint8_t *data, int log_n, int mode, int threads) {
    size_t n = (size_t)1 << log_n; domain_t d; int e = domain_init(&d, fr_load_vec(data, n), n);
    if (e) return e;
    switch (mode) { case 0: dom_fft(&d, threads); break; case 1: dom_ifft(&d, threads); break;
                    case 2: dom_coset_fft(&d, threads); break; default: dom_icoset_fft(&d, threads); break; }
    fr_store_vec(data, d.coeffs, n); free(d.coeffs); return 0;
}

/* create_proof's H computation (groth16/prover.rs, `let h = { ... }` block), SURVEY §3.2 step 4.
 * a, b, c are consumed (freed).  Returns the m-1 coefficients as Montgomery Fr in *h_out (malloc). */
static int h_poly(fe *a_own, fe *b_own, fe *c_own, size_t len, fe **h_out, size_t *h_len, uint8_t *ckpt, int cpus) {
    domain_t a, b, c; int e;
    if ((e = domain_init(&a, a_own, len)) || (e = domain_init(&b, b_own, len)) || (e = domain_init(&c, c_own, len))) return e;
    size_t m = a.m;
#define CK(k, d) do { if (ckpt) fr_store_vec(ckpt + (size_t)(k) * m * 32, (d).coeffs, m); } while (0)
    dom_ifft(&a, cpus); CK(0, a); dom_coset_fft(&a, cpus); CK(1, a);
    dom_ifft(&b, cpus); CK(2, b); dom_coset_fft(&b, cpus); CK(3, b);
    dom_ifft(&c, cpus); CK(4, c); dom_coset_fft(&c, cpus); CK(5, c);
    dom_mul_assign(&a, &b, cpus); free(b.coeffs);
    dom_sub_assign(&a, &c, cpus); free(c.coeffs);
    dom_divide_by_z_on_coset(&a, cpus); CK(6, a);
    dom_icoset_fft(&a, cpus); CK(7, a);
#undef CK
    *h_out = a.coeffs; *h_len = m - 1;      /* into_coeffs(); truncate(len - 1) */
    return 0;
}
int ora_h_poly(const uint8_t *a, const uint8_t *b, const uint8_t *c, size_t len, uint8_t *h_out, uint8_t *checkpoints, int threads) {
    fe *h; size_t hl;
    int e = h_poly(fr_load_vec(a, len), fr_load_vec(b, len), fr_load_vec(c, len), len, &h, &hl, checkpoints, threads);
    if (e) return e;
    fr_store_vec(h_out, h, hl); free(h); return 0;
}

/* ---------------------------------------------------------------- multiexp
 * bellman_ce multiexp.rs `multiexp` / `multiexp_inner`: unsigned c-bit windows,
 * c = 3 if n < 32 else ceil(ln n); one task per window scanning all exponents;
 * exp == 0 skipped, exp == 1 added straight to acc in window 0; density-filtered
 * base cursor; running-sum bucket reduction; windows joined with c doublings. */
typedef uint64_t repr_t[4];
static unsigned multiexp_c(size_t n) { return n < 32 ? 3u : (unsigned)ceil(log((double)(uint32_t)n)); }
static uint64_t repr_window(const uint64_t *e, unsigned skip, unsigned c) {     /* (exp >> skip) mod 2^c */
    unsigned limb = skip >> 6, sh = skip & 63; uint64_t v = limb < 4 ? e[limb] >> sh : 0;
    if (sh && limb + 1 < 4) v |= e[limb + 1] << (64 - sh);
    return v & (((uint64_t)1 << c) - 1);
}
static int repr_is_one(const uint64_t *e) { return e[0] == 1 && !e[1] && !e[2] && !e[3]; }
static int repr_is_zero(const uint64_t *e) { return !(e[0] | e[1] | e[2] | e[3]); }

#define MULTIEXP_IMPL(G)                                                                                          \
    typedef struct { const G##_affine *bases; size_t n_bases; const repr_t *exps; size_t n_exp;                   \
                     const uint8_t *density; unsigned c; G##_jac *results; int *errs; } G##_me_t;                 \
    static void G##_me_window(void *p, size_t lo, size_t hi, int tid) {                                           \
        (void)tid; G##_me_t *m = p;                                                                               \
        for (size_t w = lo; w < hi; w++) {                                                                        \
            unsigned skip = (unsigned)w * m->c; int handle_trivial = (w == 0);                                    \
            G##_jac acc = G##_jac_zero(); size_t nb = ((size_t)1 << m->c) - 1;                                    \
            G##_jac *buckets = malloc(sizeof(G##_jac) * nb);                                                      \
            for (size_t i = 0; i < nb; i++) buckets[i] = G##_jac_zero();                                          \
            size_t cur = 0; int err = 0;                                                                          \
            for (size_t i = 0; i < m->n_exp && !err; i++) {                                                       \
                if (m->density && !m->density[i]) continue;                                                       \
                if (cur >= m->n_bases) { err = ORA_ERR_IO; break; }                                               \
                const G##_affine *base = &m->bases[cur++];                                                        \
                const uint64_t *e = m->exps[i];                                                                   \
                if (repr_is_zero(e)) continue;                                                                    \
                if (repr_is_one(e)) {                                                                             \
                    if (handle_trivial) { if (base->inf) { err = ORA_ERR_UNEXPECTED_IDENTITY; break; } G##_add_mixed(&acc, base); } \
                    continue;                                                                                     \
                }                                                                                                 \
                uint64_t d = repr_window(e, skip, m->c);                                                          \
                if (d) { if (base->inf) { err = ORA_ERR_UNEXPECTED_IDENTITY; break; } G##_add_mixed(&buckets[d - 1], base); } \
            }                                                                                                     \
            G##_jac running = G##_jac_zero();                                                                     \
            for (size_t i = nb; i-- > 0;) { G##_add(&running, &buckets[i]); G##_add(&acc, &running); }            \
            free(buckets); m->results[w] = acc; m->errs[w] = err;                                                 \
        }                                                                                                         \
    }                                                                                                             \
    static int G##_multiexp(const G##_affine *bases, size_t n_bases, const repr_t *exps, size_t n_exp,            \
                            const uint8_t *density, G##_jac *out, int cpus) {                                     \
        G##_me_t m; m.bases = bases; m.n_bases = n_bases; m.exps = exps; m.n_exp = n_exp; m.density = density;    \
        m.c = multiexp_c(n_exp);                                                                                  \
        size_t nw = 0; for (unsigned skip = 0; ; ) { nw++; skip += m.c; if (skip >= FR_NUM_BITS) break; }         \
        m.results = malloc(sizeof(G##_jac) * nw); m.errs = malloc(sizeof(int) * nw);                              \
        /* one pool task per window (bellman spawns them all on a cpus-sized pool) */                            \
        if (cpus <= 1) G##_me_window(&m, 0, nw, 0);                                                               \
        else {                                                                                                    \
            size_t done = 0;                                                                                      \
            while (done < nw) { /* waves of `cpus` single-window tasks */                                         \
                size_t wave = nw - done < (size_t)cpus ? nw - done : (size_t)cpus;                                \
                pthread_t th[256]; job_t jobs[256];                                                               \
                for (size_t j = 0; j < wave; j++) { jobs[j].fn = G##_me_window; jobs[j].arg = &m; jobs[j].lo = done + j; jobs[j].hi = done + j + 1; jobs[j].tid = (int)j; pthread_create(&th[j], NULL, job_tramp, &jobs[j]); } \
                for (size_t j = 0; j < wave; j++) pthread_join(th[j], NULL);                                      \
                done += wave;                                                                                     \
            }                                                                                                     \
        }                                                                                                         \
        int err = 0; for (size_t w = 0; w < nw; w++) if (m.errs[w]) err = m.errs[w];                              \
        G##_jac acc = m.results[nw - 1];                                                                          \
        for (size_t w = nw - 1; w-- > 0;) { for (unsigned k = 0; k < m.c; k++) G##_double(&acc); G##_add(&acc, &m.results[w]); } \
        free(m.results); free(m.errs); *out = acc; return err;                                                    \
    }
MULTIEXP_IMPL(g1)
MULTIEXP_IMPL(g2)

static repr_t *repr_load_vec(const uint8_t *p, size_t n) { repr_t *v = malloc(sizeof(repr_t) * (n ? n : 1)); memcpy(v, p, 32 * n); return v; }
int ora_multiexp_g1(const uint8_t *bases, size_t n_bases, const uint8_t *scalars, size_t n_exp, const uint8_t *density, uint8_t *out, int threads) {
    g1_affine *b = malloc(sizeof(g1_affine) * (n_bases ? n_bases : 1));
    for (size_t i = 0; i < n_bases; i++) b[i] = g1_load(bases + 64 * i);
    repr_t *e = repr_load_vec(scalars, n_exp); g1_jac r; int err = g1_multiexp(b, n_bases, (const repr_t *)e, n_exp, density, &r, threads);
    g1_affine a = g1_into_affine(&r); g1_store(out, &a); free(b); free(e); return err;
}
int ora_multiexp_g2(const uint8_t *bases, size_t n_bases, const uint8_t *scalars, size_t n_exp, const uint8_t *density, uint8_t *out, int threads) {
    g2_affine *b = malloc(sizeof(g2_affine) * (n_bases ? n_bases : 1));
    for (size_t i = 0; i < n_bases; i++) b[i] = g2_load(bases + 128 * i);
    repr_t *e = repr_load_vec(scalars, n_exp); g2_jac r; int err = g2_multiexp(b, n_bases, (const repr_t *)e, n_exp, density, &r, threads);
    g2_affine a = g2_into_affine(&r); g2_store(out, &a); free(b); free(e); return err;
}

/* ------------------------------------------------------------- Parameters */
typedef struct {
    g1_affine alpha_g1, beta_g1; g2_affine beta_g2, gamma_g2; g1_affine delta_g1; g2_affine delta_g2;
    uint32_t n_ic; g1_affine *ic;
    uint32_t n_h, n_l, n_a, n_bg1, n_bg2;
    g1_affine *h, *l, *a, *b_g1; g2_affine *b_g2;
} params_t;

void ora_params_free(void *pp) {
    params_t *p = pp; if (!p) return;
    free(p->ic); free(p->h); free(p->l); free(p->a); free(p->b_g1); free(p->b_g2); free(p);
}
void ora_params_counts(const void *pp, uint32_t *c) {
    const params_t *p = pp; c[0] = p->n_ic; c[1] = p->n_h; c[2] = p->n_l; c[3] = p->n_a; c[4] = p->n_bg1; c[5] = p->n_bg2;
}

/* pairing_ce bn256 G1Uncompressed / G2Uncompressed: big-endian canonical coordinates,
 * infinity = bit 6 of byte 0, bit 7 (compressed flag) must be clear. */
static void be_store(uint8_t *p, const fe *a) { uint64_t c[4]; fe_to_canon(&FQ, a, c); for (int i = 0; i < 32; i++) p[i] = (uint8_t)(c[3 - i / 8] >> (56 - 8 * (i % 8))); }
static int be_load(const uint8_t *p, fe *out, int mask_top) {
    uint64_t c[4] = {0, 0, 0, 0};
    for (int i = 0; i < 32; i++) { uint8_t b = p[i]; if (i == 0 && mask_top) b &= 0x3f; c[3 - i / 8] |= (uint64_t)b << (56 - 8 * (i % 8)); }
    if (!canon_lt_mod(&FQ, c)) return ORA_ERR_BAD_ENCODING;
    *out = fe_from_canon(&FQ, c); return 0;
}
static void g1_write_be(uint8_t *p, const g1_affine *a) {
    if (a->inf) { memset(p, 0, 64); p[0] = 0x40; return; }
    be_store(p, &a->x); be_store(p + 32, &a->y);
}
static void g2_write_be(uint8_t *p, const g2_affine *a) {
    if (a->inf) { memset(p, 0, 128); p[0] = 0x40; return; }
    be_store(p, &a->x.c1); be_store(p + 32, &a->x.c0); be_store(p + 64, &a->y.c1); be_store(p + 96, &a->y.c0);
}
static const uint64_t FR_MODULUS_REPR[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static int g1_read_be(const uint8_t *p, g1_affine *a, int checked) {
    if (p[0] & 0x80) return ORA_ERR_BAD_ENCODING;
    if (p[0] & 0x40) { if ((p[0] & 0x3f) || !all_zero(p + 1, 63)) return ORA_ERR_BAD_ENCODING; *a = g1_affine_zero(); return 0; }
    int e; if ((e = be_load(p, &a->x, 1)) || (e = be_load(p + 32, &a->y, 0))) return e;
    a->inf = 0;
    if (checked) { fe b = g1_b(); if (!g1_on_curve(a, &b)) return ORA_ERR_NOT_ON_CURVE; }  /* cofactor 1: subgroup check is trivially true */
    return 0;
}
static int g2_read_be(const uint8_t *p, g2_affine *a, int checked) {
    if (p[0] & 0x80) return ORA_ERR_BAD_ENCODING;
    if (p[0] & 0x40) { if ((p[0] & 0x3f) || !all_zero(p + 1, 127)) return ORA_ERR_BAD_ENCODING; *a = g2_affine_zero(); return 0; }
    int e; if ((e = be_load(p, &a->x.c1, 1)) || (e = be_load(p + 32, &a->x.c0, 0)) || (e = be_load(p + 64, &a->y.c1, 0)) || (e = be_load(p + 96, &a->y.c0, 0))) return e;
    a->inf = 0;
    if (checked) {
        fe2 b = g2_b(); if (!g2_on_curve(a, &b)) return ORA_ERR_NOT_ON_CURVE;
        g2_jac j = g2_from_affine(a); g2_jac t = g2_mul(&j, FR_MODULUS_REPR);       /* is_in_correct_subgroup_assuming_on_curve */
        if (!g2_jac_is_zero(&t)) return ORA_ERR_NOT_IN_SUBGROUP;
    }
    return 0;
}
static void be32_store(uint8_t *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
static uint32_t be32_load(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

size_t ora_params_size(const void *pp) {
    const params_t *p = pp;
    return 64 + 64 + 128 + 128 + 64 + 128 + 4 + 64 * (size_t)p->n_ic + 4 + 64 * (size_t)p->n_h + 4 + 64 * (size_t)p->n_l +
           4 + 64 * (size_t)p->n_a + 4 + 64 * (size_t)p->n_bg1 + 4 + 128 * (size_t)p->n_bg2;
}
/* bellman Parameters::write = VerifyingKey::write then h, l, a, b_g1, b_g2 with u32 BE counts (format.rs:250) */
int ora_params_write(const void *pp, uint8_t *buf) {
    const params_t *p = pp; uint8_t *w = buf;
    g1_write_be(w, &p->alpha_g1); w += 64; g1_write_be(w, &p->beta_g1); w += 64; g2_write_be(w, &p->beta_g2); w += 128;
    g2_write_be(w, &p->gamma_g2); w += 128; g1_write_be(w, &p->delta_g1); w += 64; g2_write_be(w, &p->delta_g2); w += 128;
    be32_store(w, p->n_ic); w += 4; for (uint32_t i = 0; i < p->n_ic; i++, w += 64) g1_write_be(w, &p->ic[i]);
    be32_store(w, p->n_h); w += 4; for (uint32_t i = 0; i < p->n_h; i++, w += 64) g1_write_be(w, &p->h[i]);
    be32_store(w, p->n_l); w += 4; for (uint32_t i = 0; i < p->n_l; i++, w += 64) g1_write_be(w, &p->l[i]);
    be32_store(w, p->n_a); w += 4; for (uint32_t i = 0; i < p->n_a; i++, w += 64) g1_write_be(w, &p->a[i]);
    be32_store(w, p->n_bg1); w += 4; for (uint32_t i = 0; i < p->n_bg1; i++, w += 64) g1_write_be(w, &p->b_g1[i]);
    be32_store(w, p->n_bg2); w += 4; for (uint32_t i = 0; i < p->n_bg2; i++, w += 128) g2_write_be(w, &p->b_g2[i]);
    return 0;
}
/* bellman Parameters::read(reader, checked) (format.rs:285): points at infinity are rejected in h,l,a,b queries */
void *ora_params_read(const uint8_t *buf, size_t len, int checked, int *err) {
    params_t *p = calloc(1, sizeof(params_t)); const uint8_t *r = buf, *end = buf + len; int e = 0;
#define NEED(n) do { if ((size_t)(end - r) < (size_t)(n)) { e = ORA_ERR_IO; goto fail; } } while (0)
#define RG1(dst) do { NEED(64); if ((e = g1_read_be(r, dst, checked))) goto fail; r += 64; } while (0)
#define RG2(dst) do { NEED(128); if ((e = g2_read_be(r, dst, checked))) goto fail; r += 128; } while (0)
    RG1(&p->alpha_g1); RG1(&p->beta_g1); RG2(&p->beta_g2); RG2(&p->gamma_g2); RG1(&p->delta_g1); RG2(&p->delta_g2);
    NEED(4); p->n_ic = be32_load(r); r += 4; NEED(64 * (size_t)p->n_ic); p->ic = malloc(sizeof(g1_affine) * (p->n_ic + 1));
    for (uint32_t i = 0; i < p->n_ic; i++) RG1(&p->ic[i]);
#define RVEC1(cnt, arr) do { NEED(4); cnt = be32_load(r); r += 4; NEED(64 * (size_t)cnt); arr = malloc(sizeof(g1_affine) * ((size_t)cnt + 1)); \
        for (uint32_t i = 0; i < cnt; i++) { RG1(&arr[i]); if (arr[i].inf) { e = ORA_ERR_UNEXPECTED_IDENTITY; goto fail; } } } while (0)
    RVEC1(p->n_h, p->h); RVEC1(p->n_l, p->l); RVEC1(p->n_a, p->a); RVEC1(p->n_bg1, p->b_g1);
    NEED(4); p->n_bg2 = be32_load(r); r += 4; NEED(128 * (size_t)p->n_bg2); p->b_g2 = malloc(sizeof(g2_affine) * ((size_t)p->n_bg2 + 1));
    for (uint32_t i = 0; i < p->n_bg2; i++) { RG2(&p->b_g2[i]); if (p->b_g2[i].inf) { e = ORA_ERR_UNEXPECTED_IDENTITY; goto fail; } }
    if (err) *err = 0;
    return p;
fail:
    if (err) *err = e;
    ora_params_free(p); return NULL;
}

/* Synthetic Parameters (bench + full-size property tests; mirrors the definition in include/za_b200.h
 * za_pk_synthetic — written independently): query q entry i = ((q+1) * 2^32 + i + 1) * G,
 * alpha, beta, gamma, delta = 3, 5, 7, 11, ic[i] = 13 + i.  Not a valid CRS. */
typedef struct { g1_affine *out1; g2_affine *out2; uint64_t first; } chain_t;
static void chain_g1_fn(void *p, size_t lo, size_t hi, int tid) {
    (void)tid; chain_t *c = p; if (lo >= hi) return;
    uint8_t gen[64], tmp[64], k[32] = {0}; ora_g1_generator(gen);
    uint64_t f = c->first + lo; memcpy(k, &f, 8);
    ora_g1_mul(gen, k, tmp);                                   /* (first+lo) * G */
    g1_affine g = g1_load(gen), cur_a = g1_load(tmp); g1_jac acc = g1_from_affine(&cur_a);
    g1_jac *buf = malloc(sizeof(g1_jac) * MULT_BATCH); fe *pre = malloc(sizeof(fe) * MULT_BATCH);
    for (size_t s = lo; s < hi; s += MULT_BATCH) {
        size_t m = hi - s < MULT_BATCH ? hi - s : MULT_BATCH;
        for (size_t i = 0; i < m; i++) { buf[i] = acc; g1_add_mixed(&acc, &g); }
        fe run = fe_one(&FQ);
        for (size_t i = 0; i < m; i++) { pre[i] = run; run = fe_mul(&FQ, &run, &buf[i].z); }
        run = fe_inv(&FQ, &run);
        for (size_t i = m; i-- > 0;) {
            fe zi = fe_mul(&FQ, &run, &pre[i]); run = fe_mul(&FQ, &run, &buf[i].z);
            fe zi2 = fe_sqr(&FQ, &zi), zi3 = fe_mul(&FQ, &zi2, &zi);
            g1_affine a; a.x = fe_mul(&FQ, &buf[i].x, &zi2); a.y = fe_mul(&FQ, &buf[i].y, &zi3); a.inf = 0;
            c->out1[s + i] = a;
        }
    }
    free(buf); free(pre);
}
static void chain_g2_fn(void *p, size_t lo, size_t hi, int tid) {
    (void)tid; chain_t *c = p; if (lo >= hi) return;
    uint8_t gen[128], tmp[128], k[32] = {0}; ora_g2_generator(gen);
    uint64_t f = c->first + lo; memcpy(k, &f, 8);
    ora_g2_mul(gen, k, tmp);
    g2_affine g = g2_load(gen), cur_a = g2_load(tmp); g2_jac acc = g2_from_affine(&cur_a);
    g2_jac *buf = malloc(sizeof(g2_jac) * MULT_BATCH); fe2 *pre = malloc(sizeof(fe2) * MULT_BATCH);
    for (size_t s = lo; s < hi; s += MULT_BATCH) {
        size_t m = hi - s < MULT_BATCH ? hi - s : MULT_BATCH;
        for (size_t i = 0; i < m; i++) { buf[i] = acc; g2_add_mixed(&acc, &g); }
        fe2 run = fe2_one();
        for (size_t i = 0; i < m; i++) { pre[i] = run; run = fe2_mul(&run, &buf[i].z); }
        run = fe2_inv(&run);
        for (size_t i = m; i-- > 0;) {
            fe2 zi = fe2_mul(&run, &pre[i]); run = fe2_mul(&run, &buf[i].z);
            fe2 zi2 = fe2_sqr(&zi), zi3 = fe2_mul(&zi2, &zi);
            g2_affine a; a.x = fe2_mul(&buf[i].x, &zi2); a.y = fe2_mul(&buf[i].y, &zi3); a.inf = 0;
            c->out2[s + i] = a;
        }
    }
    free(buf); free(pre);
}
static g1_affine g1_small_multiple(uint64_t k) {
    uint8_t gen[64], tmp[64], kb[32] = {0}; ora_g1_generator(gen); memcpy(kb, &k, 8); ora_g1_mul(gen, kb, tmp); return g1_load(tmp);
}
static g2_affine g2_small_multiple(uint64_t k) {
    uint8_t gen[128], tmp[128], kb[32] = {0}; ora_g2_generator(gen); memcpy(kb, &k, 8); ora_g2_mul(gen, kb, tmp); return g2_load(tmp);
}
void *ora_params_synthetic(const uint32_t *counts, int threads) {
    params_t *p = calloc(1, sizeof(params_t));
    p->alpha_g1 = g1_small_multiple(3); p->beta_g1 = g1_small_multiple(5); p->delta_g1 = g1_small_multiple(11);
    p->beta_g2 = g2_small_multiple(5); p->gamma_g2 = g2_small_multiple(7); p->delta_g2 = g2_small_multiple(11);
    p->n_ic = counts[0]; p->ic = malloc(sizeof(g1_affine) * ((size_t)counts[0] + 1));
    for (uint32_t i = 0; i < counts[0]; i++) p->ic[i] = g1_small_multiple(13 + i);
    p->n_h = counts[1]; p->n_l = counts[2]; p->n_a = counts[3]; p->n_bg1 = counts[4]; p->n_bg2 = counts[5];
    p->h = malloc(sizeof(g1_affine) * ((size_t)p->n_h + 1)); p->l = malloc(sizeof(g1_affine) * ((size_t)p->n_l + 1));
    p->a = malloc(sizeof(g1_affine) * ((size_t)p->n_a + 1)); p->b_g1 = malloc(sizeof(g1_affine) * ((size_t)p->n_bg1 + 1));
    p->b_g2 = malloc(sizeof(g2_affine) * ((size_t)p->n_bg2 + 1));
    chain_t c; c.out2 = NULL;
    c.out1 = p->h; c.first = ((uint64_t)1 << 32) + 1; worker_scope(threads, p->n_h, chain_g1_fn, &c);
    c.out1 = p->l; c.first = ((uint64_t)2 << 32) + 1; worker_scope(threads, p->n_l, chain_g1_fn, &c);
    c.out1 = p->a; c.first = ((uint64_t)3 << 32) + 1; worker_scope(threads, p->n_a, chain_g1_fn, &c);
    c.out1 = p->b_g1; c.first = ((uint64_t)4 << 32) + 1; worker_scope(threads, p->n_bg1, chain_g1_fn, &c);
    c.out1 = NULL; c.out2 = p->b_g2; c.first = ((uint64_t)4 << 32) + 1; worker_scope(threads, p->n_bg2, chain_g2_fn, &c);
    return p;
}

/* ------------------------------------------------- ProvingAssignment::eval
 * bellman groth16/prover.rs `eval`: acc += coeff * value in insertion order, coeff == 1 skips the
 * multiply, density.inc(i) for every term regardless of coefficient or value. */
static fe eval_lc(const ora_r1cs *cs, int which, uint32_t row, const fe *inputs, const fe *aux,
                  uint8_t *input_density, uint8_t *aux_density) {
    fe acc = fe_zero(), one = fe_one(&FR);
    for (uint32_t t = cs->ptr[which][row]; t < cs->ptr[which][row + 1]; t++) {
        uint32_t v = cs->var[which][t]; fe tmp;
        if (v & ORA_VAR_AUX) { uint32_t i = v & ~ORA_VAR_AUX; tmp = aux[i]; if (aux_density) aux_density[i] = 1; }
        else { tmp = inputs[v]; if (input_density) input_density[v] = 1; }
        fe coeff = fr_load(cs->coeff[which] + 32 * (size_t)t);
        if (fe_eq(&coeff, &one)) acc = fe_add(&FR, &acc, &tmp);
        else { tmp = fe_mul(&FR, &tmp, &coeff); acc = fe_add(&FR, &acc, &tmp); }
    }
    return acc;
}

/* bellman groth16/prover.rs `create_proof` with explicit r, s. */
int ora_create_proof(const void *pp, const ora_r1cs *cs, const uint8_t *inputs_b, const uint8_t *aux_b,
                     const uint8_t *r_b, const uint8_t *s_b, uint8_t *proof, ora_trace *tr, int threads) {
    const params_t *p = pp; int err = 0;
    uint32_t ni = cs->num_inputs, na = cs->num_aux, nc = cs->num_constraints;
    fe *inputs = fr_load_vec(inputs_b, ni), *aux = fr_load_vec(aux_b, na);
    size_t len = (size_t)nc + ni;
    fe *a = malloc(sizeof(fe) * len), *b = malloc(sizeof(fe) * len), *c = malloc(sizeof(fe) * len);
    uint8_t *a_aux_d = calloc(na + 1, 1), *b_in_d = calloc(ni + 1, 1), *b_aux_d = calloc(na + 1, 1);
    /* steps 2-3: circuit rows, then one input-consistency row per input: enforce(A = input_i, B = 0, C = 0) */
    for (uint32_t k = 0; k < nc; k++) {
        a[k] = eval_lc(cs, 0, k, inputs, aux, NULL, a_aux_d);
        b[k] = eval_lc(cs, 1, k, inputs, aux, b_in_d, b_aux_d);
        c[k] = eval_lc(cs, 2, k, inputs, aux, NULL, NULL);
    }
    for (uint32_t i = 0; i < ni; i++) { a[nc + i] = inputs[i]; b[nc + i] = fe_zero(); c[nc + i] = fe_zero(); }
    if (tr) {
        if (tr->a_eval) fr_store_vec(tr->a_eval, a, len);
        if (tr->b_eval) fr_store_vec(tr->b_eval, b, len);
        if (tr->c_eval) fr_store_vec(tr->c_eval, c, len);
        if (tr->a_aux_density) memcpy(tr->a_aux_density, a_aux_d, na);
        if (tr->b_input_density) memcpy(tr->b_input_density, b_in_d, ni);
        if (tr->b_aux_density) memcpy(tr->b_aux_density, b_aux_d, na);
    }
    /* step 4: H */
    fe *h; size_t hl;
    if ((err = h_poly(a, b, c, len, &h, &hl, NULL, threads))) goto out0;
    repr_t *h_repr = malloc(sizeof(repr_t) * (hl ? hl : 1));
    for (size_t i = 0; i < hl; i++) fe_to_canon(&FR, &h[i], h_repr[i]);
    if (tr && tr->h_coeffs) memcpy(tr->h_coeffs, h_repr, 32 * hl);
    free(h);
    repr_t *in_repr = repr_load_vec(inputs_b, ni), *aux_repr = repr_load_vec(aux_b, na);
    g1_jac H, L, A_in, A_aux, B1_in, B1_aux; g2_jac B2_in, B2_aux;
    size_t a_aux_total = 0, b_in_total = 0, b_aux_total = 0;
    for (uint32_t i = 0; i < na; i++) { a_aux_total += a_aux_d[i]; b_aux_total += b_aux_d[i]; }
    for (uint32_t i = 0; i < ni; i++) b_in_total += b_in_d[i];
    /* ParameterSource for &Parameters: get_h(len) -> (h,0) etc.; EOF if a query is too short */
#define TRY(x) do { int e_ = (x); if (e_ && !err) err = e_; } while (0)
    TRY(g1_multiexp(p->h, p->n_h, (const repr_t *)h_repr, hl, NULL, &H, threads));
    TRY(g1_multiexp(p->l, p->n_l, (const repr_t *)aux_repr, na, NULL, &L, threads));
    if (p->n_a < ni) { err = ORA_ERR_IO; goto out1; }
    TRY(g1_multiexp(p->a, p->n_a, (const repr_t *)in_repr, ni, NULL, &A_in, threads));
    TRY(g1_multiexp(p->a + ni, p->n_a - ni, (const repr_t *)aux_repr, na, a_aux_d, &A_aux, threads));
    if (p->n_bg1 < b_in_total || p->n_bg2 < b_in_total) { err = ORA_ERR_IO; goto out1; }
    TRY(g1_multiexp(p->b_g1, p->n_bg1, (const repr_t *)in_repr, ni, b_in_d, &B1_in, threads));
    TRY(g1_multiexp(p->b_g1 + b_in_total, p->n_bg1 - b_in_total, (const repr_t *)aux_repr, na, b_aux_d, &B1_aux, threads));
    TRY(g2_multiexp(p->b_g2, p->n_bg2, (const repr_t *)in_repr, ni, b_in_d, &B2_in, threads));
    TRY(g2_multiexp(p->b_g2 + b_in_total, p->n_bg2 - b_in_total, (const repr_t *)aux_repr, na, b_aux_d, &B2_aux, threads));
    (void)a_aux_total; (void)b_aux_total;
    if (err) goto out1;
    if (tr && tr->msm_g1) {
        g1_jac *v[6] = {&H, &L, &A_in, &A_aux, &B1_in, &B1_aux};
        for (int i = 0; i < 6; i++) { g1_affine t = g1_into_affine(v[i]); g1_store(tr->msm_g1 + 64 * i, &t); }
    }
    if (tr && tr->msm_g2) {
        g2_affine t = g2_into_affine(&B2_in); g2_store(tr->msm_g2, &t);
        t = g2_into_affine(&B2_aux); g2_store(tr->msm_g2 + 128, &t);
    }
    /* steps 6-8 */
    if (p->delta_g1.inf || p->delta_g2.inf) { err = ORA_ERR_UNEXPECTED_IDENTITY; goto out1; }
    {
        uint64_t rr[4], ss[4], rs_c[4]; rd_le(r_b, rr); rd_le(s_b, ss);
        fe rf = fe_from_canon(&FR, rr), sf = fe_from_canon(&FR, ss), rs = fe_mul(&FR, &rf, &sf); fe_to_canon(&FR, &rs, rs_c);
        g1_jac d1 = g1_from_affine(&p->delta_g1), al = g1_from_affine(&p->alpha_g1), be1 = g1_from_affine(&p->beta_g1);
        g2_jac d2 = g2_from_affine(&p->delta_g2);
        g1_jac g_a = g1_mul(&d1, rr); g1_add_mixed(&g_a, &p->alpha_g1);
        g2_jac g_b = g2_mul(&d2, ss); g2_add_mixed(&g_b, &p->beta_g2);
        g1_jac g_c = g1_mul(&d1, rs_c); g1_jac t = g1_mul(&al, ss); g1_add(&g_c, &t); t = g1_mul(&be1, rr); g1_add(&g_c, &t);
        g1_jac a_ans = A_in; g1_add(&a_ans, &A_aux); g1_add(&g_a, &a_ans); a_ans = g1_mul(&a_ans, ss); g1_add(&g_c, &a_ans);
        g1_jac b1_ans = B1_in; g1_add(&b1_ans, &B1_aux);
        g2_jac b2_ans = B2_in; g2_add(&b2_ans, &B2_aux);
        g2_add(&g_b, &b2_ans); b1_ans = g1_mul(&b1_ans, rr); g1_add(&g_c, &b1_ans);
        g1_add(&g_c, &H); g1_add(&g_c, &L);
        g1_affine pa = g1_into_affine(&g_a), pc = g1_into_affine(&g_c); g2_affine pb = g2_into_affine(&g_b);
        g1_store(proof, &pa); g2_store(proof + 64, &pb); g1_store(proof + 192, &pc);
    }
out1:
    free(h_repr); free(in_repr); free(aux_repr);
    free(inputs); free(aux); free(a_aux_d); free(b_in_d); free(b_aux_d);
    return err;
out0:
    free(inputs); free(aux); free(a_aux_d); free(b_in_d); free(b_aux_d);
    return err;
}

/* ------------------------------------------------------ generate_parameters
 * bellman groth16/generator.rs `generate_parameters` with explicit toxic values and generators
 * (SURVEY A.5).  Plain double-and-add instead of the fixed-base wNAF tables: same group elements. */
typedef struct { uint32_t row; fe coeff; } kterm;
typedef struct { kterm *t; uint32_t n, cap; } klist;
static void klist_push(klist *l, uint32_t row, const fe *coeff) {
    if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 4; l->t = realloc(l->t, sizeof(kterm) * l->cap); }
    l->t[l->n].row = row; l->t[l->n].coeff = *coeff; l->n++;
}
typedef struct {
    const fe *lag; klist *at, *bt, *ct; g1_affine *a, *b1, *ext; g2_affine *b2;
    fe inv, alpha, beta; g1_jac g1; g2_jac g2;
} geval_t;
static fe eval_at_tau(const fe *lag, const klist *l) {
    fe acc = fe_zero();
    for (uint32_t i = 0; i < l->n; i++) { fe n = fe_mul(&FR, &lag[l->t[i].row], &l->t[i].coeff); acc = fe_add(&FR, &acc, &n); }
    return acc;
}
static void geval_fn(void *p, size_t lo, size_t hi, int tid) {
    (void)tid; geval_t *g = p;
    for (size_t i = lo; i < hi; i++) {
        fe at = eval_at_tau(g->lag, &g->at[i]), bt = eval_at_tau(g->lag, &g->bt[i]), ct = eval_at_tau(g->lag, &g->ct[i]);
        uint64_t k[4];
        if (!fe_is_zero(&at)) { fe_to_canon(&FR, &at, k); g1_jac t = g1_mul(&g->g1, k); g->a[i] = g1_into_affine(&t); } else g->a[i] = g1_affine_zero();
        if (!fe_is_zero(&bt)) {
            fe_to_canon(&FR, &bt, k); g1_jac t = g1_mul(&g->g1, k); g->b1[i] = g1_into_affine(&t);
            g2_jac t2 = g2_mul(&g->g2, k); g->b2[i] = g2_into_affine(&t2);
        } else { g->b1[i] = g1_affine_zero(); g->b2[i] = g2_affine_zero(); }
        at = fe_mul(&FR, &at, &g->beta); bt = fe_mul(&FR, &bt, &g->alpha);
        fe e = fe_add(&FR, &at, &bt); e = fe_add(&FR, &e, &ct); e = fe_mul(&FR, &e, &g->inv);
        fe_to_canon(&FR, &e, k); g1_jac t = g1_mul(&g->g1, k); g->ext[i] = g1_into_affine(&t);
    }
}
typedef struct { const fe *pow; fe coeff; g1_jac g1; g1_affine *h; } gh_t;
static void gh_fn(void *p, size_t lo, size_t hi, int tid) {
    (void)tid; gh_t *g = p;
    for (size_t i = lo; i < hi; i++) { fe e = fe_mul(&FR, &g->pow[i], &g->coeff); uint64_t k[4]; fe_to_canon(&FR, &e, k); g1_jac t = g1_mul(&g->g1, k); g->h[i] = g1_into_affine(&t); }
}

void *ora_generate_parameters(const ora_r1cs *cs, const uint8_t *alpha_b, const uint8_t *beta_b, const uint8_t *gamma_b,
                              const uint8_t *delta_b, const uint8_t *tau_b, const uint8_t *g1_b_, const uint8_t *g2_b_, int threads, int *err_out) {
    uint32_t ni = cs->num_inputs, na = cs->num_aux, nc = cs->num_constraints; int err = 0;
    fe alpha = fr_load(alpha_b), beta = fr_load(beta_b), gamma = fr_load(gamma_b), delta = fr_load(delta_b), tau = fr_load(tau_b);
    g1_affine g1a = g1_load(g1_b_); g2_affine g2a = g2_load(g2_b_);
    g1_jac g1 = g1_from_affine(&g1a); g2_jac g2 = g2_from_affine(&g2a);
    /* KeypairAssembly: per variable, (coeff, constraint index) lists */
    klist *at = calloc((size_t)ni + na, sizeof(klist)), *bt = calloc((size_t)ni + na, sizeof(klist)), *ct = calloc((size_t)ni + na, sizeof(klist));
    klist *lists[3] = {at, bt, ct};
    for (int w = 0; w < 3; w++)
        for (uint32_t k = 0; k < nc; k++)
            for (uint32_t t = cs->ptr[w][k]; t < cs->ptr[w][k + 1]; t++) {
                uint32_t v = cs->var[w][t]; size_t slot = (v & ORA_VAR_AUX) ? (size_t)ni + (v & ~ORA_VAR_AUX) : v;
                fe coeff = fr_load(cs->coeff[w] + 32 * (size_t)t); klist_push(&lists[w][slot], k, &coeff);
            }
    fe one = fe_one(&FR);
    for (uint32_t i = 0; i < ni; i++) klist_push(&at[i], nc + i, &one);           /* input-consistency rows */
    size_t ncons = (size_t)nc + ni;
    domain_t d; fe *pw = malloc(sizeof(fe) * (ncons ? ncons : 1));
    if ((err = domain_init(&d, pw, ncons))) { d.coeffs = NULL; free(pw); goto fail; }
    { fe cur = fe_one(&FR); for (size_t i = 0; i < d.m; i++) { d.coeffs[i] = cur; cur = fe_mul(&FR, &cur, &tau); } }
    params_t *p = calloc(1, sizeof(params_t));
    fe gamma_inv = fe_inv(&FR, &gamma), delta_inv = fe_inv(&FR, &delta);
    /* h[i] = g1 * (tau^i * z(tau) / delta), i < m-1 */
    p->n_h = (uint32_t)(d.m - 1); p->h = malloc(sizeof(g1_affine) * d.m);
    { gh_t g; g.pow = d.coeffs; g.coeff = dom_z(&d, &tau); g.coeff = fe_mul(&FR, &g.coeff, &delta_inv); g.g1 = g1; g.h = p->h;
      worker_scope(threads, d.m - 1, gh_fn, &g); }
    dom_ifft(&d, threads);                                                       /* Lagrange coefficients at tau */
    g1_affine *a = malloc(sizeof(g1_affine) * ((size_t)ni + na + 1)), *b1 = malloc(sizeof(g1_affine) * ((size_t)ni + na + 1));
    g2_affine *b2 = malloc(sizeof(g2_affine) * ((size_t)ni + na + 1));
    p->n_ic = ni; p->ic = malloc(sizeof(g1_affine) * (ni + 1)); p->n_l = na; p->l = malloc(sizeof(g1_affine) * (na + 1));
    geval_t g; g.lag = d.coeffs; g.alpha = alpha; g.beta = beta; g.g1 = g1; g.g2 = g2;
    g.at = at; g.bt = bt; g.ct = ct; g.a = a; g.b1 = b1; g.b2 = b2; g.ext = p->ic; g.inv = gamma_inv;
    worker_scope(threads, ni, geval_fn, &g);
    g.at = at + ni; g.bt = bt + ni; g.ct = ct + ni; g.a = a + ni; g.b1 = b1 + ni; g.b2 = b2 + ni; g.ext = p->l; g.inv = delta_inv;
    worker_scope(threads, na, geval_fn, &g);
    for (uint32_t i = 0; i < na; i++) if (p->l[i].inf) err = ORA_ERR_UNCONSTRAINED_VARIABLE;
    { uint64_t k[4]; g1_jac t; g2_jac t2;
      fe_to_canon(&FR, &alpha, k); t = g1_mul(&g1, k); p->alpha_g1 = g1_into_affine(&t);
      fe_to_canon(&FR, &beta, k); t = g1_mul(&g1, k); p->beta_g1 = g1_into_affine(&t); t2 = g2_mul(&g2, k); p->beta_g2 = g2_into_affine(&t2);
      fe_to_canon(&FR, &gamma, k); t2 = g2_mul(&g2, k); p->gamma_g2 = g2_into_affine(&t2);
      fe_to_canon(&FR, &delta, k); t = g1_mul(&g1, k); p->delta_g1 = g1_into_affine(&t); t2 = g2_mul(&g2, k); p->delta_g2 = g2_into_affine(&t2); }
    /* filter points at infinity out of a, b_g1, b_g2 (inputs first, then aux, order kept) */
    p->a = malloc(sizeof(g1_affine) * ((size_t)ni + na + 1)); p->b_g1 = malloc(sizeof(g1_affine) * ((size_t)ni + na + 1)); p->b_g2 = malloc(sizeof(g2_affine) * ((size_t)ni + na + 1));
    for (size_t i = 0; i < (size_t)ni + na; i++) {
        if (!a[i].inf) p->a[p->n_a++] = a[i];
        if (!b1[i].inf) p->b_g1[p->n_bg1++] = b1[i];
        if (!b2[i].inf) p->b_g2[p->n_bg2++] = b2[i];
    }
    free(a); free(b1); free(b2); free(d.coeffs);
    for (size_t i = 0; i < (size_t)ni + na; i++) { free(at[i].t); free(bt[i].t); free(ct[i].t); }
    free(at); free(bt); free(ct);
    if (err) { ora_params_free(p); p = NULL; }
    if (err_out) *err_out = err;
    return p;
fail:
    for (size_t i = 0; i < (size_t)ni + na; i++) { free(at[i].t); free(bt[i].t); free(ct[i].t); }
    free(at); free(bt); free(ct);
    if (err_out) *err_out = err;
    return NULL;
}

/* ------------------------------------------------------------------ pairing
 * Textbook optimal-ate pairing on BN254 (pairing_ce bn256 `Engine::{miller_loop, final_exponentiation}`
 * compute the same bilinear map up to a fixed exponent, which does not change the truth value of the
 * Groth16 check).  Tower: Fq6 = Fq2[v]/(v^3 - (9+u)), Fq12 = Fq6[w]/(w^2 - v)  (SURVEY A.1).
 * Lines are evaluated through the untwist (x', y') -> (x' w^2, y' w^3); the final exponentiation is a
 * plain square-and-multiply by (q^12 - 1)/r — slow and simple on purpose. */
typedef struct { fe2 a0, a1, a2; } fe6;
typedef struct { fe6 c0, c1; } fe12;
static fe2 fe2_mul_xi(const fe2 *a) {   /* (a0 + a1 u)(9 + u) */
    fe2 r; fe t = fe_dbl(&FQ, &a->c0); t = fe_dbl(&FQ, &t); t = fe_dbl(&FQ, &t); t = fe_add(&FQ, &t, &a->c0);   /* 9 a0 */
    r.c0 = fe_sub(&FQ, &t, &a->c1);
    t = fe_dbl(&FQ, &a->c1); t = fe_dbl(&FQ, &t); t = fe_dbl(&FQ, &t); t = fe_add(&FQ, &t, &a->c1);               /* 9 a1 */
    r.c1 = fe_add(&FQ, &t, &a->c0); return r;
}
static fe6 fe6_zero(void) { fe6 z; z.a0 = fe2_zero(); z.a1 = fe2_zero(); z.a2 = fe2_zero(); return z; }
static fe6 fe6_add(const fe6 *a, const fe6 *b) { fe6 r; r.a0 = fe2_add(&a->a0, &b->a0); r.a1 = fe2_add(&a->a1, &b->a1); r.a2 = fe2_add(&a->a2, &b->a2); return r; }
static fe6 fe6_mul(const fe6 *a, const fe6 *b) {
    fe2 t00 = fe2_mul(&a->a0, &b->a0), t01 = fe2_mul(&a->a0, &b->a1), t02 = fe2_mul(&a->a0, &b->a2);
    fe2 t10 = fe2_mul(&a->a1, &b->a0), t11 = fe2_mul(&a->a1, &b->a1), t12 = fe2_mul(&a->a1, &b->a2);
    fe2 t20 = fe2_mul(&a->a2, &b->a0), t21 = fe2_mul(&a->a2, &b->a1), t22 = fe2_mul(&a->a2, &b->a2);
    fe6 r; fe2 s = fe2_add(&t12, &t21); s = fe2_mul_xi(&s); r.a0 = fe2_add(&t00, &s);
    s = fe2_mul_xi(&t22); r.a1 = fe2_add(&t01, &t10); r.a1 = fe2_add(&r.a1, &s);
    r.a2 = fe2_add(&t02, &t11); r.a2 = fe2_add(&r.a2, &t20); return r;
}
static fe6 fe6_mul_v(const fe6 *a) { fe6 r; r.a0 = fe2_mul_xi(&a->a2); r.a1 = a->a0; r.a2 = a->a1; return r; }
static fe12 fe12_one(void) { fe12 r; r.c0 = fe6_zero(); r.c1 = fe6_zero(); r.c0.a0 = fe2_one(); return r; }
static fe12 fe12_mul(const fe12 *a, const fe12 *b) {
    fe6 t0 = fe6_mul(&a->c0, &b->c0), t1 = fe6_mul(&a->c1, &b->c1), t2 = fe6_mul(&a->c0, &b->c1), t3 = fe6_mul(&a->c1, &b->c0);
    fe12 r; fe6 v = fe6_mul_v(&t1); r.c0 = fe6_add(&t0, &v); r.c1 = fe6_add(&t2, &t3); return r;
}
static int fe12_eq(const fe12 *a, const fe12 *b) { return memcmp(a, b, sizeof(fe12)) == 0; }
static fe12 fe12_pow(const fe12 *a, const uint64_t *e, int nlimbs) {
    fe12 r = fe12_one(); int started = 0;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        if (started) r = fe12_mul(&r, &r);
        if ((e[i >> 6] >> (i & 63)) & 1) { r = started ? fe12_mul(&r, a) : *a; started = 1; }
    }
    return r;
}
static const uint64_t FINAL_EXP[44] = {0x86964b64ca86f120ULL, 0x40a4efb7e54523a4ULL, 0x837fa97896e84abbULL, 0x361102b6b9b2b918ULL, 0xc0de81def35692daULL, 0xbe04c7e8a6c3c760ULL, 0xd766f9c9d570bb7fULL, 0xc230974d83561841ULL, 0x5bba1668c3be69a3ULL, 0x7f3811c410526294ULL, 0x29baee7ddadda71cULL, 0xbf813b8d145da900ULL, 0x641bbadf423f9a2cULL, 0xa80bb4ea44eacc5eULL, 0xcd65664814fde37cULL, 0x4a0364b9580291d2ULL, 0xee93dfb10826f0ddULL, 0x6b42db8dc5514724ULL, 0xbb10cf430b0f3785ULL, 0x40494e406f804216ULL, 0x55cfe107acf3aafbULL, 0x2088ec80e0ebae87ULL, 0x846a3ed011a337a0ULL, 0x48a45a4a1e3a5195ULL, 0xe5664568dfc50e16ULL, 0xab6a41294c0cc4ebULL, 0x82d0d602d268c7daULL, 0x6668449aed3cc48aULL, 0x5062cd0fb2015dfcULL, 0x7f2940a8b1ddb3d1ULL, 0x77f5b63a2a226448ULL, 0xfef0781361e443aeULL, 0xf977870e88d5c6c8ULL, 0x790364a61f676baaULL, 0x5887e72eceaddea3ULL, 0x1377e563a09a1b70ULL, 0x0c54efee1bd8c3b2ULL, 0x3ec3d15ad524d8f7ULL, 0xdaf15466b2383a5dULL, 0xe1e30a73bb94fec0ULL, 0x6a1c71015f3f7be2ULL, 0x842d43bf6369b1ffULL, 0x20fddadf107d20bcULL, 0x0000002f4b6dc970ULL};
static const uint64_t EXP_QM1_3[4] = {0x69602eb24829a9c2ULL, 0xdd2b2385cd7b4384ULL, 0xe81ac1e7808072c9ULL, 0x10216f7ba065e00dULL};
static const uint64_t EXP_QM1_2[4] = {0x9e10460b6c3e7ea3ULL, 0xcbc0b548b438e546ULL, 0xdc2822db40c0ac2eULL, 0x183227397098d014ULL};
static const uint64_t EXP_Q2M1_3[8] = {0x691c1d8b62747890ULL, 0x8cab57b9adf8eb00ULL, 0x18c55d8979dcee49ULL, 0x56cd8a31d35b6b98ULL, 0xb7a4a8c966ece684ULL, 0xe5592c705cbd1cacULL, 0x1dde2529566d9b5eULL, 0x030c96e827699534ULL};
static const uint64_t EXP_Q2M1_2[8] = {0x9daa2c5113aeb4d8ULL, 0x5301039684f56080ULL, 0x25280c4e36cb656eULL, 0x82344f4abd092164ULL, 0x1376fd2e1a6359c6ULL, 0x5805c2a88b1bab03ULL, 0x2ccd37be01a4690eULL, 0x0492e25c3b1e5fceULL};
static const uint64_t ATE_LOOP[2] = {0x9d797039be763ba8ULL, 0x1ULL};   /* 6u+2, u = 4965661367192848881 */

/* l_{T,S}(P) for the line through twist points with slope lam, evaluated at P = (xp, yp) in G1 */
static fe12 line_eval(const fe2 *lam, const g2_affine *T, const fe *xp, const fe *yp) {
    fe12 l; l.c0 = fe6_zero(); l.c1 = fe6_zero();
    l.c0.a0.c0 = *yp;                                                            /* y_P            * 1   */
    fe2 t = fe2_mul_fe(lam, xp); l.c1.a0 = fe2_neg(&t);                          /* -lam x_P       * w   */
    t = fe2_mul(lam, &T->x); l.c1.a1 = fe2_sub(&t, &T->y);                       /* lam x_T - y_T  * w^3 */
    return l;
}
/* T = T + S (affine, twist), returns the line value; handles doubling when S == T */
static fe12 line_step(g2_affine *T, const g2_affine *S, const fe *xp, const fe *yp) {
    fe2 lam;
    if (fe2_eq(&T->x, &S->x)) {
        if (!fe2_eq(&T->y, &S->y) || fe2_is_zero(&T->y)) {                        /* vertical line: value in Fq6, killed by the final exponentiation */
            T->inf = 1; return fe12_one();
        }
        fe2 x2 = fe2_sqr(&T->x), n = fe2_dbl(&x2); n = fe2_add(&n, &x2);
        fe2 dnm = fe2_dbl(&T->y); dnm = fe2_inv(&dnm); lam = fe2_mul(&n, &dnm);
    } else {
        fe2 n = fe2_sub(&S->y, &T->y), dnm = fe2_sub(&S->x, &T->x); dnm = fe2_inv(&dnm); lam = fe2_mul(&n, &dnm);
    }
    fe12 l = line_eval(&lam, T, xp, yp);
    fe2 x3 = fe2_sqr(&lam); x3 = fe2_sub(&x3, &T->x); x3 = fe2_sub(&x3, &S->x);
    fe2 y3 = fe2_sub(&T->x, &x3); y3 = fe2_mul(&y3, &lam); y3 = fe2_sub(&y3, &T->y);
    T->x = x3; T->y = y3;
    return l;
}
static fe12 miller_loop(const g1_affine *P, const g2_affine *Q) {
    fe12 f = fe12_one();
    if (P->inf || Q->inf) return f;
    g2_affine T = *Q;
    for (int i = 63; i >= 0; i--) {              /* bit 64 is the leading one */
        f = fe12_mul(&f, &f);
        fe12 l = line_step(&T, &T, &P->x, &P->y); f = fe12_mul(&f, &l);
        if ((ATE_LOOP[0] >> i) & 1) { l = line_step(&T, Q, &P->x, &P->y); f = fe12_mul(&f, &l); }
    }
    fe2 xi; xi.c0 = fe_from_u64(&FQ, 9); xi.c1 = fe_from_u64(&FQ, 1);
    fe2 g12 = fe2_pow(&xi, EXP_QM1_3, 4), g13 = fe2_pow(&xi, EXP_QM1_2, 4);
    fe2 g22 = fe2_pow(&xi, EXP_Q2M1_3, 8), g23 = fe2_pow(&xi, EXP_Q2M1_2, 8);
    g2_affine Q1, Q2; fe2 t;
    t = fe2_conj(&Q->x); Q1.x = fe2_mul(&t, &g12); t = fe2_conj(&Q->y); Q1.y = fe2_mul(&t, &g13); Q1.inf = 0;
    Q2.x = fe2_mul(&Q->x, &g22); t = fe2_mul(&Q->y, &g23); Q2.y = fe2_neg(&t); Q2.inf = 0;      /* -pi^2(Q) */
    fe12 l = line_step(&T, &Q1, &P->x, &P->y); f = fe12_mul(&f, &l);
    l = line_step(&T, &Q2, &P->x, &P->y); f = fe12_mul(&f, &l);
    return f;
}
static fe12 final_exponentiation(const fe12 *f) { return fe12_pow(f, FINAL_EXP, 44); }

int ora_pairing(const uint8_t *g1, const uint8_t *g2, uint8_t *out) {
    g1_affine P = g1_load(g1); g2_affine Q = g2_load(g2);
    fe12 f = miller_loop(&P, &Q); f = final_exponentiation(&f);
    const fe *c = (const fe *)&f;
    for (int i = 0; i < 12; i++) fq_store(out + 32 * i, &c[i]);
    return 0;
}

/* bellman groth16/verifier.rs `verify_proof` (called at prover.rs:200, helper.rs:158), SURVEY A.6 */
int ora_verify_proof(const void *pp, const uint8_t *proof, const uint8_t *public_inputs, size_t n_public) {
    const params_t *p = pp;
    if (n_public + 1 != p->n_ic) return ORA_ERR_MALFORMED_VK;
    g1_affine A = g1_load(proof), C = g1_load(proof + 192); g2_affine B = g2_load(proof + 64);
    g1_jac acc = g1_from_affine(&p->ic[0]);
    for (size_t i = 0; i < n_public; i++) {
        uint64_t k[4]; rd_le(public_inputs + 32 * i, k);
        g1_jac b = g1_from_affine(&p->ic[i + 1]); g1_jac t = g1_mul(&b, k); g1_add(&acc, &t);
    }
    g1_affine acc_a = g1_into_affine(&acc);
    g2_affine ng = p->gamma_g2, nd = p->delta_g2;
    if (!ng.inf) ng.y = fe2_neg(&ng.y);
    if (!nd.inf) nd.y = fe2_neg(&nd.y);
    fe12 f = miller_loop(&A, &B), t = miller_loop(&acc_a, &ng); f = fe12_mul(&f, &t);
    t = miller_loop(&C, &nd); f = fe12_mul(&f, &t);
    f = final_exponentiation(&f);
    fe12 rhs = miller_loop(&p->alpha_g1, &p->beta_g2); rhs = final_exponentiation(&rhs);
    return fe12_eq(&f, &rhs);
}
