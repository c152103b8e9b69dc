/* ORACLE — TEST INFRASTRUCTURE ONLY (see field.h).  PARITY UNPINNED.
 *
 * C interface of the CPU restatement of the Groth16 proving path that
 * /root/reference/prover/src/groth16/prover.rs:173 enters (bellman_ce
 * create_proof / multiexp / EvaluationDomain, pairing_ce bn256).
 *
 * Interchange encodings (shared with include/za_b200.h):
 *   Fr scalar : 32 bytes, little-endian canonical integer (< r)
 *   G1 affine : 64 bytes, x || y, each 32-byte LE canonical; infinity = 64 zero bytes
 *   G2 affine : 128 bytes, x.c0 || x.c1 || y.c0 || y.c1; infinity = 128 zero bytes
 *   proof     : a (G1, 64) || b (G2, 128) || c (G1, 64)
 * The on-disk big-endian `Parameters` layout (format.rs:250,285) is handled by
 * ora_params_write / ora_params_read.
 */
#ifndef ZA_ORACLE_H
#define ZA_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORA_VAR_AUX 0x80000000u /* term variable: bit 31 set = aux index, else input index */

/* What bellman's ConstraintSystem sees after CircomCircuit::synthesize
 * (prover.rs:45-103): rows of enforce(A, B, C) as three CSR matrices; terms
 * keep insertion order.  Input 0 is the constant one. */
typedef struct {
    uint32_t num_inputs;      /* including the leading `one` */
    uint32_t num_aux;
    uint32_t num_constraints; /* circuit rows only; the input-consistency rows are appended internally */
    const uint32_t *ptr[3];   /* [num_constraints+1] term offsets for A, B, C */
    const uint32_t *var[3];   /* variable per term */
    const uint8_t *coeff[3];  /* 32-byte LE canonical coefficient per term */
} ora_r1cs;

/* optional intermediates of ora_create_proof, any pointer may be NULL */
typedef struct {
    uint8_t *a_eval, *b_eval, *c_eval; /* (num_constraints+num_inputs)*32 : ProvingAssignment a,b,c */
    uint8_t *h_coeffs;                 /* (m-1)*32 */
    uint8_t *msm_g1;                   /* 7*64: h, l, a_inputs, a_aux, b1_inputs, b1_aux, (unused) */
    uint8_t *msm_g2;                   /* 2*128: b2_inputs, b2_aux */
    uint8_t *a_aux_density, *b_input_density, *b_aux_density; /* one byte per variable */
} ora_trace;

/* op: 0 add, 1 sub, 2 mul, 3 inv(a), 4 neg(a).  field: 0 Fr, 1 Fq */
void ora_field_op(int field, int op, const uint8_t *a, const uint8_t *b, uint8_t *out);
void ora_fq2_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *out); /* 64-byte operands */

int ora_g1_mul(const uint8_t *p, const uint8_t *k, uint8_t *out);
int ora_g1_add(const uint8_t *p, const uint8_t *q, uint8_t *out);
int ora_g2_mul(const uint8_t *p, const uint8_t *k, uint8_t *out);
int ora_g2_add(const uint8_t *p, const uint8_t *q, uint8_t *out);
int ora_g1_on_curve(const uint8_t *p);
int ora_g2_on_curve(const uint8_t *p);
/* out[i] = (i+1)*base, by an affine addition chain */
int ora_g1_multiples(const uint8_t *base, size_t n, uint8_t *out);
int ora_g2_multiples(const uint8_t *base, size_t n, uint8_t *out);
void ora_g1_generator(uint8_t *out);
void ora_g2_generator(uint8_t *out);

/* EvaluationDomain.  mode: 0 fft, 1 ifft, 2 coset_fft, 3 icoset_fft.  In place, natural order. */
int ora_fft(uint8_t *data, int log_n, int mode, int threads);
void ora_domain_omega(int log_n, uint8_t *omega_out);
/* create_proof step 4: a,b,c evaluations (len each) -> h coefficients ((m-1)*32, m = next pow2 >= len).
 * checkpoints (optional, 8*m*32): a.ifft, a.coset_fft, b.ifft, b.coset_fft, c.ifft, c.coset_fft,
 * (a*b-c)/Z, icoset_fft result. */
int ora_h_poly(const uint8_t *a, const uint8_t *b, const uint8_t *c, size_t len, uint8_t *h_out,
               uint8_t *checkpoints, int threads);

/* bellman multiexp: density == NULL means FullDensity (n_exp bases consumed). */
int ora_multiexp_g1(const uint8_t *bases, size_t n_bases, const uint8_t *scalars, size_t n_exp,
                    const uint8_t *density, uint8_t *out, int threads);
int ora_multiexp_g2(const uint8_t *bases, size_t n_bases, const uint8_t *scalars, size_t n_exp,
                    const uint8_t *density, uint8_t *out, int threads);

/* Groth16 */
void *ora_generate_parameters(const ora_r1cs *cs, const uint8_t *alpha, const uint8_t *beta,
                              const uint8_t *gamma, const uint8_t *delta, const uint8_t *tau,
                              const uint8_t *g1, const uint8_t *g2, int threads, int *err);
size_t ora_params_size(const void *params);
int ora_params_write(const void *params, uint8_t *buf);
void *ora_params_read(const uint8_t *buf, size_t len, int checked, int *err);
void ora_params_free(void *params);
/* counts[6] = ic,h,l,a,b_g1,b_g2: bases are known multiples of the generators (not a valid CRS) */
void *ora_params_synthetic(const uint32_t *counts, int threads);
void ora_params_counts(const void *params, uint32_t *counts /* ic,h,l,a,b_g1,b_g2 */);
int ora_create_proof(const void *params, const ora_r1cs *cs, const uint8_t *inputs, const uint8_t *aux,
                     const uint8_t *r, const uint8_t *s, uint8_t *proof, ora_trace *trace, int threads);
/* returns 1 valid, 0 invalid, <0 malformed */
int ora_verify_proof(const void *params, const uint8_t *proof, const uint8_t *public_inputs, size_t n_public);
/* e(P,Q) as 12 Fq coefficients (384 bytes), for bilinearity tests */
int ora_pairing(const uint8_t *g1, const uint8_t *g2, uint8_t *out);

enum {
    ORA_OK = 0,
    ORA_ERR_UNEXPECTED_IDENTITY = -1,
    ORA_ERR_POLY_DEGREE_TOO_LARGE = -2,
    ORA_ERR_UNCONSTRAINED_VARIABLE = -3,
    ORA_ERR_IO = -4,
    ORA_ERR_NOT_ON_CURVE = -5,
    ORA_ERR_NOT_IN_SUBGROUP = -6,
    ORA_ERR_MALFORMED_VK = -7,
    ORA_ERR_BAD_ENCODING = -8,
    ORA_ERR_ASSIGNMENT_MISSING = -9
};
#ifdef __cplusplus
}
#endif
#endif
