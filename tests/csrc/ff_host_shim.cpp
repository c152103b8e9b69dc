// Host build of za_b200/csrc/ff.cuh + ec.cuh for CPU-side unit tests (no GPU needed).
// Test infrastructure only: exercises the exact limb schedule the device runs,
// on the emulated carry flag.
#define ZA_FF_EMULATE_PTX 1
#include "../../za_b200/csrc/ff.cuh"
#include <string.h>
using namespace za;
extern "C" {
uint64_t shim_lost_carries() { return host_cc().lost; }
#define BINOP(name, T, expr) void name(const uint32_t* a, const uint32_t* b, uint32_t* r) { T x, y; memcpy(&x, a, sizeof(T)); memcpy(&y, b, sizeof(T)); T z = expr; memcpy(r, &z, sizeof(T)); }
#define UNOP(name, T, expr) void name(const uint32_t* a, uint32_t* r) { T x; memcpy(&x, a, sizeof(T)); T z = expr; memcpy(r, &z, sizeof(T)); }
BINOP(shim_fr_mul, Fr, x * y)
BINOP(shim_fr_add, Fr, x + y)
BINOP(shim_fr_sub, Fr, x - y)
UNOP(shim_fr_neg, Fr, -x)
UNOP(shim_fr_inv, Fr, inv(x))
UNOP(shim_fr_to_mont, Fr, fp_to_mont(x))
UNOP(shim_fr_from_mont, Fr, fp_from_mont(x))
BINOP(shim_fq_mul, Fq, x * y)
BINOP(shim_fq_add, Fq, x + y)
BINOP(shim_fq_sub, Fq, x - y)
UNOP(shim_fq_neg, Fq, -x)
UNOP(shim_fq_inv, Fq, inv(x))
UNOP(shim_fq_to_mont, Fq, fp_to_mont(x))
UNOP(shim_fq_from_mont, Fq, fp_from_mont(x))
BINOP(shim_fq2_mul, Fq2, x * y)
UNOP(shim_fq2_sqr, Fq2, sqr(x))
UNOP(shim_fq2_inv, Fq2, inv(x))
UNOP(shim_fr_inv_kaliski, Fr, fp_inv_kaliski(x))
UNOP(shim_fq_inv_kaliski, Fq, fp_inv_kaliski(x))
}
#include "../../za_b200/csrc/ec.cuh"
extern "C" {
static G1Affine ld_aff(const uint32_t* p) { G1Affine a; memcpy(&a, p, sizeof a); return a; }
static void st_aff(uint32_t* p, const G1Affine& a) { memcpy(p, &a, sizeof a); }
void shim_g1_add_mixed(const uint32_t* a, const uint32_t* b, uint32_t* r) {
    G1XYZZ acc = G1XYZZ::from_affine(ld_aff(a)); xyzz_madd<Fq>(acc, ld_aff(b)); st_aff(r, xyzz_to_affine<Fq>(acc));
}
void shim_g1_add(const uint32_t* a, const uint32_t* b, uint32_t* r) {
    G1XYZZ acc = G1XYZZ::from_affine(ld_aff(a)); xyzz_add<Fq>(acc, G1XYZZ::from_affine(ld_aff(b))); st_aff(r, xyzz_to_affine<Fq>(acc));
}
void shim_g1_mul(const uint32_t* a, const uint32_t* k, uint32_t* r) {
    st_aff(r, xyzz_to_affine<Fq>(xyzz_mul<Fq>(G1XYZZ::from_affine(ld_aff(a)), k)));
}
}
// round 2: dedicated squaring, 512-bit products, stand-alone reduction, lazily reduced Fq2 product
extern "C" {
UNOP(shim_fr_sqr, Fr, fp_sqr(x))
UNOP(shim_fq_sqr, Fq, fp_sqr(x))
BINOP(shim_fq2_mul_lazy, Fq2, fq2_mul_lazy(x, y))
BINOP(shim_fq2_mul_karatsuba, Fq2, fq2_mul_karatsuba(x, y))
void shim_mul_wide(const uint32_t* a, const uint32_t* b, uint32_t* t) { uint32_t w[16]; memset(w, 0xa5, sizeof w); u256_mul_wide(w, a, b); memcpy(t, w, sizeof w); }
void shim_sqr_wide(const uint32_t* a, uint32_t* t) { uint32_t w[16]; memset(w, 0xa5, sizeof w); u256_sqr_wide(w, a); memcpy(t, w, sizeof w); }
void shim_fq_redc_wide(const uint32_t* t, uint32_t* r) { Fq z = fp_redc_wide<FqParams>(t); memcpy(r, &z, sizeof z); }
void shim_fr_redc_wide(const uint32_t* t, uint32_t* r) { Fr z = fp_redc_wide<FrParams>(t); memcpy(r, &z, sizeof z); }
}
