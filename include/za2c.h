/* za2c.h — the OUTER C ABI of za's bindings (libza2c), as the reference pins it twice:
 *   Rust  /root/reference/binding/c/native/src/lib.rs:10-13 (return codes), :33-117 (symbols)
 *   C     /root/reference/binding/go/lib.go:6-9 (prototypes the Go binding links against)
 * za_b200/libza2c.so exports these four symbols on top of libza_b200.so so that a binding built against the reference's
 * libza2c links unchanged.  `verify` is complete (helper::verify, helper.rs:149-158: JsonVerifyingKey + JsonProofAndInput
 * -> pairing check).  `setup` and `prove` need za's front-end (parser, evaluator, optimiser: SURVEY.md §8f N4, out of
 * scope of the hot path): they fail loudly with ERR_CUSTOM and a message naming the kernel-level entry points
 * (za_generate_parameters / za_pkfile_read + za_synthesize + za_create_proof in include/za_b200.h) that take the
 * constraint system and the signal values instead of source text.
 *
 * Conventions (lib.rs:22-31): the caller allocates every output buffer; strings are NUL-terminated UTF-8; a string of
 * length >= the buffer size is "too small" (return 1) and nothing is written. */
#ifndef ZA2C_H
#define ZA2C_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ZA2C_ERR_NONE 0                 /* lib.rs:10 */
#define ZA2C_ERR_BUFFER_TOO_SMALL 1     /* lib.rs:11 */
#define ZA2C_ERR_VERIFICATION_FAILED 2  /* lib.rs:12 */
#define ZA2C_ERR_CUSTOM 100             /* lib.rs:13: error text in err_buf */

/* lib.rs:33-49.  The reference initialises its logger here (and panics on a second call); this library has no logger:
 * the flag only switches the one-line notices of libza2c itself on stderr. */
void verbose(int on);
/* lib.rs:51-81 */
int setup(const char *circuit_path, const char *pk_path, const char *verifier_type, char *verifier_buf,
          size_t verifier_buf_size, char *err_buf, size_t err_buf_size);
/* lib.rs:83-101 */
int prove(const char *pk_path, const char *inputs_json, char *proof_buf, size_t proof_buf_size, char *err_buf,
          size_t err_buf_size);
/* lib.rs:103-117: 0 = valid, 2 = the proof does not verify, 100 = malformed JSON / bad coordinates (text in err_buf) */
int verify(const char *vk_json, const char *proof_with_inputs_json, char *err_buf, size_t err_buf_size);

#ifdef __cplusplus
}
#endif
#endif
