/* za2c.h — the OUTER C ABI of za's bindings (libza2c), as the reference pins it twice:
 *   Rust  /root/reference/binding/c/native/src/lib.rs:10-13 (return codes), :33-117 (symbols)
 *   C     /root/reference/binding/go/lib.go:6-9 (prototypes the Go binding links against)
 * za_b200/libza2c.so exports these four symbols on top of libza_b200.so so that a binding built against the reference's
 * libza2c links unchanged:
 *   setup  = helper::setup  (helper.rs:22-89):  compile the circuit (parser + evaluator in constraint mode + optimiser,
 *            host C++ under za_b200/csrc/frontend/), generate_random_parameters on the GPU, write_pk, verifier text
 *   prove  = helper::prove  (helper.rs:91-147): read_pk, witness from the stored syntax tree and the inputs, constraint
 *            check + create_random_proof on the GPU, self-verification, proof.json
 *   verify = helper::verify (helper.rs:149-158): JsonVerifyingKey + JsonProofAndInput -> pairing check (host)
 * There is no CPU fallback for the GPU stages: without a CUDA device setup and prove return ERR_CUSTOM with the CUDA
 * error text.
 *
 * Conventions (lib.rs:22-31): the caller allocates every output buffer; strings are NUL-terminated UTF-8; a string of
 * length >= the buffer size is "too small" (return 1) and nothing is written.  Errors are the Debug text of the
 * reference's error enums (e.g. Evaluator(NotFound("template T"))). */
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

/* lib.rs:33-49.  The reference initialises its logger here (and panics on a second call) and sets BELLMAN_VERBOSE; here
 * the flag switches the info! lines of helper.rs / prover.rs (stage names and times) on stderr and the per-multiexp
 * timeline of the GPU backend. */
void verbose(int on);
/* lib.rs:51-81 */
int setup(const char *circuit_path, const char *pk_path, const char *verifier_type, char *verifier_buf,
          size_t verifier_buf_size, char *err_buf, size_t err_buf_size);
/* lib.rs:83-101 */
int prove(const char *pk_path, const char *inputs_json, char *proof_buf, size_t proof_buf_size, char *err_buf,
          size_t err_buf_size);
/* lib.rs:103-117: 0 = valid, 2 = the proof does not verify, 100 = malformed JSON / bad coordinates (text in err_buf) */
int verify(const char *vk_json, const char *proof_with_inputs_json, char *err_buf, size_t err_buf_size);


/* ---- extensions (not in the reference's ABI) ------------------------------------------------------------------
 * za2c_release: drops the proving key that prove() keeps loaded between calls.
 * The front-end seam below needs no GPU; tests and tools use it to look at what the parser and the evaluator produce.
 *   za2c_parse: what = 0 expression, 1 statement, 2 body element -> the Debug text of parser/src/display.rs;
 *               3 body -> hex of its bincode image (after a serialise / deserialise round trip); 4 -> preprocessed text
 *   za2c_eval:  mode 1 = Mode::GenConstraints, 2 = Mode::GenWitness over `source` (eval_inline) or `file_path`
 *               (eval_file); deferred_json = {"main.a": "4", ...} input values by full signal name; check != 0 in
 *               witness mode also evaluates the constraints of the same text on the witness.  out = JSON with
 *               "signals" ("name:Type:value", signal.rs:162-165), "constraints" (QEQ text, qeq.rs:20-32), "scope" and,
 *               in constraint mode, "optimized" / "removed" (optimizer/mod.rs)
 *   za2c_test:  `za test` of one file (compiler/src/tester/embeeded.rs): every #[test] template, witness then
 *               constraints, compared and checked */
void za2c_release(void);
int za2c_parse(int what, const char *text, char *out, size_t out_size, char *err_buf, size_t err_buf_size);
int za2c_eval(int mode, const char *source, const char *file_path, const char *deferred_json, int check, char *out,
              size_t out_size, char *err_buf, size_t err_buf_size);
int za2c_test(const char *file_path, const char *prefix, char *out, size_t out_size, char *err_buf, size_t err_buf_size);

#ifdef __cplusplus
}
#endif
#endif
