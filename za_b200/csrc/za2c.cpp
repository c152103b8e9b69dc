// libza2c: the outer C ABI of za's bindings on top of libza_b200 (include/za2c.h).
// Reference: /root/reference/binding/c/native/src/lib.rs:10-117, binding/go/lib.go:6-9.
#include "../../include/za2c.h"
#include "../../include/za_b200.h"
#include <stdio.h>
#include <string.h>
#include <string>

static int g_verbose = 0;

// lib.rs:22-31 `return_string`: len >= size is "too small", otherwise copy and return `ret`
static int return_string(const std::string& s, char* buf, size_t size, int ret) {
    if (!buf || s.size() >= size) return ZA2C_ERR_BUFFER_TOO_SMALL;
    memcpy(buf, s.c_str(), s.size() + 1);
    return ret;
}

static const char* FRONT_END_TEXT =
    ": this build replaces the Groth16 hot path only; compiling .za source and evaluating the witness need za's front-end. "
    "Use the kernel-level ABI of include/za_b200.h with the constraint system and the signal values: ";

extern "C" {

void verbose(int on) { g_verbose = on != 0; }

int setup(const char* circuit_path, const char* pk_path, const char* verifier_type, char* verifier_buf, size_t verifier_buf_size, char* err_buf,
          size_t err_buf_size) {
    (void)verifier_buf; (void)verifier_buf_size;
    if (!circuit_path || !pk_path || !verifier_type) return return_string("NULL argument", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    if (strcmp(verifier_type, "json") != 0 && strcmp(verifier_type, "solidity") != 0)
        return return_string("invalid validator type", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);      // lib.rs:68
    if (g_verbose) fprintf(stderr, "[za2c] setup(%s): front-end not available\n", circuit_path);
    return return_string(std::string("setup") + FRONT_END_TEXT + "za_circuit_upload + za_generate_parameters + za_pkfile_write + za_vk_to_json",
                         err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
}

int prove(const char* pk_path, const char* inputs_json, char* proof_buf, size_t proof_buf_size, char* err_buf, size_t err_buf_size) {
    (void)proof_buf; (void)proof_buf_size;
    if (!pk_path || !inputs_json) return return_string("NULL argument", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    if (g_verbose) fprintf(stderr, "[za2c] prove(%s): front-end not available\n", pk_path);
    return return_string(std::string("prove") + FRONT_END_TEXT + "za_pkfile_read + za_synthesize + za_create_proof + za_proof_to_json",
                         err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
}

int verify(const char* vk_json, const char* proof_with_inputs_json, char* err_buf, size_t err_buf_size) {
    if (!vk_json || !proof_with_inputs_json) return return_string("NULL argument", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    int valid = 0;
    const int rc = za_verify_json(vk_json, proof_with_inputs_json, &valid);
    if (rc != ZA_OK) return return_string(za_last_error(), err_buf, err_buf_size, ZA2C_ERR_CUSTOM);   // lib.rs:115
    if (g_verbose) fprintf(stderr, "[za2c] verify: %s\n", valid ? "valid" : "not valid");
    return valid ? ZA2C_ERR_NONE : ZA2C_ERR_VERIFICATION_FAILED;                                      // lib.rs:113-114
}

}  // extern "C"
