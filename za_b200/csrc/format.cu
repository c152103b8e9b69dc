// proof.json writer — JsonProofAndInput of /root/reference/prover/src/groth16/format.rs:80-128.
// serde_json compact output, field order a, b, c, public_inputs; a, c = [x, y];
// b = [[x.c0, x.c1], [y.c0, y.c1]] (format.rs:74-77); coordinates are FqRepr::to_string() =
// "0x" + 64 lowercase hex digits (format.rs:51); public inputs are decimal strings (format.rs:94-97).
#include "common.cuh"
#include <string.h>
#include <string>

using namespace za;

static std::string hex_coord(const uint8_t* le) {
    static const char* d = "0123456789abcdef";
    std::string s = "\"0x";
    for (int i = 31; i >= 0; i--) { s.push_back(d[le[i] >> 4]); s.push_back(d[le[i] & 15]); }
    s.push_back('"');
    return s;
}
// 256-bit little-endian -> decimal
static std::string dec_scalar(const uint8_t* le) {
    uint32_t w[8];
    memcpy(w, le, 32);
    std::string out;
    bool nz = true;
    while (nz) {
        uint64_t rem = 0;
        nz = false;
        for (int i = 7; i >= 0; i--) {
            uint64_t cur = (rem << 32) | w[i];
            w[i] = (uint32_t)(cur / 1000000000u);
            rem = cur % 1000000000u;
            if (w[i]) nz = true;
        }
        char b[16];
        snprintf(b, sizeof b, nz ? "%09u" : "%u", (unsigned)rem);
        out = std::string(b) + out;
    }
    return out;
}

extern "C" int za_proof_to_json(const uint8_t* proof, const uint8_t* public_inputs, size_t n_public, char* buf, size_t size) {
    if (!proof || !buf || (n_public && !public_inputs)) return fail(ZA_ERR_INVALID, "NULL argument");
    std::string s = "{\"a\":[" + hex_coord(proof) + "," + hex_coord(proof + 32) + "],";
    s += "\"b\":[[" + hex_coord(proof + 64) + "," + hex_coord(proof + 96) + "],[" + hex_coord(proof + 128) + "," + hex_coord(proof + 160) + "]],";
    s += "\"c\":[" + hex_coord(proof + 192) + "," + hex_coord(proof + 224) + "],\"public_inputs\":[";
    for (size_t i = 0; i < n_public; i++) {
        if (i) s += ",";
        s += "\"" + dec_scalar(public_inputs + 32 * i) + "\"";
    }
    s += "]}";
    if (s.size() >= size) return fail(ZA_ERR_BUFFER_TOO_SMALL, "proof json needs %zu bytes", s.size() + 1);   // binding/c lib.rs:23
    memcpy(buf, s.c_str(), s.size() + 1);
    return ZA_OK;
}
