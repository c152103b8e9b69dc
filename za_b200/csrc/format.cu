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

// ---------------------------------------------------------------------------------------------------------
// proving.key container — write_pk / read_pk of /root/reference/prover/src/groth16/format.rs:223-293:
//   u32BE len || bincode(Vec<BodyElementP>)            (the AST: opaque here, carried through untouched)
//   u32BE nC  || nC x (u32BE len || bincode(QEQ))      QEQ{a,b,c: LC}, LC(Vec<(usize, FS)>), FS(BigUint)
//   u32BE nI  || nI x u32BE signal id                   (signals removed by the optimiser, ascending)
//   bellman Parameters::write                           (za_pk_load takes it from params_offset)
// bincode 1.2 defaults: little-endian, u64 lengths, usize as u64; num-bigint's BigUint serialises as a
// sequence of u32 digits, least significant first.
// Constraints cross the ABI as three CSR matrices over SIGNAL ids with 32-byte LE coefficients, za's sign
// convention a*b + c = 0 (qeq.rs:9-13).
// ---------------------------------------------------------------------------------------------------------
namespace {
struct Rd {
    const uint8_t* p; const uint8_t* end;
    void need(size_t n) { if ((size_t)(end - p) < n) throw ZaError(ZA_ERR_IO, "proving.key: unexpected end of file"); }
    uint32_t be32() { need(4); uint32_t v = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; p += 4; return v; }
    uint64_t le64() { need(8); uint64_t v = 0; for (int i = 7; i >= 0; i--) v = (v << 8) | p[i]; p += 8; return v; }
    uint32_t le32() { need(4); uint32_t v = 0; for (int i = 3; i >= 0; i--) v = (v << 8) | p[i]; p += 4; return v; }
};
// one LC: calls f(signal, coeff_le32bytes) per term, returns the term count
template <class Fn>
uint64_t read_lc(Rd& r, Fn f) {
    uint64_t n = r.le64();
    if (n > (uint64_t)(r.end - r.p)) throw ZaError(ZA_ERR_IO, "proving.key: implausible term count");
    for (uint64_t t = 0; t < n; t++) {
        uint64_t sig = r.le64();
        uint64_t nd = r.le64();
        if (nd > 8) throw ZaError(ZA_ERR_BAD_ENCODING, "proving.key: coefficient wider than 256 bits");
        uint8_t c[32] = {0};
        for (uint64_t d = 0; d < nd; d++) { uint32_t w = r.le32(); memcpy(c + 4 * d, &w, 4); }
        if (sig > 0xffffffffull) throw ZaError(ZA_ERR_BAD_ENCODING, "proving.key: signal id does not fit 32 bits");
        f((uint32_t)sig, c);
    }
    return n;
}
void put_be32(std::string& o, uint32_t v) { o.push_back((char)(v >> 24)); o.push_back((char)(v >> 16)); o.push_back((char)(v >> 8)); o.push_back((char)v); }
void put_le64(std::string& o, uint64_t v) { for (int i = 0; i < 8; i++) o.push_back((char)(v >> (8 * i))); }
}  // namespace

extern "C" {

// info[6] = num_constraints, num_ignore, nnz_a, nnz_b, nnz_c, max signal id + 1; *params_offset = start of Parameters
int za_pkfile_scan(const uint8_t* file, size_t len, uint64_t* info, size_t* params_offset, size_t* ast_offset, size_t* ast_len) {
    if (!file || !info || !params_offset) return fail(ZA_ERR_INVALID, "NULL argument");
    try {
        Rd r{file, file + len};
        uint32_t al = r.be32(); r.need(al);
        if (ast_offset) *ast_offset = 4;
        if (ast_len) *ast_len = al;
        r.p += al;
        uint32_t nc = r.be32();
        uint64_t nnz[3] = {0, 0, 0}, maxsig = 0;
        for (uint32_t k = 0; k < nc; k++) {
            uint32_t ql = r.be32(); r.need(ql);
            Rd q{r.p, r.p + ql};
            for (int w = 0; w < 3; w++) nnz[w] += read_lc(q, [&](uint32_t s, const uint8_t*) { if ((uint64_t)s + 1 > maxsig) maxsig = (uint64_t)s + 1; });
            if (q.p != q.end) throw ZaError(ZA_ERR_BAD_ENCODING, "proving.key: trailing bytes in a QEQ record");
            r.p += ql;
        }
        uint32_t ni = r.be32(); r.need(4ull * ni); r.p += 4ull * ni;
        info[0] = nc; info[1] = ni; info[2] = nnz[0]; info[3] = nnz[1]; info[4] = nnz[2]; info[5] = maxsig;
        *params_offset = (size_t)(r.p - file);
        return ZA_OK;
    } catch (const ZaError& e) { return fail(e.code, "%s", e.what()); }
}

// second pass: fill caller buffers sized from za_pkfile_scan.  ptr[w]: nc+1, sig[w]: nnz_w, coeff[w]: nnz_w*32, ignore: num_ignore
int za_pkfile_read(const uint8_t* file, size_t len, uint32_t* const* ptr, uint32_t* const* sig, uint8_t* const* coeff, uint32_t* ignore) {
    if (!file || !ptr || !sig || !coeff) return fail(ZA_ERR_INVALID, "NULL argument");
    try {
        Rd r{file, file + len};
        uint32_t al = r.be32(); r.need(al); r.p += al;
        uint32_t nc = r.be32();
        uint64_t pos[3] = {0, 0, 0};
        for (int w = 0; w < 3; w++) ptr[w][0] = 0;
        for (uint32_t k = 0; k < nc; k++) {
            uint32_t ql = r.be32(); r.need(ql);
            Rd q{r.p, r.p + ql};
            for (int w = 0; w < 3; w++) {
                read_lc(q, [&](uint32_t s, const uint8_t* c) { sig[w][pos[w]] = s; memcpy(coeff[w] + 32 * pos[w], c, 32); pos[w]++; });
                ptr[w][k + 1] = (uint32_t)pos[w];
            }
            r.p += ql;
        }
        uint32_t ni = r.be32();
        for (uint32_t i = 0; i < ni; i++) { uint32_t v = r.be32(); if (ignore) ignore[i] = v; }
        return ZA_OK;
    } catch (const ZaError& e) { return fail(e.code, "%s", e.what()); }
}

// write_pk (format.rs:223-252).  ast: the opaque bincode(Vec<BodyElementP>) blob (8 zero bytes = empty vector).
// Returns ZA_ERR_BUFFER_TOO_SMALL with *out_len set when size is insufficient.
int za_pkfile_write(const uint8_t* ast, size_t ast_len, uint32_t nc, const uint32_t* const* ptr, const uint32_t* const* sig,
                    const uint8_t* const* coeff, const uint32_t* ignore, uint32_t num_ignore, const uint8_t* params, size_t params_len,
                    uint8_t* out, size_t size, size_t* out_len) {
    if (!ptr || !sig || !coeff || !out_len || (ast_len && !ast) || (params_len && !params)) return fail(ZA_ERR_INVALID, "NULL argument");
    std::string o;
    put_be32(o, (uint32_t)ast_len);
    o.append((const char*)ast, ast_len);
    put_be32(o, nc);
    for (uint32_t k = 0; k < nc; k++) {
        std::string q;
        for (int w = 0; w < 3; w++) {
            put_le64(q, ptr[w][k + 1] - ptr[w][k]);
            for (uint32_t t = ptr[w][k]; t < ptr[w][k + 1]; t++) {
                put_le64(q, sig[w][t]);
                const uint8_t* c = coeff[w] + 32 * (size_t)t;
                int nd = 8;
                while (nd > 0 && c[4 * nd - 1] == 0 && c[4 * nd - 2] == 0 && c[4 * nd - 3] == 0 && c[4 * nd - 4] == 0) nd--;   // BigUint has no leading zero digits
                put_le64(q, (uint64_t)nd);
                q.append((const char*)c, 4 * nd);
            }
        }
        put_be32(o, (uint32_t)q.size());
        o += q;
    }
    put_be32(o, num_ignore);
    for (uint32_t i = 0; i < num_ignore; i++) put_be32(o, ignore[i]);
    o.append((const char*)params, params_len);
    *out_len = o.size();
    if (!out || size < o.size()) return fail(ZA_ERR_BUFFER_TOO_SMALL, "proving.key needs %zu bytes", o.size());
    memcpy(out, o.data(), o.size());
    return ZA_OK;
}

// CircomCircuit::synthesize (prover.rs:45-103) as data: signals 1..n-1 in id order become bellman variables —
// input (alloc_input) if is_public[id] (a depth-1 output or public input, signal.rs:58-62), aux (alloc) otherwise,
// none if the id is in the (ascending) ignore list; signal 0 is input 0 = one.  Constraint k contributes
// enforce(A = a, B = b, C = -c) (prover.rs:96-98).
// var_of_signal[n_signals] receives the variable (ZA_VAR_AUX | i for aux, 0xffffffff for ignored);
// out_var[w] (nnz_w) the variable per term, out_c_coeff (nnz_c * 32) the negated C coefficients.
int za_synthesize(uint32_t n_signals, const uint8_t* is_public, const uint32_t* ignore, uint32_t num_ignore, uint32_t nc,
                  const uint32_t* const* ptr, const uint32_t* const* sig, const uint8_t* c_coeff, uint32_t* var_of_signal,
                  uint32_t* const* out_var, uint8_t* out_c_coeff, uint32_t* num_inputs, uint32_t* num_aux) {
    if (!is_public || !ptr || !sig || !var_of_signal || !out_var || !num_inputs || !num_aux || (num_ignore && !ignore)) return fail(ZA_ERR_INVALID, "NULL argument");
    if (n_signals == 0) return fail(ZA_ERR_INVALID, "signal 0 (one) is missing");
    uint32_t ni = 1, na = 0, ig = 0;
    var_of_signal[0] = 0;
    for (uint32_t s = 1; s < n_signals; s++) {
        if (ig < num_ignore && ignore[ig] == s) { var_of_signal[s] = 0xffffffffu; ig++; continue; }
        if (is_public[s]) var_of_signal[s] = ni++;
        else var_of_signal[s] = ZA_VAR_AUX | na++;
    }
    for (int w = 0; w < 3; w++)
        for (uint32_t t = 0; t < ptr[w][nc]; t++) {
            uint32_t s = sig[w][t];
            if (s >= n_signals || var_of_signal[s] == 0xffffffffu) return fail(ZA_ERR_INVALID, "signal %u not defined", s);   // format.rs:215-217
            out_var[w][t] = var_of_signal[s];
        }
    if (out_c_coeff) {
        for (uint32_t t = 0; t < ptr[2][nc]; t++) {
            Fr c; memcpy(c.v, c_coeff + 32 * (size_t)t, 32);
            if (!fp_is_canonical<FrParams>(c.v)) return fail(ZA_ERR_NOT_CANONICAL, "C coefficient %u is not reduced", t);
            Fr n = fp_neg<FrParams>(c);              // negation is the same in canonical and Montgomery form
            memcpy(out_c_coeff + 32 * (size_t)t, n.v, 32);
        }
    }
    *num_inputs = ni; *num_aux = na;
    return ZA_OK;
}

}  // extern "C"
