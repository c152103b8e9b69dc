// libza2c: the outer C ABI of za's bindings (include/za2c.h) — verbose / setup / prove / verify — on top of libza_b200
// (the GPU backend) and the front-end under csrc/frontend/ (parser, evaluator, optimiser: host C++).
// Reference: /root/reference/binding/c/native/src/lib.rs:10-117 (symbols, return codes, buffer rule),
// /root/reference/prover/src/groth16/helper.rs:22-158 (what setup / prove / verify do),
// /root/reference/prover/src/groth16/prover.rs:105-208 (setup, generate_verified_proof),
// /root/reference/prover/src/groth16/format.rs:296-335 (flatten_json).
#include "../../include/za2c.h"
#include "../../include/za_b200.h"
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <chrono>
#include <fstream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include "frontend/frontend.hpp"

using namespace zafe;

static int g_verbose = 0;
static std::mutex g_mutex;             // Signals hold Rc in the reference (single threaded); one call at a time here too

// lib.rs:22-31 `return_string`: len >= size is "too small", otherwise copy and return `ret`
static int return_string(const std::string& s, char* buf, size_t size, int ret) {
    if (!buf || s.size() >= size) return ZA2C_ERR_BUFFER_TOO_SMALL;
    memcpy(buf, s.c_str(), s.size() + 1);
    return ret;
}

namespace {

struct Span {                                       // the info!("... time: {:?}") lines of helper.rs / prover.rs
    const char* text;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit Span(const char* t) : text(t) {}
    ~Span() {
        if (!g_verbose) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[za2c] %s: %.3fms\n", text, ms);
    }
};
void info(const std::string& s) { if (g_verbose) fprintf(stderr, "[za2c] %s\n", s.c_str()); }

[[noreturn]] void backend_fail(const char* what) { throw FeError("Unexpected", std::string(what) + ": " + za_last_error()); }
void zcheck(int rc, const char* what) { if (rc != ZA_OK) backend_fail(what); }

struct CtxHandle { za_ctx* p = nullptr; ~CtxHandle() { if (p) za_ctx_destroy(p); } };
struct PkHandle { za_pk* p = nullptr; ~PkHandle() { if (p) za_pk_free(p); } };
struct CircuitHandle { za_circuit* p = nullptr; ~CircuitHandle() { if (p) za_circuit_free(p); } };
struct BasesHandle { za_bases* p = nullptr; ~BasesHandle() { if (p) za_bases_free(p); } };

int device_from_env() { const char* e = getenv("ZA2C_DEVICE"); return e ? atoi(e) : 0; }

void fs_bytes(const FS& f, uint8_t* out) { memcpy(out, f.n.v, 32); }        // 32-byte LE canonical (little-endian host)

// a uniform non-zero scalar below r (thread_rng of prover.rs:112,146)
FS random_scalar() {
    std::ifstream ur("/dev/urandom", std::ios::binary);
    if (!ur) throw FeError("Unexpected", "cannot open /dev/urandom");
    while (true) {
        U256 x;
        ur.read((char*)x.v, 32);
        if (!ur) throw FeError("Unexpected", "short read from /dev/urandom");
        x.v[3] &= 0x3fffffffffffffffull;            // 254 bits, then rejection
        if (cmp(x, field_r()) < 0 && !x.is_zero()) return FS(x);
    }
}

// Constraints as three CSR matrices over SIGNAL ids (the layout za_pkfile_write / za_synthesize take)
struct Csr {
    std::vector<uint32_t> ptr[3], sig[3];
    std::vector<uint8_t> coeff[3];
    void from(const Constraints& c) {
        for (int w = 0; w < 3; w++) { ptr[w].assign(1, 0); sig[w].clear(); coeff[w].clear(); }
        for (auto& q : c.q) {
            const LC* l[3] = {&q.a, &q.b, &q.c};
            for (int w = 0; w < 3; w++) {
                for (auto& t : l[w]->t) {
                    if (t.first > 0xfffffffeull) throw FeError("Unexpected", "signal id does not fit 32 bits");
                    sig[w].push_back((uint32_t)t.first);
                    uint8_t b[32];
                    fs_bytes(FS(mod_r(t.second.n)), b);
                    coeff[w].insert(coeff[w].end(), b, b + 32);
                }
                ptr[w].push_back((uint32_t)sig[w].size());
            }
        }
        for (int w = 0; w < 3; w++) { if (sig[w].empty()) sig[w].push_back(0); if (coeff[w].empty()) coeff[w].assign(32, 0); }
    }
    uint32_t rows() const { return (uint32_t)ptr[0].size() - 1; }
};

// CircomCircuit::synthesize (prover.rs:45-103) through za_synthesize, then the device copy of the constraint system
struct Synth {
    std::vector<uint32_t> var_of_signal, var[3];
    std::vector<uint8_t> c_coeff;
    uint32_t ni = 0, na = 0;
    void run(const Signals& signals, const std::vector<uint32_t>& ignore, const Csr& m) {
        const uint32_t n = (uint32_t)signals.len();
        std::vector<uint8_t> is_public(n, 0);
        for (uint32_t i = 1; i < n; i++) is_public[i] = signals.ids[i].is_main_public_input() ? 1 : 0;
        var_of_signal.assign(n, 0);
        for (int w = 0; w < 3; w++) var[w].assign(m.sig[w].size(), 0);
        c_coeff.assign(m.coeff[2].size(), 0);
        const uint32_t* ptr[3] = {m.ptr[0].data(), m.ptr[1].data(), m.ptr[2].data()};
        const uint32_t* sig[3] = {m.sig[0].data(), m.sig[1].data(), m.sig[2].data()};
        uint32_t* ov[3] = {var[0].data(), var[1].data(), var[2].data()};
        static const uint32_t none = 0;
        zcheck(za_synthesize(n, is_public.data(), ignore.empty() ? &none : ignore.data(), (uint32_t)ignore.size(), m.rows(), ptr, sig, m.coeff[2].data(),
                             var_of_signal.data(), ov, c_coeff.data(), &ni, &na), "synthesize");
    }
    void upload(za_ctx* ctx, const Csr& m, za_circuit** out) const {
        za_r1cs cs;
        cs.num_inputs = ni; cs.num_aux = na; cs.num_constraints = m.rows();
        for (int w = 0; w < 3; w++) { cs.ptr[w] = m.ptr[w].data(); cs.var[w] = var[w].data(); }
        cs.coeff[0] = m.coeff[0].data(); cs.coeff[1] = m.coeff[1].data(); cs.coeff[2] = c_coeff.data();
        zcheck(za_circuit_upload(ctx, &cs, out), "circuit upload");
    }
};

std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw FeError("IoError", "Os { code: 2, kind: NotFound, message: \"No such file or directory\" }");
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// ---- flatten_json (format.rs:296-335) ----------------------------------------------------------------------------------
struct Json {
    const std::string& s;
    size_t p = 0;
    int depth = 0;
    explicit Json(const std::string& str) : s(str) {}
    [[noreturn]] void bad(const std::string& what) const { throw FeError("Json", "Error(" + rust_debug_str(what) + ", column: " + std::to_string(p + 1) + ")"); }
    void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) p++; }
    std::string string() {
        std::string out;
        p++;
        while (true) {
            if (p >= s.size()) bad("EOF while parsing a string");
            const char c = s[p++];
            if (c == '"') return out;
            if (c != '\\') { out += c; continue; }
            if (p >= s.size()) bad("EOF while parsing a string");
            const char e = s[p++];
            switch (e) {
                case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break;
                case 'r': out += '\r'; break; case 't': out += '\t'; break;
                case 'u': {
                    if (p + 4 > s.size()) bad("EOF while parsing a string");
                    unsigned cp = 0;
                    for (int i = 0; i < 4; i++) {
                        const char h = s[p++];
                        cp = cp * 16 + (h >= '0' && h <= '9' ? h - '0' : h >= 'a' && h <= 'f' ? h - 'a' + 10 : h >= 'A' && h <= 'F' ? h - 'A' + 10 : 0);
                    }
                    if (cp < 0x80) out += (char)cp;
                    else if (cp < 0x800) { out += (char)(0xc0 | (cp >> 6)); out += (char)(0x80 | (cp & 63)); }
                    else { out += (char)(0xe0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 63)); out += (char)(0x80 | (cp & 63)); }
                    break;
                }
                default: bad("invalid escape");
            }
        }
    }
    void value(const std::string& prefix, std::vector<std::pair<std::string, FS>>& out) {
        if (++depth > 128) bad("recursion limit exceeded");
        ws();
        if (p >= s.size()) bad("EOF while parsing a value");
        const char c = s[p];
        if (c == '[') {
            p++;
            ws();
            size_t i = 0;
            if (p < s.size() && s[p] == ']') p++;
            else
                while (true) {
                    value(prefix + "[" + std::to_string(i++) + "]", out);
                    ws();
                    if (p < s.size() && s[p] == ',') { p++; continue; }
                    if (p < s.size() && s[p] == ']') { p++; break; }
                    bad("expected `,` or `]`");
                }
        } else if (c == '{') {
            p++;
            ws();
            if (p < s.size() && s[p] == '}') p++;
            else
                while (true) {
                    ws();
                    if (p >= s.size() || s[p] != '"') bad("key must be a string");
                    const std::string key = string();
                    ws();
                    if (p >= s.size() || s[p] != ':') bad("expected `:`");
                    p++;
                    value(prefix + "." + key, out);
                    ws();
                    if (p < s.size() && s[p] == ',') { p++; continue; }
                    if (p < s.size() && s[p] == '}') { p++; break; }
                    bad("expected `,` or `}`");
                }
        } else if (c == '"') {
            const std::string v = string();
            out.emplace_back(prefix, FS::parse(v));                     // format.rs:313-317
        } else if (c == '-' || (c >= '0' && c <= '9')) {
            size_t q = p;
            if (s[q] == '-') q++;
            while (q < s.size() && s[q] >= '0' && s[q] <= '9') q++;
            const bool integral = q >= s.size() || (s[q] != '.' && s[q] != 'e' && s[q] != 'E');
            const std::string text = s.substr(p, q - p);
            if (!integral || text[0] == '-' || text.size() > 20) throw FeError("BadFormat", "bad value Number(" + text + ")");      // as_u64, format.rs:318-324
            unsigned __int128 v = 0;
            for (char d : text) v = v * 10 + (unsigned)(d - '0');
            if (v >> 64) throw FeError("BadFormat", "bad value Number(" + text + ")");
            p = q;
            out.emplace_back(prefix, FS::from_u64((uint64_t)v));
        } else {
            throw FeError("BadFormat", "Cannot decode value");          // Bool / Null, format.rs:326
        }
        depth--;
    }
};
std::vector<std::pair<std::string, FS>> flatten_json(const std::string& prefix, const std::string& json) {
    std::vector<std::pair<std::string, FS>> out;
    Json j(json);
    j.value(prefix, out);
    j.ws();
    if (j.p != json.size()) j.bad("trailing characters");
    return out;
}

std::string json_escape(const std::string& s) {
    std::string o;
    for (char c : s) {
        if (c == '"') o += "\\\"";
        else if (c == '\\') o += "\\\\";
        else if (c == '\n') o += "\\n";
        else if ((unsigned char)c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o += c;
    }
    return o;
}

// One device context per process, created on first use (setup and prove are host calls that own no GPU state in the
// reference; here the context holds the stream and the cached NTT domains).
za_ctx* shared_ctx() {
    static CtxHandle h;
    if (!h.p) zcheck(za_ctx_create(device_from_env(), &h.p), "GPU context");
    return h.p;
}

// A proving.key stays loaded between prove() calls while the file is unchanged (the reference re-reads and re-validates it
// on every call, helper.rs:92-98).
struct LoadedKey {
    std::string path;
    long long mtime_ns = -1;
    long long size = -1;
    std::vector<BodyElement> asts;
    Csr csr;
    std::vector<uint32_t> ignore;
    PkHandle pk;
    CircuitHandle circuit;
    Synth synth;
    bool circuit_ready = false;
    size_t n_signals_seen = 0;
};
std::unique_ptr<LoadedKey> g_key;

}  // namespace

extern "C" {

void verbose(int on) {                              // lib.rs:33-49, prover.rs:28-34 (BELLMAN_VERBOSE)
    g_verbose = on != 0;
    if (on) setenv("ZA_DEBUG_TIMELINE", "1", 1); else unsetenv("ZA_DEBUG_TIMELINE");
}

// helper::setup (helper.rs:22-89) + prover::setup (prover.rs:105-137)
static std::string do_setup(const std::string& circuit_path, const std::string& pk_path, bool solidity) {
    Evaluator eval(Mode::GenConstraints);
    info("Compiling circuit...");
    {
        Span t("Compilation time");
        Scope scope(true, nullptr, circuit_path);
        eval.eval_file(scope, ".", circuit_path);
    }
    info("[compile] " + std::to_string(eval.signals.len()) + " signals, " + std::to_string(eval.constraints.len()) + " constraints");
    Constraints constraints;
    std::vector<SignalId> removed;
    {
        Span t("Optimization time");
        optimize(eval.constraints, eval.signals.main_input_ids(), constraints, removed);
    }
    info("Optimize L1 " + std::to_string(constraints.len()) + " " + std::to_string(removed.size()));
    info("[optimized] " + std::to_string(eval.signals.len() - removed.size()) + " signals, " + std::to_string(constraints.len()) + " constraints");
    info("Running setup");
    za_ctx* ctx = shared_ctx();
    Csr csr;
    csr.from(constraints);
    std::vector<uint32_t> ignore(removed.begin(), removed.end());
    Synth syn;
    syn.run(eval.signals, ignore, csr);
    CircuitHandle circuit;
    syn.upload(ctx, csr, &circuit.p);
    std::vector<uint8_t> params;
    {
        Span t("Setup time");
        // generate_random_parameters: g1, g2 random group elements, alpha, beta, gamma, delta, tau random scalars
        uint8_t sc[7][32], g1[64], g2[128];
        for (int i = 0; i < 7; i++) fs_bytes(random_scalar(), sc[i]);
        BasesHandle b1, b2;
        zcheck(za_bases_generate(ctx, 1, 1, 1, &b1.p), "G1 generator");
        zcheck(za_bases_generate(ctx, 2, 1, 1, &b2.p), "G2 generator");
        zcheck(za_multiexp(ctx, b1.p, 0, sc[5], 1, nullptr, g1), "random G1 element");
        zcheck(za_multiexp(ctx, b2.p, 0, sc[6], 1, nullptr, g2), "random G2 element");
        params.resize(za_parameters_max_size(circuit.p));
        size_t len = 0;
        const int rc = za_generate_parameters(ctx, circuit.p, sc[0], sc[1], sc[2], sc[3], sc[4], g1, g2, params.data(), params.size(), &len);
        if (rc == ZA_ERR_UNCONSTRAINED_VARIABLE) throw FeError("Synthesis", "UnconstrainedVariable");
        if (rc == ZA_ERR_POLY_DEGREE_TOO_LARGE) throw FeError("Synthesis", "PolynomialDegreeTooLarge");
        zcheck(rc, "generate_parameters");
        params.resize(len);
    }
    {
        Span t("Proving key write time");
        const std::vector<uint8_t> ast = ast_serialize(eval.collected_asts);
        const uint32_t* ptr[3] = {csr.ptr[0].data(), csr.ptr[1].data(), csr.ptr[2].data()};
        const uint32_t* sig[3] = {csr.sig[0].data(), csr.sig[1].data(), csr.sig[2].data()};
        const uint8_t* coeff[3] = {csr.coeff[0].data(), csr.coeff[1].data(), csr.coeff[2].data()};
        static const uint32_t none = 0;
        size_t need = 0;
        za_pkfile_write(ast.data(), ast.size(), csr.rows(), ptr, sig, coeff, ignore.empty() ? &none : ignore.data(), (uint32_t)ignore.size(), params.data(),
                        params.size(), nullptr, 0, &need);
        std::vector<uint8_t> file(need);
        zcheck(za_pkfile_write(ast.data(), ast.size(), csr.rows(), ptr, sig, coeff, ignore.empty() ? &none : ignore.data(), (uint32_t)ignore.size(),
                               params.data(), params.size(), file.data(), file.size(), &need), "write_pk");
        std::ofstream f(pk_path, std::ios::binary | std::ios::trunc);
        if (!f) throw FeError("IoError", "Os { code: 2, kind: NotFound, message: \"No such file or directory\" }");
        f.write((const char*)file.data(), (std::streamsize)need);
        if (!f) throw FeError("IoError", "Custom { kind: Other, error: \"short write\" }");
    }
    // params.vk and the names of main's public signals (prover.rs:134-136)
    PkHandle pk;
    zcheck(za_pk_load(ctx, params.data(), params.size(), 0, &pk.p), "Parameters::read");
    uint32_t counts[6];
    zcheck(za_pk_counts(pk.p, counts), "pk counts");
    std::vector<uint8_t> vk(576 + 64 * (size_t)counts[0]);
    zcheck(za_pk_vk(pk.p, vk.data(), vk.size()), "vk");
    const std::vector<std::string> inputs = eval.signals.main_public_input_names();
    std::vector<const char*> names;
    for (auto& s : inputs) names.push_back(s.c_str());
    std::string out;
    if (solidity) {
        size_t need = 0;
        char probe = 0;                                  // a first call that only sizes the text
        za_vk_to_solidity(vk.data(), counts[0], names.data(), names.size(), nullptr, &probe, 0, &need);
        out.resize(need + 1);
        zcheck(za_vk_to_solidity(vk.data(), counts[0], names.data(), names.size(), nullptr, &out[0], out.size(), &need), "generate_solidity");
        out.resize(strlen(out.c_str()));
    } else {
        size_t cap = 4096 + 700 * (size_t)counts[0];
        for (auto& s : inputs) cap += s.size() + 8;
        out.resize(cap);
        zcheck(za_vk_to_json(vk.data(), counts[0], names.data(), names.size(), &out[0], out.size()), "vk json");
        out.resize(strlen(out.c_str()));
    }
    return out;
}

int setup(const char* circuit_path, const char* pk_path, const char* verifier_type, char* verifier_buf, size_t verifier_buf_size, char* err_buf,
          size_t err_buf_size) {
    if (!circuit_path || !pk_path || !verifier_type) return return_string("NULL argument", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    if (strcmp(verifier_type, "json") != 0 && strcmp(verifier_type, "solidity") != 0)
        return return_string("invalid validator type", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);      // lib.rs:68
    std::lock_guard<std::mutex> lock(g_mutex);
    try {
        g_key.reset();                              // the file at pk_path is about to change
        const std::string verifier = do_setup(circuit_path, pk_path, strcmp(verifier_type, "solidity") == 0);
        return return_string(verifier, verifier_buf, verifier_buf_size, ZA2C_ERR_NONE);
    } catch (const FeError& e) {
        return return_string(error_debug(e), err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    } catch (const std::exception& e) {
        return return_string(std::string("Unexpected(") + rust_debug_str(e.what()) + ")", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    }
}

// helper::prove (helper.rs:91-147) + generate_verified_proof (prover.rs:139-208)
static std::string do_prove(const std::string& pk_path, const std::vector<std::pair<std::string, FS>>& inputs) {
    za_ctx* ctx = shared_ctx();
    struct stat st;
    if (stat(pk_path.c_str(), &st) != 0) throw FeError("IoError", "Os { code: 2, kind: NotFound, message: \"No such file or directory\" }");
    const long long mtime_ns = (long long)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
    if (!g_key || g_key->path != pk_path || g_key->mtime_ns != mtime_ns || g_key->size != (long long)st.st_size) {
        Span t("Proving key read time");
        std::unique_ptr<LoadedKey> k(new LoadedKey());
        k->path = pk_path; k->mtime_ns = mtime_ns; k->size = (long long)st.st_size;
        const std::string file = read_file(pk_path);
        uint64_t pinfo[6];
        size_t params_off = 0, ast_off = 0, ast_len = 0;
        if (za_pkfile_scan((const uint8_t*)file.data(), file.size(), pinfo, &params_off, &ast_off, &ast_len) != ZA_OK)
            throw FeError("Bincode", std::string("Custom(") + rust_debug_str(za_last_error()) + ")");
        k->asts = ast_deserialize((const uint8_t*)file.data() + ast_off, ast_len);
        Csr& m = k->csr;
        const size_t nc = (size_t)pinfo[0];
        for (int w = 0; w < 3; w++) {
            m.ptr[w].assign(nc + 1, 0);
            m.sig[w].assign(std::max<size_t>((size_t)pinfo[2 + w], 1), 0);
            m.coeff[w].assign(std::max<size_t>((size_t)pinfo[2 + w], 1) * 32, 0);
        }
        k->ignore.assign(std::max<size_t>((size_t)pinfo[1], 1), 0);
        uint32_t* ptr[3] = {m.ptr[0].data(), m.ptr[1].data(), m.ptr[2].data()};
        uint32_t* sig[3] = {m.sig[0].data(), m.sig[1].data(), m.sig[2].data()};
        uint8_t* coeff[3] = {m.coeff[0].data(), m.coeff[1].data(), m.coeff[2].data()};
        zcheck(za_pkfile_read((const uint8_t*)file.data(), file.size(), ptr, sig, coeff, k->ignore.data()), "read_pk");
        k->ignore.resize((size_t)pinfo[1]);
        zcheck(za_pk_load(ctx, (const uint8_t*)file.data() + params_off, file.size() - params_off, 1, &k->pk.p), "Parameters::read");   // format.rs:285 (checked)
        g_key = std::move(k);
    }
    LoadedKey& key = *g_key;

    info("Generating witness...");
    Evaluator ev(Mode::GenWitness);
    {
        Span t("Witness generation time");
        for (auto& kv : inputs) ev.set_deferred_value(kv.first, Value::of(kv.second));
        Scope scope(true, nullptr, "");
        ev.eval_asts(scope, key.asts);
    }
    info("Checking constraints...");
    if (ev.constraints.len() != 0) throw FeError("Unexpected", "Constrains generated in witnes");
    info("Checking signals...");
    for (size_t n = 1; n < ev.signals.len(); n++) {
        const Signal& s = ev.signals.ids[n];
        if (!s.has_value) throw FeError("Unexpected", "signal '" + s.full_name + "' value is not defined");
    }
    info("Creating and self-verifying proof...");
    if (!key.circuit_ready || key.n_signals_seen != ev.signals.len()) {
        key.synth.run(ev.signals, key.ignore, key.csr);
        if (key.circuit.p) { za_circuit_free(key.circuit.p); key.circuit.p = nullptr; }
        key.synth.upload(ctx, key.csr, &key.circuit.p);
        key.circuit_ready = true;
        key.n_signals_seen = ev.signals.len();
    }
    const Synth& syn = key.synth;
    std::vector<uint8_t> in((size_t)syn.ni * 32, 0), aux(std::max<size_t>(syn.na, 1) * 32, 0);
    for (size_t n = 0; n < ev.signals.len(); n++) {
        const uint32_t v = syn.var_of_signal[n];
        if (v == 0xffffffffu) continue;
        const Signal& s = ev.signals.ids[n];
        FS val = FS::one();
        if (n) {
            if (s.value.kind != Value::FieldScalar) throw FeError("Unexpected", "signal '" + s.full_name + "' has no scalar value");   // value_to_bellman_fr panics
            val = s.value.fs;
            if (cmp(val.n, field_r()) >= 0) throw FeError("Unexpected", "signal '" + s.full_name + "' is not a field element");       // Fr::from_str(..).unwrap()
        }
        fs_bytes(val, (v & ZA_VAR_AUX) ? &aux[(size_t)(v & 0x7fffffffu) * 32] : &in[(size_t)v * 32]);
    }
    {
        Span t("Constraint check time");
        int64_t bad = -1;
        zcheck(za_circuit_satisfied(ctx, key.circuit.p, in.data(), aux.data(), &bad), "constraint check");
        if (bad >= 0) throw FeError("Unexpected", "check_constrains_eval_zero failed: constraint " + std::to_string(bad));      // prover.rs:155-157
    }
    uint8_t proof[256];
    {
        Span t("Proof generation time");
        uint8_t r[32], s[32];
        fs_bytes(random_scalar(), r);
        fs_bytes(random_scalar(), s);
        const int rc = za_create_proof(ctx, key.pk.p, key.circuit.p, in.data(), aux.data(), r, s, proof, nullptr);
        if (rc != ZA_OK) throw FeError("Synthesis", za_last_error());
    }
    std::string out;
    {
        Span t("Proof verification time");
        uint32_t counts[6];
        zcheck(za_pk_counts(key.pk.p, counts), "pk counts");
        std::vector<uint8_t> vk(576 + 64 * (size_t)counts[0]);
        zcheck(za_pk_vk(key.pk.p, vk.data(), vk.size()), "vk");
        // public inputs = main's public signals in signal order = bellman inputs 1.. (prover.rs:181-189)
        int valid = 0;
        const int rc = za_verify_proof(vk.data(), counts[0], proof, in.data() + 32, syn.ni - 1, &valid);
        if (rc != ZA_OK) throw FeError("Synthesis", za_last_error());
        out.resize(2048 + 100 * (size_t)syn.ni);
        std::vector<std::string> names;
        for (size_t n = 1; n < ev.signals.len(); n++) if (ev.signals.ids[n].is_main_public_input()) names.push_back(ev.signals.ids[n].full_name);
        (void)names;                                   // JsonProofAndInput stores the values only (format.rs:80-99)
        (void)valid;                                   // the reference ignores verify_proof's bool as well (prover.rs:200 `?` on the Result only)
        zcheck(za_proof_to_json(proof, in.data() + 32, syn.ni - 1, &out[0], out.size()), "proof json");
        out.resize(strlen(out.c_str()));
    }
    return out;
}

int prove(const char* pk_path, const char* inputs_json, char* proof_buf, size_t proof_buf_size, char* err_buf, size_t err_buf_size) {
    if (!pk_path || !inputs_json) return return_string("NULL argument", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    std::lock_guard<std::mutex> lock(g_mutex);
    try {
        const std::vector<std::pair<std::string, FS>> inputs = flatten_json("main", inputs_json);    // lib.rs:97
        const std::string proof = do_prove(pk_path, inputs);
        return return_string(proof, proof_buf, proof_buf_size, ZA2C_ERR_NONE);
    } catch (const FeError& e) {
        return return_string(error_debug(e), err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    } catch (const std::exception& e) {
        return return_string(std::string("Unexpected(") + rust_debug_str(e.what()) + ")", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    }
}

int verify(const char* vk_json, const char* proof_with_inputs_json, char* err_buf, size_t err_buf_size) {
    if (!vk_json || !proof_with_inputs_json) return return_string("NULL argument", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    int valid = 0;
    const int rc = za_verify_json(vk_json, proof_with_inputs_json, &valid);
    if (rc != ZA_OK) return return_string(za_last_error(), err_buf, err_buf_size, ZA2C_ERR_CUSTOM);   // lib.rs:115
    if (g_verbose) fprintf(stderr, "[za2c] verify: %s\n", valid ? "valid" : "not valid");
    return valid ? ZA2C_ERR_NONE : ZA2C_ERR_VERIFICATION_FAILED;                                      // lib.rs:113-114
}

void za2c_release(void) {                           // drops the cached proving key and the device context's users (tests)
    std::lock_guard<std::mutex> lock(g_mutex);
    g_key.reset();
}

// ---- front-end seam for tests and tools (no GPU) ---------------------------------------------------------------------
// what: 0 expression, 1 statement, 2 body element (Debug text, display.rs), 3 whole body -> bincode -> body round trip
// (returns the hex of the image), 4 preprocess
int za2c_parse(int what, const char* text, char* out, size_t out_size, char* err_buf, size_t err_buf_size) {
    if (!text) return return_string("NULL argument", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    try {
        std::string r;
        if (what == 0) r = debug_string(*parse_expression(text));
        else if (what == 1) r = debug_string(*parse_statement(text));
        else if (what == 2) r = debug_string(parse_body_element(text));
        else if (what == 3) {
            const std::vector<BodyElement> body = parse_body(text);
            const std::vector<uint8_t> img = ast_serialize(body);
            const std::vector<BodyElement> back = ast_deserialize(img.data(), img.size());
            if (ast_serialize(back) != img) throw FeError("Unexpected", "bincode round trip differs");
            std::string a, b;
            for (auto& e : body) a += debug_string(e) + "\n";
            for (auto& e : back) b += debug_string(e) + "\n";
            if (a != b) throw FeError("Unexpected", "bincode round trip changes the tree");
            static const char* H = "0123456789abcdef";
            for (uint8_t x : img) { r += H[x >> 4]; r += H[x & 15]; }
        } else if (what == 4) r = preprocess(text);
        else return return_string("bad selector", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
        return return_string(r, out, out_size, ZA2C_ERR_NONE);
    } catch (const FeError& e) {
        return return_string(error_debug(e), err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    }
}

static std::string dump_eval(Evaluator& ev, Scope& scope, bool with_optimized) {
    std::string j = "{\"signals\":[";
    for (size_t i = 0; i < ev.signals.len(); i++) {
        if (i) j += ",";
        j += "\"" + json_escape(ev.signals.to_string(i)) + "\"";
    }
    j += "],\"constraints\":[";
    auto nm = [&](SignalId id) { return ev.signals.name_of(id); };
    for (size_t i = 0; i < ev.constraints.len(); i++) {
        if (i) j += ",";
        j += "\"" + json_escape(ev.constraints.q[i].format(nm)) + "\"";
    }
    j += "],\"scope\":{";
    bool first = true;
    for (auto& kv : scope.vars) {
        if (kv.second.kind == ScopeValue::Function || kv.second.kind == ScopeValue::Template) continue;
        if (!first) j += ",";
        first = false;
        j += "\"" + json_escape(kv.first) + "\":\"" + json_escape(kv.second.debug()) + "\"";
    }
    j += "}";
    if (with_optimized) {
        Constraints oc;
        std::vector<SignalId> removed;
        optimize(ev.constraints, ev.signals.main_input_ids(), oc, removed);
        j += ",\"optimized\":[";
        for (size_t i = 0; i < oc.len(); i++) { if (i) j += ","; j += "\"" + json_escape(oc.q[i].format(nm)) + "\""; }
        j += "],\"removed\":[";
        for (size_t i = 0; i < removed.size(); i++) { if (i) j += ","; j += std::to_string(removed[i]); }
        j += "]";
    }
    j += ",\"ast_bytes\":" + std::to_string(ast_serialize(ev.collected_asts).size()) + "}";
    return j;
}

// mode: 1 = GenConstraints, 2 = GenWitness.  source != NULL: eval_inline (evaluator/test.rs:36-50); else file_path:
// eval_file(".", file_path).  deferred_json: {"main.a": "4", ...} or NULL.  check != 0 (witness mode): also generate the
// constraints of the same text and test them against the witness (test.rs:62-78).  out: JSON dump of signals
// ("name:Type:value"), constraints (QEQ text over signal names), root scope variables (Debug text).
int za2c_eval(int mode, const char* source, const char* file_path, const char* deferred_json, int check, char* out, size_t out_size, char* err_buf,
              size_t err_buf_size) {
    if ((!source && !file_path) || (mode != 1 && mode != 2)) return return_string("bad argument", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    try {
        auto run = [&](Mode m, Evaluator& ev, Scope& scope) {
            (void)m;
            if (deferred_json && ev.mode == Mode::GenWitness)
                for (auto& kv : flatten_json("", deferred_json)) ev.set_deferred_value(kv.first.substr(1), Value::of(kv.second));      // keys are full names
            if (source) ev.eval_inline(scope, source); else ev.eval_file(scope, ".", file_path);
        };
        Evaluator ev(mode == 1 ? Mode::GenConstraints : Mode::GenWitness);
        Scope scope(true, nullptr, "root");
        run(ev.mode, ev, scope);
        if (mode == 2 && check) {
            Evaluator ec(Mode::GenConstraints);
            Scope sc(true, nullptr, "root");
            run(ec.mode, ec, sc);
            const std::string msg = ec.constraints.satisfies_with_signals(ev.signals);
            if (!msg.empty()) throw FeError("Unexpected", msg);
        }
        return return_string(dump_eval(ev, scope, mode == 1), out, out_size, ZA2C_ERR_NONE);
    } catch (const FeError& e) {
        return return_string(error_debug(e), err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    }
}

// `za test` for one file (compiler/src/tester/embeeded.rs:11-121): every #[test] template whose name starts with
// `prefix` is run as a witness, then as constraints, the signal tables are compared and the constraints are evaluated
// on the witness.  out: JSON list of {"name", "signals", "constraints"}.
int za2c_test(const char* file_path, const char* prefix, char* out, size_t out_size, char* err_buf, size_t err_buf_size) {
    if (!file_path) return return_string("NULL argument", err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    try {
        Evaluator scan(Mode::Collect);
        Scope scan_scope(true, nullptr, file_path);
        scan.eval_file(scan_scope, ".", file_path);
        std::vector<std::string> tests;
        for (auto& kv : scan_scope.vars)
            if (kv.second.kind == ScopeValue::Template) {
                bool tagged = false;
                for (auto& a : kv.second.attrs) if (a == "test") tagged = true;
                if (tagged && kv.first.compare(0, strlen(prefix ? prefix : ""), prefix ? prefix : "") == 0) tests.push_back(kv.first);
            }
        std::sort(tests.begin(), tests.end());
        std::string j = "[";
        for (size_t i = 0; i < tests.size(); i++) {
            Evaluator wi(Mode::GenWitness), cn(Mode::GenConstraints);
            { Scope s = scan_scope; wi.eval_template(s, tests[i]); }
            { Scope s = scan_scope; cn.eval_template(s, tests[i]); }
            const size_t upto = std::min(wi.signals.len(), cn.signals.len());
            for (size_t n = 1; n < upto; n++)
                if (wi.signals.ids[n].full_name != cn.signals.ids[n].full_name)
                    throw FeError("Unexpected", "constrain & witness signals differ #cn=" + cn.signals.ids[n].full_name + ",#wi=" + wi.signals.ids[n].full_name);
            if (wi.signals.len() != cn.signals.len()) throw FeError("Unexpected", "constrain & witness signals differ");
            const std::string msg = cn.constraints.satisfies_with_signals(wi.signals);
            if (!msg.empty()) throw FeError("Unexpected", msg);
            if (i) j += ",";
            j += "{\"name\":\"" + json_escape(tests[i]) + "\",\"signals\":" + std::to_string(cn.signals.len()) + ",\"constraints\":" +
                 std::to_string(cn.constraints.len()) + "}";
        }
        j += "]";
        return return_string(j, out, out_size, ZA2C_ERR_NONE);
    } catch (const FeError& e) {
        return return_string(error_debug(e), err_buf, err_buf_size, ZA2C_ERR_CUSTOM);
    }
}

}  // extern "C"
