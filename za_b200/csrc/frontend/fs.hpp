// Field scalars of za's front-end: canonical integers modulo the BN254 scalar field with the operator semantics of
// /root/reference/compiler/src/algebra/fs.rs (FS over num_bigint::BigUint): which operators reduce, which do not, what
// is an error.  Host only; independent of the device field layer (ff.cuh) on purpose — the front-end is a CPU component
// and must compile with plain g++.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

namespace zafe {

struct FeError : std::runtime_error {
    std::string kind;      // variant name of the reference's error enum (evaluator/error.rs, algebra/error.rs)
    std::string text;
    uint64_t meta_start = 0, meta_end = 0;
    bool has_meta = false;
    FeError(const std::string& k, const std::string& t) : std::runtime_error(k + "(\"" + t + "\")"), kind(k), text(t) {}
};
[[noreturn]] inline void fail(const char* kind, const std::string& text) { throw FeError(kind, text); }

struct U256 {
    uint64_t v[4] = {0, 0, 0, 0};
    U256() {}
    explicit U256(uint64_t x) { v[0] = x; }
    bool is_zero() const { return !(v[0] | v[1] | v[2] | v[3]); }
    bool fits_u64() const { return !(v[1] | v[2] | v[3]); }
    int bit(int i) const { return (int)((v[i >> 6] >> (i & 63)) & 1); }
    int bits() const {
        for (int i = 3; i >= 0; i--)
            if (v[i]) return 64 * i + 64 - __builtin_clzll(v[i]);
        return 0;
    }
};
inline int cmp(const U256& a, const U256& b) {
    for (int i = 3; i >= 0; i--)
        if (a.v[i] != b.v[i]) return a.v[i] < b.v[i] ? -1 : 1;
    return 0;
}
inline bool operator==(const U256& a, const U256& b) { return cmp(a, b) == 0; }
inline bool operator<(const U256& a, const U256& b) { return cmp(a, b) < 0; }
inline uint64_t add_to(U256& a, const U256& b) {
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) { c += (unsigned __int128)a.v[i] + b.v[i]; a.v[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
inline uint64_t sub_from(U256& a, const U256& b) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        const uint64_t x = a.v[i], y = b.v[i];
        const uint64_t d = x - y - borrow;
        borrow = (x < y) || (x == y && borrow);
        a.v[i] = d;
    }
    return borrow;
}
inline U256 shl_small(const U256& a, unsigned k) {     // k < 256, bits shifted out are lost
    U256 r;
    const unsigned w = k >> 6, b = k & 63;
    for (int i = 3; i >= 0; i--) {
        uint64_t x = 0;
        if (i >= (int)w) {
            x = a.v[i - w] << b;
            if (b && i - (int)w - 1 >= 0) x |= a.v[i - w - 1] >> (64 - b);
        }
        r.v[i] = x;
    }
    return r;
}
inline U256 shr_small(const U256& a, unsigned k) {
    U256 r;
    const unsigned w = k >> 6, b = k & 63;
    for (int i = 0; i < 4; i++) {
        uint64_t x = 0;
        if (i + w < 4) {
            x = a.v[i + w] >> b;
            if (b && i + w + 1 < 4) x |= a.v[i + w + 1] << (64 - b);
        }
        r.v[i] = x;
    }
    return r;
}
// plain integer division (BigUint `/` and `%`), b != 0
inline void divmod(const U256& a, const U256& b, U256& q, U256& r) {
    q = U256(); r = U256();
    for (int i = a.bits() - 1; i >= 0; i--) {
        r = shl_small(r, 1);
        r.v[0] |= (uint64_t)a.bit(i);
        if (cmp(r, b) >= 0) { sub_from(r, b); q.v[i >> 6] |= 1ull << (i & 63); }
    }
}

// BN254 scalar field modulus, fs.rs:15-16
inline const U256& field_r() {
    static const U256 r = [] {
        U256 x;
        x.v[0] = 0x43e1f593f0000001ull; x.v[1] = 0x2833e84879b97091ull; x.v[2] = 0xb85045b68181585dull; x.v[3] = 0x30644e72e131a029ull;
        return x;
    }();
    return r;
}

// Montgomery machinery used only inside mulmod (values at rest are canonical integers, as in the reference)
struct MontCtx {
    uint64_t inv;      // -r^-1 mod 2^64
    U256 r2;           // 2^512 mod r
    MontCtx() {
        const U256& r = field_r();
        uint64_t x = 1;
        for (int i = 0; i < 6; i++) x *= 2 - r.v[0] * x;      // Newton: r^-1 mod 2^64
        inv = (uint64_t)0 - x;
        U256 t(1);
        for (int i = 0; i < 512; i++) {
            const uint64_t top = t.v[3] >> 63;
            t = shl_small(t, 1);
            if (top || cmp(t, r) >= 0) sub_from(t, r);
        }
        r2 = t;
    }
    U256 mul(const U256& a, const U256& b) const {           // a b 2^-256 mod r (CIOS)
        const U256& r = field_r();
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            unsigned __int128 c = 0;
            for (int j = 0; j < 4; j++) { c += (unsigned __int128)a.v[j] * b.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
            const uint64_t m = t[0] * inv;
            c = (unsigned __int128)m * r.v[0] + t[0];
            c >>= 64;
            for (int j = 1; j < 4; j++) { c += (unsigned __int128)m * r.v[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
        }
        U256 out;
        memcpy(out.v, t, 32);
        if (t[4] || cmp(out, r) >= 0) sub_from(out, r);
        return out;
    }
};
inline const MontCtx& mont() { static const MontCtx m; return m; }

inline U256 reduce_once(U256 a, uint64_t carry = 0) {       // a + carry 2^256 < 2 r
    if (carry || cmp(a, field_r()) >= 0) sub_from(a, field_r());
    return a;
}
inline U256 mod_r(const U256& a) {                          // any 256-bit value: 2^256 < 6 r, so at most five subtractions
    U256 x = a;                                             // (this was a 256-step long division: half of the evaluator's time)
    while (cmp(x, field_r()) >= 0) sub_from(x, field_r());
    return x;
}
inline U256 addmod(const U256& a, const U256& b) { U256 s = a; const uint64_t c = add_to(s, b); return reduce_once(s, c); }
inline U256 negmod(const U256& a) { if (a.is_zero()) return a; U256 s = field_r(); sub_from(s, a); return s; }
inline U256 mulmod(const U256& a, const U256& b) { return mont().mul(mont().mul(a, b), mont().r2); }
// a^-1 mod r for 0 < a < r by the binary extended Euclid (r is prime and odd): shifts and additions only, ~15x cheaper than
// a^(r-2) through mulmod — the divisions of a witness (every BabyAdd has two) were 60 % of the evaluator's time on eddsamimc.
inline U256 invmod(const U256& a) {
    const U256& r = field_r();
    auto half_mod = [&](U256& x) {                          // x / 2 mod r
        uint64_t carry = 0;
        if (x.v[0] & 1) carry = add_to(x, r);
        x = shr_small(x, 1);
        if (carry) x.v[3] |= 1ull << 63;
    };
    auto sub_mod = [&](U256& x, const U256& y) {            // x - y mod r, both below r
        if (cmp(x, y) >= 0) sub_from(x, y);
        else { U256 t = r; sub_from(t, y); add_to(x, t); }  // x + (r - y) < r
    };
    U256 u = a, v = r, x1(1), x2;
    const U256 one(1);
    while (!(u == one) && !(v == one)) {
        while (!(u.v[0] & 1)) { u = shr_small(u, 1); half_mod(x1); }
        while (!(v.v[0] & 1)) { v = shr_small(v, 1); half_mod(x2); }
        if (cmp(u, v) >= 0) { sub_from(u, v); sub_mod(x1, x2); }
        else { sub_from(v, u); sub_mod(x2, x1); }
    }
    return u == one ? x1 : x2;
}
inline U256 powmod(const U256& a, const U256& e) {
    U256 acc(1);
    for (int i = e.bits() - 1; i >= 0; i--) { acc = mulmod(acc, acc); if (e.bit(i)) acc = mulmod(acc, a); }
    return acc;
}

// Arbitrary-size non-negative integers as the u32 digit vectors num-bigint serialises (little endian, no leading zeros):
// source literals (lang.lalrpop DECNUMBER / HEXNUMBER) and the bincode image of FS / BigInt.
struct BigDigits {
    std::vector<uint32_t> d;
    void mul_add_small(uint32_t m, uint32_t a) {
        uint64_t c = a;
        for (auto& x : d) { c += (uint64_t)x * m; x = (uint32_t)c; c >>= 32; }
        if (c) d.push_back((uint32_t)c);
    }
    static BigDigits parse(const std::string& s, int base) {      // digits already validated by the lexer
        BigDigits b;
        for (char ch : s) {
            const uint32_t v = ch >= '0' && ch <= '9' ? ch - '0' : ch >= 'a' && ch <= 'f' ? ch - 'a' + 10 : ch >= 'A' && ch <= 'F' ? ch - 'A' + 10 : 99;
            if (v >= (uint32_t)base) fail("InvalidFormat", s + (base == 16 ? " is not hexadecimal" : " is not decimal"));
            b.mul_add_small((uint32_t)base, v);
        }
        return b;
    }
    static BigDigits from_u256(const U256& a) {
        BigDigits b;
        for (int i = 0; i < 4; i++) { b.d.push_back((uint32_t)a.v[i]); b.d.push_back((uint32_t)(a.v[i] >> 32)); }
        while (!b.d.empty() && b.d.back() == 0) b.d.pop_back();
        return b;
    }
    bool fits_u256() const { return d.size() <= 8; }
    U256 to_u256() const {                                         // low 256 bits
        U256 a;
        for (size_t i = 0; i < d.size() && i < 8; i++) a.v[i >> 1] |= (uint64_t)d[i] << (32 * (i & 1));
        return a;
    }
    U256 mod_field() const {                                       // FS::from(&BigInt): n % r
        if (fits_u256()) return mod_r(to_u256());
        U256 acc, two32(1ull << 32);
        for (size_t i = d.size(); i-- > 0;) acc = addmod(mulmod(acc, two32), U256(d[i]));
        return acc;
    }
    std::string to_decimal() const {
        if (d.empty()) return "0";
        std::vector<uint32_t> t = d;
        std::string out;
        while (!t.empty()) {
            uint64_t rem = 0;
            for (size_t i = t.size(); i-- > 0;) { const uint64_t cur = (rem << 32) | t[i]; t[i] = (uint32_t)(cur / 1000000000u); rem = cur % 1000000000u; }
            while (!t.empty() && t.back() == 0) t.pop_back();
            char buf[16];
            snprintf(buf, sizeof buf, t.empty() ? "%u" : "%09u", (unsigned)rem);
            out.insert(0, buf);
        }
        return out;
    }
};

// FS: fs.rs:33-34.  The value is NOT always below the modulus in the reference (FS::parse and `&` / `%` / `>>` / intdiv
// build FS(v) directly); every such value is still below 2^256 here because the operands are.
struct FS {
    U256 n;
    FS() {}
    explicit FS(const U256& x) : n(x) {}
    static FS from_u64(uint64_t x) { return FS(U256(x)); }               // FS::from(u64): x % r = x
    static FS reduced(const U256& x) { return FS(mod_r(x)); }            // FS::from(BigUint)
    static FS zero() { return FS(); }
    static FS one() { return from_u64(1); }
    bool is_zero() const { return n.is_zero(); }
    bool is_one() const { return n == U256(1); }
    // fs.rs:72-74: greater than (r - 1) / 2
    bool is_neg() const {
        U256 h = field_r();
        h.v[0] -= 1;
        h = shr_small(h, 1);
        return cmp(n, h) > 0;
    }
    std::string to_string() const { return BigDigits::from_u256(n).to_decimal(); }
    // fs.rs:78-86
    std::string format(bool plus_sign_at_start) const {
        if (is_neg()) return "-" + neg().to_string();
        return (plus_sign_at_start ? "+" : "") + to_string();
    }
    bool try_to_u64(uint64_t& out) const { if (!n.fits_u64()) return false; out = n.v[0]; return true; }
    // fs.rs:46-60 FS::parse: "0x" hexadecimal or decimal, stored WITHOUT reduction; a value that does not fit 256 bits
    // cannot be a field element (the reference fails later, in Fr::from_str) and is rejected here.
    static FS parse(const std::string& expr) {
        const bool hex = expr.size() >= 2 && expr[0] == '0' && expr[1] == 'x';
        const std::string digits = hex ? expr.substr(2) : expr;
        if (digits.empty()) fail("InvalidFormat", expr + (hex ? " is not hexadecimal" : " is not decimal"));
        BigDigits b = BigDigits::parse(digits, hex ? 16 : 10);
        if (!b.fits_u256()) fail("InvalidFormat", expr + " does not fit a field element");
        return FS(b.to_u256());
    }
    FS neg() const {                                                     // fs.rs:207-213: FS::from(r - x)
        if (cmp(n, field_r()) > 0) fail("InvalidOperation", "negation of a value above the field modulus");
        U256 s = field_r();
        sub_from(s, n);
        return reduced(s);
    }
    FS add(const FS& o) const {                                          // fs.rs:216-222
        U256 s = n;
        const uint64_t c = add_to(s, o.n);
        if (!c) return reduced(s);
        // 257-bit sum: reduce 2^256 + s
        U256 two256 = mod_r(negate256(field_r()));                      // 2^256 - r, then mod r
        return FS(addmod(mod_r(s), two256));
    }
    FS mul(const FS& o) const { return FS(mulmod(mod_r(n), mod_r(o.n))); }   // fs.rs:225-231
    FS div(const FS& o) const {                                          // fs.rs:234-254: a * b^-1, gcd(b, r) must be 1
        const U256 b = mod_r(o.n);
        if (o.n.is_zero()) fail("InvalidOperation", "Cannot find inv gcd=" + FS(field_r()).to_string());
        if (b.is_zero()) fail("InvalidOperation", "Cannot find inv gcd=" + FS(field_r()).to_string());
        return FS(mulmod(mod_r(n), invmod(b)));
    }
    FS intdiv(const FS& o) const {                                       // fs.rs:116-118 (BigUint division; by zero it panics)
        if (o.n.is_zero()) fail("InvalidOperation", "Divison by zero");
        U256 q, r;
        divmod(n, o.n, q, r);
        return reduced(q);
    }
    FS rem(const FS& o) const {                                          // fs.rs:265-274, not reduced
        if (o.n.is_zero()) fail("InvalidOperation", "Divison by zero");
        U256 q, r;
        divmod(n, o.n, q, r);
        return FS(r);
    }
    FS shl(const FS& o) const {                                          // fs.rs:277-288: (x << k) mod r
        uint64_t k;
        if (!o.try_to_u64(k)) fail("InvalidOperation", "Only can shl on 64 bit values");
        U256 two(2);
        return FS(mulmod(mod_r(n), powmod(two, U256(k))));
    }
    FS shr(const FS& o) const {                                          // fs.rs:291-302
        uint64_t k;
        if (!o.try_to_u64(k)) fail("InvalidOperation", "Only can shr on 64 bit values");
        return reduced(k >= 256 ? U256() : shr_small(n, (unsigned)k));
    }
    FS bit_and(const FS& o) const { FS r; for (int i = 0; i < 4; i++) r.n.v[i] = n.v[i] & o.n.v[i]; return r; }          // fs.rs:305-310, not reduced
    FS bit_or(const FS& o) const { U256 r; for (int i = 0; i < 4; i++) r.v[i] = n.v[i] | o.n.v[i]; return reduced(r); }  // fs.rs:313-318
    FS bit_xor(const FS& o) const { U256 r; for (int i = 0; i < 4; i++) r.v[i] = n.v[i] ^ o.n.v[i]; return reduced(r); } // fs.rs:321-326
    FS pow(const FS& o) const { return FS(powmod(mod_r(n), o.n)); }      // fs.rs:112-114

   private:
    static U256 negate256(const U256& a) { U256 z; sub_from(z, a); return z; }      // 2^256 - a
};
inline bool operator==(const FS& a, const FS& b) { return a.n == b.n; }
inline bool operator!=(const FS& a, const FS& b) { return !(a.n == b.n); }

}  // namespace zafe
