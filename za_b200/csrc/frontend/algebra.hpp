// Linear combinations, quadratic equations and the value algebra of za's evaluator:
// /root/reference/compiler/src/algebra/{lc.rs, qeq.rs, value.rs}.  Term ORDER inside a linear combination is part of
// the behaviour (it is what the constraint text, the proving-key file and the optimiser see), so every operation keeps
// the reference's order: terms stay where they first appeared, new ones are appended, zero coefficients are dropped
// only where the reference drops them.
#pragma once
#include <functional>
#include <utility>
#include "fs.hpp"

namespace zafe {

typedef uint64_t SignalId;
static const SignalId SIGNAL_ONE = 0;

struct LC {
    std::vector<std::pair<SignalId, FS>> t;
    static LC from_signal(SignalId s, const FS& f) { LC l; l.t.emplace_back(s, f); return l; }
    static LC from_fs(const FS& f) { return from_signal(SIGNAL_ONE, f); }            // lc.rs:103-107
    bool is_zero() const { for (auto& e : t) if (!e.second.is_zero()) return false; return true; }   // lc.rs:94-100
    LC neg() const { LC r; for (auto& e : t) r.t.emplace_back(e.first, e.second.neg()); return r; }  // lc.rs:109-115
    LC add_fs(const FS& f) const {                                                   // lc.rs:117-131
        LC r = *this;
        bool found = false;
        for (auto& e : r.t) if (e.first == SIGNAL_ONE) { e.second = e.second.add(f); found = true; break; }
        if (!found) r.t.emplace_back(SIGNAL_ONE, f);
        r.retain_nonzero();
        return r;
    }
    LC mul_fs(const FS& f) const {                                                   // lc.rs:133-143
        LC r;
        if (f.is_zero()) return r;
        for (auto& e : t) r.t.emplace_back(e.first, e.second.mul(f));
        return r;
    }
    LC add_lc(const LC& o) const {                                                   // lc.rs:145-160
        LC r = *this;
        for (auto& e : o.t) {
            bool found = false;
            for (auto& x : r.t) if (x.first == e.first) { x.second = x.second.add(e.second); found = true; break; }
            if (!found) r.t.emplace_back(e.first, e.second);
        }
        r.retain_nonzero();
        return r;
    }
    std::string format(const std::function<std::string(SignalId)>& name) const {     // lc.rs:40-55
        if (t.empty()) return "0";
        std::string s = t[0].second.format(false) + name(t[0].first);
        for (size_t i = 1; i < t.size(); i++) s += t[i].second.format(true) + name(t[i].first);
        return s;
    }
   private:
    void retain_nonzero() {
        size_t w = 0;
        for (size_t i = 0; i < t.size(); i++) if (!t[i].second.is_zero()) { if (w != i) t[w] = t[i]; w++; }
        t.resize(w);
    }
};

struct QEQ {                                                                         // a * b + c = 0
    LC a, b, c;
    bool is_zero() const { return (a.is_zero() || b.is_zero()) && c.is_zero(); }     // qeq.rs:43-45
    QEQ add_fs(const FS& f) const { QEQ r = *this; r.c = c.add_fs(f); return r; }    // qeq.rs:60-70
    QEQ mul_fs(const FS& f) const { QEQ r; r.a = a.mul_fs(f); r.b = b; r.c = c.mul_fs(f); return r; }   // qeq.rs:72-82
    QEQ add_lc(const LC& l) const { QEQ r = *this; r.c = c.add_lc(l); return r; }    // qeq.rs:84-94
    QEQ neg() const { QEQ r; r.a = a.neg(); r.b = b; r.c = c.neg(); return r; }      // qeq.rs:96-106
    std::string format(const std::function<std::string(SignalId)>& name) const {     // qeq.rs:20-32
        auto f = [&](const LC& v) { return v.t.empty() ? std::string(" ") : v.format(name); };
        return "[" + f(a) + "]*[" + f(b) + "]+[" + f(c) + "]";
    }
};

struct Value {                                                                       // value.rs:12-17
    enum Kind { FieldScalar, LinearCombination, QuadraticEquation } kind = FieldScalar;
    FS fs;
    LC lc;
    QEQ qeq;
    Value() {}
    static Value of(const FS& f) { Value v; v.kind = FieldScalar; v.fs = f; return v; }
    static Value of(const LC& l) { Value v; v.kind = LinearCombination; v.lc = l; return v; }
    static Value of(const QEQ& q) { Value v; v.kind = QuadraticEquation; v.qeq = q; return v; }
    static Value from_signal(SignalId s) { return of(LC::from_signal(s, FS::one())); }
    QEQ into_qeq() const {                                                           // value.rs:23-30
        QEQ q;
        if (kind == FieldScalar) q.c = LC::from_fs(fs);
        else if (kind == LinearCombination) q.c = lc;
        else q = qeq;
        return q;
    }
    std::string to_string() const {
        auto nm = [](SignalId s) { return "s" + std::to_string(s); };
        return kind == FieldScalar ? fs.to_string() : kind == LinearCombination ? lc.format(nm) : qeq.format(nm);
    }
};

enum class Opcode : uint32_t {                                                       // parser/src/ast.rs Opcode, in declaration order
    Mul, Div, Add, Sub, Pow, IntDiv, Mod, ShiftL, ShiftR, LesserEq, GreaterEq, Lesser, Greater, Eq, NotEq, BoolOr, BoolAnd, BoolNot,
    BitOr, BitAnd, BitXor, Assig, AssigAdd, AssigSub, AssigMul, AssigDiv, AssigMod, AssigShiftL, AssigShiftR, AssigBitAnd, AssigBitOr,
    AssigBitXor, SignalWireLeft, SignalWireRight, SignalContrainLeft, SignalContrainRight, SignalContrainEq, COUNT
};
inline const char* opcode_text(Opcode op) {                                          // display.rs Debug for Opcode
    static const char* T[] = {"*", "/", "+", "-", "**", "\\", "%", "<<", ">>", "<=", ">=", "<", ">", "==", "!=", "||", "&&", "!", "|", "&", "^",
                              "=", "+=", "-=", "*=", "/=", "%=", "<<=", ">>=", "&=", "|=", "^=", "<--", "-->", "<==", "==>", "==="};
    return T[(uint32_t)op];
}
inline const char* opcode_name(Opcode op) {                                          // derive(Debug) is not used for Opcode; names for messages
    return opcode_text(op);
}

// value.rs:116-181
inline Value eval_infix(const Value& l, Opcode op, const Value& r) {
    typedef Value V;
    const int FSk = V::FieldScalar, LCk = V::LinearCombination, QQk = V::QuadraticEquation;
    const int lk = l.kind, rk = r.kind;
    switch (op) {
        case Opcode::Add:
            if (lk == FSk && rk == FSk) return V::of(l.fs.add(r.fs));
            if (lk == LCk && rk == LCk) return V::of(l.lc.add_lc(r.lc));
            if (lk == FSk && rk == LCk) return V::of(r.lc.add_fs(l.fs));
            if (lk == LCk && rk == FSk) return V::of(l.lc.add_fs(r.fs));
            if (lk == FSk && rk == QQk) return V::of(r.qeq.add_fs(l.fs));
            if (lk == QQk && rk == FSk) return V::of(l.qeq.add_fs(r.fs));
            if (lk == LCk && rk == QQk) return V::of(r.qeq.add_lc(l.lc));
            if (lk == QQk && rk == LCk) return V::of(l.qeq.add_lc(r.lc));
            break;
        case Opcode::Sub:
            if (lk == FSk && rk == FSk) return V::of(l.fs.add(r.fs.neg()));
            if (lk == LCk && rk == LCk) return V::of(l.lc.add_lc(r.lc.neg()));
            if (lk == FSk && rk == LCk) return V::of(r.lc.neg().add_fs(l.fs));
            if (lk == LCk && rk == FSk) return V::of(l.lc.add_fs(r.fs.neg()));
            if (lk == FSk && rk == QQk) return V::of(r.qeq.neg().add_fs(l.fs));
            if (lk == QQk && rk == FSk) return V::of(l.qeq.add_fs(r.fs.neg()));
            if (lk == LCk && rk == QQk) return V::of(r.qeq.neg().add_lc(l.lc));
            if (lk == QQk && rk == LCk) return V::of(l.qeq.add_lc(r.lc.neg()));
            break;
        case Opcode::Mul:
            if (lk == FSk && rk == FSk) return V::of(l.fs.mul(r.fs));
            if (lk == LCk && rk == LCk) { QEQ q; q.a = l.lc; q.b = r.lc; return V::of(q); }
            if (lk == LCk && rk == FSk) return V::of(l.lc.mul_fs(r.fs));
            if (lk == FSk && rk == LCk) return V::of(r.lc.mul_fs(l.fs));
            if (lk == QQk && rk == FSk) return V::of(l.qeq.mul_fs(r.fs));
            if (lk == FSk && rk == QQk) return V::of(r.qeq.mul_fs(l.fs));
            break;
        case Opcode::Div: if (lk == FSk && rk == FSk) return V::of(l.fs.div(r.fs)); break;
        case Opcode::IntDiv: if (lk == FSk && rk == FSk) return V::of(l.fs.intdiv(r.fs)); break;
        case Opcode::Mod: if (lk == FSk && rk == FSk) return V::of(l.fs.rem(r.fs)); break;
        case Opcode::ShiftL: if (lk == FSk && rk == FSk) return V::of(l.fs.shl(r.fs)); break;
        case Opcode::ShiftR: if (lk == FSk && rk == FSk) return V::of(l.fs.shr(r.fs)); break;
        case Opcode::BitAnd: if (lk == FSk && rk == FSk) return V::of(l.fs.bit_and(r.fs)); break;
        case Opcode::BitOr: if (lk == FSk && rk == FSk) return V::of(l.fs.bit_or(r.fs)); break;
        case Opcode::BitXor: if (lk == FSk && rk == FSk) return V::of(l.fs.bit_xor(r.fs)); break;
        case Opcode::Pow: if (lk == FSk && rk == FSk) return V::of(l.fs.pow(r.fs)); break;
        default: break;
    }
    fail("InvalidOperation", std::string("Cannot apply operator ") + opcode_text(op) + " on " + l.to_string() + " over " + r.to_string());
}
// value.rs:183-197
inline Value eval_prefix(Opcode op, const Value& r) {
    if (op == Opcode::Sub) {
        if (r.kind == Value::FieldScalar) return Value::of(r.fs.neg());
        if (r.kind == Value::LinearCombination) return Value::of(r.lc.neg());
        return Value::of(r.qeq.neg());
    }
    fail("InvalidOperation", std::string("Cannot apply operator ") + opcode_text(op) + " on " + r.to_string());
}

}  // namespace zafe
