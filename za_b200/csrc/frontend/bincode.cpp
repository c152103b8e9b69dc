// bincode 1.x image of Vec<BodyElementP> — the first section of a proving.key (format.rs:231-234, read back at
// format.rs:262-266).  Layout rules of bincode's default configuration as serde derives them for ast.rs: little endian,
// fixed-width integers, usize and sequence / string lengths as u64, enum variants as u32 in declaration order, Option
// as one tag byte, Box and newtype structs transparent, struct fields in declaration order.  num-bigint's BigInt is the
// tuple (Sign as i8: -1 / 0 / 1, BigUint as a sequence of u32 digits, least significant first).
#include "ast.hpp"

namespace zafe {
namespace {

struct Writer {
    std::vector<uint8_t> out;
    void u8(uint8_t v) { out.push_back(v); }
    void u32(uint32_t v) { for (int i = 0; i < 4; i++) out.push_back((uint8_t)(v >> (8 * i))); }
    void u64(uint64_t v) { for (int i = 0; i < 8; i++) out.push_back((uint8_t)(v >> (8 * i))); }
    void str(const std::string& s) { u64(s.size()); out.insert(out.end(), s.begin(), s.end()); }
    void strs(const std::vector<std::string>& v) { u64(v.size()); for (auto& s : v) str(s); }
    void meta(const Meta& m) { u64(m.start); u64(m.end); strs(m.attrs); }
    void bigint(const BigDigits& b) {
        u8(b.d.empty() ? 0 : 1);                    // Sign::NoSign / Sign::Plus
        u64(b.d.size());
        for (uint32_t d : b.d) u32(d);
    }
    void opcode(Opcode op) { u32((uint32_t)op); }
    void variable(const Variable& v) {
        meta(v.meta);
        str(v.name);
        u64(v.sels.size());
        for (auto& s : v.sels) {
            u32(s.is_pin ? 0 : 1);
            meta(s.meta);
            if (s.is_pin) str(s.name); else expr(*s.pos);
        }
    }
    void exprs(const std::vector<ExprP>& l) { u64(l.size()); for (auto& e : l) expr(*e); }
    void expr(const Expression& e) {
        u32((uint32_t)e.kind);
        meta(e.meta);
        switch (e.kind) {
            case ExprKind::FunctionCall: str(e.name); exprs(e.list); break;
            case ExprKind::Variable: variable(*e.var); break;
            case ExprKind::Number: bigint(e.number); break;
            case ExprKind::PrefixOp: opcode(e.op); expr(*e.rhe); break;
            case ExprKind::InfixOp: expr(*e.lhe); opcode(e.op); expr(*e.rhe); break;
            case ExprKind::Array: exprs(e.list); break;
        }
    }
    void vartype(const VariableType& t) {
        u32((uint32_t)t.kind);
        if (t.kind == VarKind::Signal) u32((uint32_t)t.signal);
    }
    void stmt(const Statement& s) {
        u32((uint32_t)s.kind);
        meta(s.meta);
        switch (s.kind) {
            case StmtKind::IfThenElse:
                expr(*s.cond); stmt(*s.xthen);
                if (s.xelse) { u8(1); stmt(*s.xelse); } else u8(0);
                break;
            case StmtKind::For: stmt(*s.init); expr(*s.cond); stmt(*s.step); stmt(*s.body); break;
            case StmtKind::While: expr(*s.cond); stmt(*s.body); break;
            case StmtKind::Return: expr(*s.value); break;
            case StmtKind::Declaration:
                vartype(s.xtype); variable(*s.name);
                if (s.has_init) { u8(1); opcode(s.op); expr(*s.value); } else u8(0);
                break;
            case StmtKind::Substitution: variable(*s.name); opcode(s.op); expr(*s.value); break;
            case StmtKind::Block: u64(s.stmts.size()); for (auto& t : s.stmts) stmt(*t); break;
            case StmtKind::SignalLeft: variable(*s.name); opcode(s.op); expr(*s.value); break;
            case StmtKind::SignalRight: expr(*s.value); opcode(s.op); variable(*s.name); break;
            case StmtKind::SignalEq: expr(*s.lhe); opcode(s.op); expr(*s.value); break;
            case StmtKind::InternalCall: str(s.call_name); exprs(s.args); break;
        }
    }
    void body(const BodyElement& b) {
        u32((uint32_t)b.kind);
        meta(b.meta);
        switch (b.kind) {
            case BodyKind::Include: str(b.path); break;
            case BodyKind::FunctionDef:
            case BodyKind::TemplateDef: str(b.name); strs(b.args); stmt(*b.stmt); break;
            case BodyKind::Declaration: stmt(*b.stmt); break;
        }
    }
};

struct Reader {
    const uint8_t* p;
    size_t n, at = 0;
    int depth = 0;
    [[noreturn]] void bad(const char* what) const { fail("Bincode", std::string("proving key AST section: ") + what + " at byte " + std::to_string(at)); }
    void need(size_t k) const { if (k > n - at) bad("unexpected end of data"); }
    uint8_t u8() { need(1); return p[at++]; }
    uint32_t u32() { need(4); uint32_t v = 0; for (int i = 0; i < 4; i++) v |= (uint32_t)p[at + i] << (8 * i); at += 4; return v; }
    uint64_t u64() { need(8); uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)p[at + i] << (8 * i); at += 8; return v; }
    uint64_t len(size_t min_elem_bytes) { const uint64_t l = u64(); if (l > (n - at) / (min_elem_bytes ? min_elem_bytes : 1)) bad("length exceeds the data"); return l; }
    std::string str() { const uint64_t l = len(1); std::string s((const char*)p + at, (size_t)l); at += (size_t)l; return s; }
    std::vector<std::string> strs() { const uint64_t l = len(8); std::vector<std::string> v; for (uint64_t i = 0; i < l; i++) v.push_back(str()); return v; }
    Meta meta() { Meta m; m.start = u64(); m.end = u64(); m.attrs = strs(); return m; }
    struct Depth { Reader& r; explicit Depth(Reader& rr) : r(rr) { if (++r.depth > 2000) r.bad("nesting too deep"); } ~Depth() { r.depth--; } };
    BigDigits bigint() {
        const int8_t sign = (int8_t)u8();
        if (sign != 0 && sign != 1) bad(sign == -1 ? "negative literal (the grammar cannot produce one)" : "bad BigInt sign");
        BigDigits b;
        const uint64_t l = len(4);
        for (uint64_t i = 0; i < l; i++) b.d.push_back(u32());
        while (!b.d.empty() && b.d.back() == 0) b.d.pop_back();
        return b;
    }
    Opcode opcode() { const uint32_t v = u32(); if (v >= (uint32_t)Opcode::COUNT) bad("bad Opcode"); return (Opcode)v; }
    VarP variable() {
        Depth d(*this);
        Variable v;
        v.meta = meta();
        v.name = str();
        const uint64_t l = len(4);
        for (uint64_t i = 0; i < l; i++) {
            Selector s;
            const uint32_t k = u32();
            if (k > 1) bad("bad SelectorP");
            s.is_pin = k == 0;
            s.meta = meta();
            if (s.is_pin) s.name = str(); else s.pos = expr();
            v.sels.push_back(std::move(s));
        }
        return std::make_shared<const Variable>(std::move(v));
    }
    std::vector<ExprP> exprs() { const uint64_t l = len(4); std::vector<ExprP> v; for (uint64_t i = 0; i < l; i++) v.push_back(expr()); return v; }
    ExprP expr() {
        Depth d(*this);
        Expression e;
        const uint32_t k = u32();
        if (k > 5) bad("bad ExpressionP");
        e.kind = (ExprKind)k;
        e.meta = meta();
        switch (e.kind) {
            case ExprKind::FunctionCall: e.name = str(); e.list = exprs(); break;
            case ExprKind::Variable: e.var = variable(); break;
            case ExprKind::Number: e.number = bigint(); break;
            case ExprKind::PrefixOp: e.op = opcode(); e.rhe = expr(); break;
            case ExprKind::InfixOp: e.lhe = expr(); e.op = opcode(); e.rhe = expr(); break;
            case ExprKind::Array: e.list = exprs(); break;
        }
        return std::make_shared<const Expression>(std::move(e));
    }
    VariableType vartype() {
        VariableType t;
        const uint32_t k = u32();
        if (k > 3) bad("bad VariableType");
        t.kind = (VarKind)k;
        if (t.kind == VarKind::Signal) { const uint32_t s = u32(); if (s > 3) bad("bad SignalType"); t.signal = (SignalType)s; }
        return t;
    }
    StmtP stmt() {
        Depth d(*this);
        Statement s;
        const uint32_t k = u32();
        if (k > 10) bad("bad StatementP");
        s.kind = (StmtKind)k;
        s.meta = meta();
        switch (s.kind) {
            case StmtKind::IfThenElse: {
                s.cond = expr(); s.xthen = stmt();
                const uint8_t t = u8();
                if (t > 1) bad("bad Option tag");
                if (t) s.xelse = stmt();
                break;
            }
            case StmtKind::For: s.init = stmt(); s.cond = expr(); s.step = stmt(); s.body = stmt(); break;
            case StmtKind::While: s.cond = expr(); s.body = stmt(); break;
            case StmtKind::Return: s.value = expr(); break;
            case StmtKind::Declaration: {
                s.xtype = vartype(); s.name = variable();
                const uint8_t t = u8();
                if (t > 1) bad("bad Option tag");
                if (t) { s.has_init = true; s.op = opcode(); s.value = expr(); }
                break;
            }
            case StmtKind::Substitution: s.name = variable(); s.op = opcode(); s.value = expr(); break;
            case StmtKind::Block: { const uint64_t l = len(4); for (uint64_t i = 0; i < l; i++) s.stmts.push_back(stmt()); break; }
            case StmtKind::SignalLeft: s.name = variable(); s.op = opcode(); s.value = expr(); break;
            case StmtKind::SignalRight: s.value = expr(); s.op = opcode(); s.name = variable(); break;
            case StmtKind::SignalEq: s.lhe = expr(); s.op = opcode(); s.value = expr(); break;
            case StmtKind::InternalCall: s.call_name = str(); s.args = exprs(); break;
        }
        return std::make_shared<const Statement>(std::move(s));
    }
    BodyElement body() {
        BodyElement b;
        const uint32_t k = u32();
        if (k > 3) bad("bad BodyElementP");
        b.kind = (BodyKind)k;
        b.meta = meta();
        switch (b.kind) {
            case BodyKind::Include: b.path = str(); break;
            case BodyKind::FunctionDef:
            case BodyKind::TemplateDef: b.name = str(); b.args = strs(); b.stmt = stmt(); break;
            case BodyKind::Declaration: b.stmt = stmt(); break;
        }
        return b;
    }
};

}  // namespace

std::vector<uint8_t> ast_serialize(const std::vector<BodyElement>& body) {
    Writer w;
    w.u64(body.size());
    for (auto& b : body) w.body(b);
    return std::move(w.out);
}

std::vector<BodyElement> ast_deserialize(const uint8_t* data, size_t len) {
    Reader r{data, len};
    std::vector<BodyElement> out;
    const uint64_t n = r.len(4);
    for (uint64_t i = 0; i < n; i++) out.push_back(r.body());
    if (r.at != len) r.bad("trailing bytes");
    return out;
}

}  // namespace zafe
