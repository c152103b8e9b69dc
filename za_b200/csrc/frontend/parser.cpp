// Parser of za's circuit language.  Reference: /root/reference/parser/src/parse.rs (comment preprocessor),
// /root/reference/parser/src/lang.lalrpop (grammar, operator tiers, token set), /root/reference/parser/src/display.rs
// (the Debug text of the tree, which the reference's own parser tests compare).  The reference generates an LR(1)
// parser with LALRPOP; this is a hand-written recursive-descent parser for the same language.  Meta.start / Meta.end are
// what LALRPOP's `@L` yields at the same places: the byte offset of the next unconsumed token (the end of the last
// token at end of input).
#include "ast.hpp"

namespace zafe {

// parse.rs:6-75.  Comments become spaces so that offsets keep pointing into the original text; the state machine
// (including what it does with "**/") is the reference's.
std::string preprocess(const std::string& text) {
    std::string expr;
    expr.reserve(text.size());
    for (size_t i = 0; i < text.size();) {          // "/*#[" -> "  #[", then "]#*/" -> "]   "
        if (text.compare(i, 4, "/*#[") == 0) { expr += "  #["; i += 4; }
        else expr += text[i++];
    }
    {
        std::string t;
        for (size_t i = 0; i < expr.size();) {
            if (expr.compare(i, 4, "]#*/") == 0) { t += "]   "; i += 4; }
            else t += expr[i++];
        }
        expr.swap(t);
    }
    // iterate over characters (UTF-8 code points), as the reference iterates over `chars()`
    std::vector<std::string> chars;
    for (size_t i = 0; i < expr.size();) {
        const unsigned char c = (unsigned char)expr[i];
        size_t len = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1;
        if (i + len > expr.size()) len = expr.size() - i;
        chars.push_back(expr.substr(i, len));
        i += len;
    }
    std::string pp;
    int state = 0;
    uint64_t loc = 0, block_comment_start = 0;
    for (size_t i = 0; i < chars.size(); i++) {
        const std::string& c0 = chars[i];
        loc += 1;
        if (state == 0 && c0 == "/") {
            loc += 1;
            if (i + 1 < chars.size()) {
                const std::string& c1 = chars[++i];
                if (c1 == "/") { state = 1; pp += "  "; }
                else if (c1 == "*") { block_comment_start = loc; state = 2; pp += "  "; }
                else { pp += c0; pp += c1; }
            } else { pp += c0; break; }
        } else if (state == 0) {
            pp += c0;
        } else if (state == 1 && c0 == "\n") {
            pp += c0;
            state = 0;
        } else if (state == 2 && c0 == "*") {
            loc += 1;
            if (i + 1 < chars.size()) {
                const std::string& c1 = chars[++i];
                pp += "  ";
                if (c1 == "/") state = 0;
            } else {
                FeError e("ParseError", "unterminated /* */");
                e.has_meta = true; e.meta_start = e.meta_end = block_comment_start;
                throw e;
            }
        } else {
            pp += ' ';
        }
    }
    return pp;
}

namespace {

enum TokKind { T_LIT, T_IDENT, T_DEC, T_HEX, T_STRING, T_EOF };
struct Token { TokKind kind; std::string text; uint64_t start, end; };

const char* const KEYWORDS[] = {"include", "function", "template", "if", "else", "for", "while", "return", "var", "component", "signal", "input", "private", "output"};
// every punctuation literal of lang.lalrpop, longest first
const char* const PUNCT[] = {"<<=", ">>=", "===", "<==", "==>", "<--", "-->", "#[", "**", "<<", ">>", "<=", ">=", "==", "!=", "&&", "||", "+=", "-=", "*=", "/=", "%=",
                             "&=", "|=", "^=", "(", ")", "{", "}", "[", "]", ";", ",", ".", "=", "!", "+", "-", "*", "/", "\\", "%", "<", ">", "|", "^", "&"};

bool is_alpha(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
bool is_digit(char c) { return c >= '0' && c <= '9'; }
bool is_hex(char c) { return is_digit(c) || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F'); }
bool is_ident_tail(char c) { return is_alpha(c) || is_digit(c) || c == '$' || c == '_'; }

[[noreturn]] void parse_error(const std::string& what, uint64_t l, uint64_t r) {
    FeError e("ParseError", what);
    e.has_meta = true; e.meta_start = l; e.meta_end = r;
    throw e;
}

std::vector<Token> lex(const std::string& s) {
    std::vector<Token> out;
    size_t i = 0;
    const size_t n = s.size();
    while (true) {
        while (i < n && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r' || s[i] == '\f' || s[i] == '\v')) i++;
        if (i >= n) break;
        const char c = s[i];
        size_t j = i;
        if (is_alpha(c)) {                                              // IDENTIFIER: [a-zA-Z][a-zA-Z$_0-9]*, keywords win ties
            while (j < n && is_ident_tail(s[j])) j++;
            std::string t = s.substr(i, j - i);
            bool kw = false;
            for (const char* k : KEYWORDS) if (t == k) kw = true;
            out.push_back({kw ? T_LIT : T_IDENT, t, i, j});
        } else if (is_digit(c)) {
            if (c == '0' && i + 1 < n && s[i + 1] == 'x') {              // HEXNUMBER: 0x[0-9A-Fa-f]* (longest match against DECNUMBER "0")
                j = i + 2;
                while (j < n && is_hex(s[j])) j++;
                out.push_back({T_HEX, s.substr(i + 2, j - i - 2), i, j});
            } else {
                while (j < n && is_digit(s[j])) j++;
                out.push_back({T_DEC, s.substr(i, j - i), i, j});
            }
        } else if (c == '"') {                                         // STRING: "[^"]*"
            j = i + 1;
            while (j < n && s[j] != '"') j++;
            if (j >= n) parse_error("InvalidToken { location: " + std::to_string(i) + " }", i, i);
            out.push_back({T_STRING, s.substr(i + 1, j - i - 1), i, j + 1});
            j++;
        } else {
            const char* hit = nullptr;
            for (const char* p : PUNCT) if (s.compare(i, strlen(p), p) == 0) { hit = p; break; }
            if (!hit) parse_error("InvalidToken { location: " + std::to_string(i) + " }", i, i);
            j = i + strlen(hit);
            out.push_back({T_LIT, hit, i, j});
        }
        i = j;
    }
    const uint64_t endpos = out.empty() ? 0 : out.back().end;
    out.push_back({T_EOF, "", endpos, endpos});
    return out;
}

template <class T>
std::shared_ptr<const T> share(T& v) { return std::make_shared<const T>(std::move(v)); }

struct Parser {
    std::vector<Token> toks;
    size_t p = 0;
    explicit Parser(const std::string& text) : toks(lex(text)) {}

    const Token& cur() const { return toks[p]; }
    const Token& peek(size_t k) const { return toks[std::min(p + k, toks.size() - 1)]; }
    uint64_t loc() const { return toks[p].start; }                      // @L
    bool is(const char* lit) const { return cur().kind == T_LIT && cur().text == lit; }
    bool at_eof() const { return cur().kind == T_EOF; }
    [[noreturn]] void unexpected(const char* expected) const {
        const Token& t = cur();
        if (t.kind == T_EOF) parse_error(std::string("UnrecognizedEOF { location: ") + std::to_string(t.start) + ", expected: [" + expected + "] }", 0, 0);
        parse_error("UnrecognizedToken { token: (" + std::to_string(t.start) + ", \"" + t.text + "\", " + std::to_string(t.end) + "), expected: [" + expected + "] }",
                    t.start, t.end);
    }
    void expect(const char* lit) { if (!is(lit)) unexpected(lit); p++; }
    bool accept(const char* lit) { if (is(lit)) { p++; return true; } return false; }
    std::string ident() { if (cur().kind != T_IDENT) unexpected("IDENTIFIER"); return toks[p++].text; }
    Meta meta(uint64_t s, uint64_t e) const { Meta m; m.start = s; m.end = e; return m; }

    // ---- expressions (lang.lalrpop:327-417) ------------------------------------------------------------------
    static int infix_level(const Token& t, Opcode& op) {
        if (t.kind != T_LIT) return 0;
        static const struct { const char* s; Opcode op; int lvl; } T[] = {
            {"||", Opcode::BoolOr, 12}, {"&&", Opcode::BoolAnd, 11}, {"==", Opcode::Eq, 10}, {"!=", Opcode::NotEq, 10}, {"<", Opcode::Lesser, 10},
            {">", Opcode::Greater, 10}, {"<=", Opcode::LesserEq, 10}, {">=", Opcode::GreaterEq, 10}, {"|", Opcode::BitOr, 9}, {"^", Opcode::BitXor, 8},
            {"&", Opcode::BitAnd, 7}, {"<<", Opcode::ShiftL, 6}, {">>", Opcode::ShiftR, 6}, {"+", Opcode::Add, 5}, {"-", Opcode::Sub, 5}, {"*", Opcode::Mul, 4},
            {"/", Opcode::Div, 4}, {"\\", Opcode::IntDiv, 4}, {"%", Opcode::Mod, 4}, {"**", Opcode::Pow, 3}};
        for (auto& e : T) if (t.text == e.s) { op = e.op; return e.lvl; }
        return 0;
    }
    ExprP expression() { return tier(12); }
    ExprP tier(int level) {
        if (level == 2) return prefix();
        const uint64_t s = loc();
        ExprP lhs = tier(level - 1);
        Opcode op;
        while (infix_level(cur(), op) == level) {
            p++;
            ExprP rhs = tier(level - 1);
            Expression e;
            e.kind = ExprKind::InfixOp; e.meta = meta(s, loc()); e.lhe = lhs; e.op = op; e.rhe = rhs;
            lhs = share(e);
        }
        return lhs;
    }
    ExprP prefix() {                                                   // Expression2: unary - and ! over Expression1
        if (is("-") || is("!")) {
            const uint64_t s = loc();
            const Opcode op = is("-") ? Opcode::Sub : Opcode::BoolNot;
            p++;
            ExprP rhe = expression1();
            Expression e;
            e.kind = ExprKind::PrefixOp; e.meta = meta(s, loc()); e.op = op; e.rhe = rhe;
            return share(e);
        }
        return expression1();
    }
    std::vector<ExprP> expression_list(const char* close) {           // (Expression ",")* Expression?
        std::vector<ExprP> v;
        while (!is(close)) {
            v.push_back(expression());
            if (!accept(",")) break;
        }
        return v;
    }
    ExprP expression1() {
        const uint64_t s = loc();
        if (cur().kind == T_IDENT && peek(1).kind == T_LIT && peek(1).text == "(") {
            Expression e;
            e.kind = ExprKind::FunctionCall;
            e.name = ident();
            expect("(");
            e.list = expression_list(")");
            expect(")");
            e.meta = meta(s, loc());
            return share(e);
        }
        if (is("[")) {
            p++;
            Expression e;
            e.kind = ExprKind::Array;
            e.list = expression_list("]");
            expect("]");
            e.meta = meta(s, loc());
            return share(e);
        }
        return expression0();
    }
    ExprP expression0() {
        const uint64_t s = loc();
        if (cur().kind == T_IDENT) {
            Expression e;
            e.kind = ExprKind::Variable;
            e.var = variable();
            e.meta = meta(s, loc());
            return share(e);
        }
        if (cur().kind == T_DEC || cur().kind == T_HEX) {
            Expression e;
            e.kind = ExprKind::Number;
            if (cur().kind == T_HEX && cur().text.empty()) parse_error("failed to parse base16", cur().start, cur().end);
            e.number = BigDigits::parse(cur().text, cur().kind == T_HEX ? 16 : 10);
            p++;
            e.meta = meta(s, loc());
            return share(e);
        }
        if (accept("(")) {
            ExprP inner = expression();
            expect(")");
            return inner;
        }
        unexpected("expression");
    }
    // ---- variables (lang.lalrpop:262-322) ------------------------------------------------------------------------
    Selector index_selector() {
        Selector sel;
        const uint64_t s = loc();
        expect("[");
        sel.pos = expression();
        expect("]");
        sel.meta = meta(s, loc());
        return sel;
    }
    VarP variable() {                                                  // IDENTIFIER PinOrIndexSelector*
        Variable v;
        const uint64_t s = loc();
        v.name = ident();
        while (true) {
            if (is(".")) {
                Selector sel;
                sel.is_pin = true;
                const uint64_t ss = loc();
                p++;
                sel.name = ident();
                sel.meta = meta(ss, loc());
                v.sels.push_back(std::move(sel));
            } else if (is("[")) {
                v.sels.push_back(index_selector());
            } else break;
        }
        v.meta = meta(s, loc());
        return share(v);
    }
    VarP decl_variable(bool allow_index) {                             // IndexVariableDecl / SimpleVariableDecl
        Variable v;
        const uint64_t s = loc();
        v.name = ident();
        while (allow_index && is("[")) v.sels.push_back(index_selector());
        v.meta = meta(s, loc());
        return share(v);
    }
    // ---- declarations and substitutions (lang.lalrpop:183-256) ---------------------------------------------------
    bool at_declaration() const { return is("var") || is("component") || is("signal"); }
    Statement declaration() {
        Statement st;
        st.kind = StmtKind::Declaration;
        const uint64_t s = loc();
        if (is("var") || is("component")) {
            st.xtype.kind = is("var") ? VarKind::Var : VarKind::Component;
            p++;
            // `var a[2]` (indexed, no initialiser) or `var a = expr` (plain name with initialiser)
            if (cur().kind == T_IDENT && peek(1).kind == T_LIT && peek(1).text == "=") {
                st.name = decl_variable(false);
                expect("=");
                st.has_init = true;
                st.op = Opcode::Assig;
                st.value = expression();
            } else {
                st.name = decl_variable(true);
            }
        } else {
            expect("signal");
            st.xtype.kind = VarKind::Signal;
            if (accept("private")) { expect("input"); st.xtype.signal = SignalType::PrivateInput; }
            else if (accept("input")) st.xtype.signal = SignalType::PublicInput;
            else if (accept("output")) st.xtype.signal = SignalType::Output;
            else st.xtype.signal = SignalType::Internal;
            st.name = decl_variable(true);
        }
        st.meta = meta(s, loc());
        return st;
    }
    static bool assign_op(const Token& t, Opcode& op) {
        if (t.kind != T_LIT) return false;
        static const struct { const char* s; Opcode op; } T[] = {{"=", Opcode::Assig}, {"+=", Opcode::AssigAdd}, {"-=", Opcode::AssigSub}, {"*=", Opcode::AssigMul},
            {"/=", Opcode::AssigDiv}, {"%=", Opcode::AssigMod}, {"<<=", Opcode::AssigShiftL}, {">>=", Opcode::AssigShiftR}, {"&=", Opcode::AssigBitAnd},
            {"|=", Opcode::AssigBitOr}, {"^=", Opcode::AssigBitXor}};
        for (auto& e : T) if (t.text == e.s) { op = e.op; return true; }
        return false;
    }
    Statement substitution() {                                         // Variable OpAssigClass Expression
        Statement st;
        st.kind = StmtKind::Substitution;
        const uint64_t s = loc();
        st.name = variable();
        if (!assign_op(cur(), st.op)) unexpected("assignment operator");
        p++;
        st.value = expression();
        st.meta = meta(s, loc());
        return st;
    }
    // ---- statements (lang.lalrpop:70-181) ----------------------------------------------------------------------------
    std::vector<std::string> attrs_opt() {
        std::vector<std::string> a;
        if (!accept("#[")) return a;
        while (!is("]")) {
            a.push_back(ident());
            if (!accept(",")) break;
        }
        expect("]");
        return a;
    }
    StmtP block() {
        Statement st;
        st.kind = StmtKind::Block;
        const uint64_t s = loc();
        expect("{");
        while (!is("}")) {
            if (at_eof()) unexpected("}");
            st.stmts.push_back(statement());
        }
        expect("}");
        st.meta = meta(s, loc());
        return share(st);
    }
    Statement if_then_else() {                                         // after the "if" keyword
        Statement st;
        st.kind = StmtKind::IfThenElse;
        const uint64_t s = loc();
        expect("(");
        st.cond = expression();
        expect(")");
        st.xthen = block();
        if (accept("else")) {
            if (accept("if")) { Statement inner = if_then_else(); st.xelse = share(inner); }
            else st.xelse = block();
        }
        st.meta = meta(s, loc());
        return st;
    }
    StmtP statement() {
        const uint64_t s = loc();
        std::vector<std::string> attrs = attrs_opt();
        if (accept("if")) {
            Statement st = if_then_else();
            st.meta.attrs = attrs;
            return share(st);
        }
        if (accept("for")) {
            Statement st;
            st.kind = StmtKind::For;
            expect("(");
            { Statement init = at_declaration() ? declaration() : substitution(); st.init = share(init); }
            expect(";");
            st.cond = expression();
            expect(";");
            { Statement step = substitution(); st.step = share(step); }
            expect(")");
            st.body = block();
            st.meta = meta(s, loc());
            st.meta.attrs = attrs;
            return share(st);
        }
        if (accept("while")) {
            Statement st;
            st.kind = StmtKind::While;
            expect("(");
            st.cond = expression();
            expect(")");
            st.body = block();
            st.meta = meta(s, loc());
            st.meta.attrs = attrs;
            return share(st);
        }
        if (accept("return")) {
            Statement st;
            st.kind = StmtKind::Return;
            st.value = expression();
            expect(";");
            st.meta = meta(s, loc());
            st.meta.attrs = attrs;
            return share(st);
        }
        if (at_declaration()) {
            Statement st = declaration();
            expect(";");
            st.meta.attrs = attrs;
            return share(st);
        }
        if (is("{")) {
            Statement st = *block();
            st.meta.attrs = attrs;
            return share(st);
        }
        if (cur().kind == T_IDENT && peek(1).kind == T_LIT && peek(1).text == "!" && peek(2).kind == T_LIT && peek(2).text == "(") {
            Statement st;
            st.kind = StmtKind::InternalCall;
            st.call_name = ident();
            expect("!");
            expect("(");
            st.args = expression_list(")");
            expect(")");
            expect(";");
            st.meta = meta(s, loc());
            st.meta.attrs = attrs;
            return share(st);
        }
        // Substitution | SignalLeft | SignalRight | SignalEq: all start with an expression (a Variable in the first two)
        const uint64_t es = loc();
        ExprP first = expression();
        Opcode op;
        if (assign_op(cur(), op)) {
            if (first->kind != ExprKind::Variable) unexpected("=== ==> -->");
            Statement st;
            st.kind = StmtKind::Substitution;
            st.name = first->var;
            st.op = op;
            p++;
            st.value = expression();
            st.meta = meta(es, loc());                                 // the Substitution's own span (lang.lalrpop:250-256)
            expect(";");
            st.meta.attrs = attrs;
            return share(st);
        }
        if (is("<--") || is("<==")) {
            if (first->kind != ExprKind::Variable) unexpected("=== ==> -->");
            Statement st;
            st.kind = StmtKind::SignalLeft;
            st.name = first->var;
            st.op = is("<--") ? Opcode::SignalWireLeft : Opcode::SignalContrainLeft;
            p++;
            st.value = expression();
            expect(";");
            st.meta = meta(s, loc());
            st.meta.attrs = attrs;
            return share(st);
        }
        if (is("-->") || is("==>")) {
            Statement st;
            st.kind = StmtKind::SignalRight;
            st.value = first;
            st.op = is("-->") ? Opcode::SignalWireRight : Opcode::SignalContrainRight;
            p++;
            st.name = variable();
            expect(";");
            st.meta = meta(s, loc());
            st.meta.attrs = attrs;
            return share(st);
        }
        if (is("===")) {
            Statement st;
            st.kind = StmtKind::SignalEq;
            st.lhe = first;
            st.op = Opcode::SignalContrainEq;
            p++;
            st.value = expression();
            expect(";");
            st.meta = meta(s, loc());
            st.meta.attrs = attrs;
            return share(st);
        }
        unexpected("= <-- <== --> ==> ===");
    }
    // ---- body (lang.lalrpop:13-66) -----------------------------------------------------------------------------
    std::vector<std::string> parameter_list() {
        std::vector<std::string> v;
        while (!is(")")) {
            v.push_back(ident());
            if (!accept(",")) break;
        }
        return v;
    }
    BodyElement body_element() {
        BodyElement b;
        const uint64_t s = loc();
        if (accept("include")) {
            b.kind = BodyKind::Include;
            if (cur().kind != T_STRING) unexpected("STRING");
            b.path = toks[p++].text;
            expect(";");
            b.meta = meta(s, loc());
            return b;
        }
        std::vector<std::string> attrs = attrs_opt();
        if (is("function") || is("template")) {
            b.kind = is("function") ? BodyKind::FunctionDef : BodyKind::TemplateDef;
            p++;
            b.name = ident();
            expect("(");
            b.args = parameter_list();
            expect(")");
            b.stmt = block();
        } else if (at_declaration()) {
            b.kind = BodyKind::Declaration;
            Statement d = declaration();
            b.stmt = share(d);
            expect(";");
        } else {
            unexpected("include function template var component signal");
        }
        b.meta = meta(s, loc());
        b.meta.attrs = attrs;
        return b;
    }
    void finish() { if (!at_eof()) unexpected("end of input"); }
};

}  // namespace

std::vector<BodyElement> parse_body(const std::string& text) {         // parse.rs:78-97
    Parser ps(preprocess(text));
    std::vector<BodyElement> out;
    while (!ps.at_eof()) out.push_back(ps.body_element());
    return out;
}
StmtP parse_statement(const std::string& text) { Parser ps(text); StmtP s = ps.statement(); ps.finish(); return s; }
ExprP parse_expression(const std::string& text) { Parser ps(text); ExprP e = ps.expression(); ps.finish(); return e; }
BodyElement parse_body_element(const std::string& text) { Parser ps(text); BodyElement b = ps.body_element(); ps.finish(); return b; }

// ---- display.rs ----------------------------------------------------------------------------------------------------
static std::string join_exprs(const std::vector<ExprP>& l) {
    std::string s;
    for (size_t i = 0; i < l.size(); i++) { if (i) s += ","; s += debug_string(*l[i]); }
    return s;
}
std::string debug_string(const Variable& v) {
    std::string s = v.name;
    for (auto& sel : v.sels) s += sel.is_pin ? "." + sel.name : "[" + debug_string(*sel.pos) + "]";
    return s;
}
std::string debug_string(const Expression& e) {
    switch (e.kind) {
        case ExprKind::Variable: return debug_string(*e.var);
        case ExprKind::Number: return e.number.to_decimal();
        case ExprKind::PrefixOp: return std::string("(") + opcode_text(e.op) + " " + debug_string(*e.rhe) + ")";
        case ExprKind::InfixOp: return "(" + debug_string(*e.lhe) + " " + opcode_text(e.op) + " " + debug_string(*e.rhe) + ")";
        case ExprKind::Array: return "[" + join_exprs(e.list) + "]";
        case ExprKind::FunctionCall: return e.name + "(" + join_exprs(e.list) + ")";
    }
    return "";
}
static std::string type_text(const VariableType& t) {
    switch (t.kind) {
        case VarKind::Empty: return "";
        case VarKind::Var: return "var";
        case VarKind::Component: return "component";
        case VarKind::Signal:
            return t.signal == SignalType::Internal ? "signal" : t.signal == SignalType::PublicInput ? "signal input"
                   : t.signal == SignalType::PrivateInput ? "signal private input" : "signal output";
    }
    return "";
}
static std::string for_item(const Statement& s) {
    if (s.kind == StmtKind::Declaration)
        return s.has_init ? type_text(s.xtype) + " " + debug_string(*s.name) + " " + opcode_text(s.op) + " " + debug_string(*s.value)
                          : type_text(s.xtype) + " " + debug_string(*s.name);
    return debug_string(*s.name) + " " + opcode_text(s.op) + " " + debug_string(*s.value);
}
std::string debug_string(const Statement& s) {
    switch (s.kind) {
        case StmtKind::Block: {
            std::string t = "{";
            for (size_t i = 0; i < s.stmts.size(); i++) { if (i) t += " "; t += debug_string(*s.stmts[i]); }
            return t + "}";
        }
        case StmtKind::IfThenElse:
            return "if (" + debug_string(*s.cond) + ") " + debug_string(*s.xthen) + (s.xelse ? " else " + debug_string(*s.xelse) : "");
        case StmtKind::For:
            return "for (" + for_item(*s.init) + ";" + debug_string(*s.cond) + ";" + for_item(*s.step) + ") " + debug_string(*s.body);
        case StmtKind::While: return "while (" + debug_string(*s.cond) + ") " + debug_string(*s.body);
        case StmtKind::Return: return "return " + debug_string(*s.value) + ";";
        case StmtKind::Declaration: return for_item(s) + ";";
        case StmtKind::Substitution: return for_item(s) + ";";
        case StmtKind::SignalLeft: return debug_string(*s.name) + " " + opcode_text(s.op) + " " + debug_string(*s.value) + ";";
        case StmtKind::SignalRight: return debug_string(*s.value) + " " + opcode_text(s.op) + " " + debug_string(*s.name) + ";";
        case StmtKind::SignalEq: return debug_string(*s.lhe) + " " + opcode_text(s.op) + " " + debug_string(*s.value) + ";";
        case StmtKind::InternalCall: return s.call_name + "!(" + join_exprs(s.args) + ");";
    }
    return "";
}
std::string debug_string(const BodyElement& b) {
    auto join = [](const std::vector<std::string>& a) { std::string s; for (size_t i = 0; i < a.size(); i++) { if (i) s += ","; s += a[i]; } return s; };
    switch (b.kind) {
        case BodyKind::Include: return "include \"" + b.path + "\";";
        case BodyKind::FunctionDef: return "function " + b.name + "(" + join(b.args) + ") " + debug_string(*b.stmt);
        case BodyKind::TemplateDef: return "template " + b.name + "(" + join(b.args) + ") " + debug_string(*b.stmt);
        case BodyKind::Declaration: return debug_string(*b.stmt);
    }
    return "";
}

}  // namespace zafe
