// Evaluator of za's circuit language: one tree walk that either generates the constraint system (Mode::GenConstraints)
// or computes the witness (Mode::GenWitness).  Behavioural restatement of
// /root/reference/compiler/src/evaluator/eval.rs (cited per function), scope.rs and types.rs: signal naming and numbering,
// the order constraints are emitted in and the order of terms inside them are observable (proving-key file, optimiser,
// the bellman variable numbering of prover.rs:45-103) and follow the reference exactly.
#include <fstream>
#include <sstream>
#include "frontend.hpp"

namespace zafe {

// ---- small types -------------------------------------------------------------------------------------------------------
std::string rust_debug_str(const std::string& s) {
    std::string o = "\"";
    for (char c : s) {
        if (c == '"') o += "\\\"";
        else if (c == '\\') o += "\\\\";
        else if (c == '\n') o += "\\n";
        else if (c == '\t') o += "\\t";
        else if (c == '\r') o += "\\r";
        else o += c;
    }
    return o + "\"";
}

std::string error_debug(const FeError& e) {
    static const char* ALG[] = {"InvalidOperation", "InvalidFormat"};
    static const char* TOP[] = {"BadFormat", "Unexpected", "Json", "Bincode", "Synthesis", "IoError"};
    for (const char* k : TOP)
        if (e.kind == k) {
            if (e.kind == "IoError") return "Io(" + e.text + ")";
            if (e.kind == "Json" || e.kind == "Bincode" || e.kind == "Synthesis") return e.kind + "(" + e.text + ")";
            return e.kind + "(" + rust_debug_str(e.text) + ")";
        }
    for (const char* k : ALG)
        if (e.kind == k) return "Evaluator(Algebra(" + e.kind + "(" + rust_debug_str(e.text) + ")))";
    if (e.kind == "ParseError") return "Evaluator(Parse(" + rust_debug_str(e.text) + "))";
    if (e.kind == "CannotConvertToU64") return "Evaluator(CannotConvertToU64(" + e.text + "))";
    if (e.kind == "Io") return "Evaluator(Io(" + e.text + "))";
    return "Evaluator(" + e.kind + "(" + rust_debug_str(e.text) + "))";
}

static const char* signal_type_name(SignalType t) {
    return t == SignalType::Output ? "Output" : t == SignalType::PublicInput ? "PublicInput" : t == SignalType::PrivateInput ? "PrivateInput" : "Internal";
}
std::string Signals::to_string(SignalId id) const {
    const Signal& s = ids[id];
    return s.full_name + ":" + signal_type_name(s.xtype) + ":" + (s.has_value ? "Some(" + s.value.to_string() + ")" : "None");
}

std::string Constraints::satisfies_with_signals(const Signals& s) const {
    std::string err;
    auto eval_lc = [&](const LC& lc, FS& out) {
        FS acc = FS::zero();
        for (auto& t : lc.t) {
            FS v;
            if (t.first == 0) v = FS::one();
            else {
                const Signal& sg = s.ids.at(t.first);
                if (!sg.has_value || sg.value.kind != Value::FieldScalar) { err = "signal bad value " + s.to_string(t.first); return false; }
                v = sg.value.fs;
            }
            acc = acc.add(t.second.mul(v));
        }
        out = acc;
        return true;
    };
    for (size_t n = 0; n < q.size(); n++) {
        FS a, b, c;
        if (!eval_lc(q[n].a, a) || !eval_lc(q[n].b, b) || !eval_lc(q[n].c, c)) return err;
        const FS zero = a.mul(b).add(c);
        if (!zero.is_zero())
            return "constrain '" + s.format(Value::of(q[n])) + "' (" + debug[n] + ") evals to non-zero (" + zero.format(false) + ")";
    }
    return "";
}

List List::make(const std::vector<size_t>& sizes, size_t at) {
    List l;
    if (at >= sizes.size()) return l;               // List::Algebra(Value::default()) = 0
    l.is_value = false;
    for (size_t i = 0; i < sizes[at]; i++) l.items.push_back(make(sizes, at + 1));
    return l;
}
const List& List::get(const std::vector<size_t>& idx, size_t at) const {
    if (at >= idx.size()) return *this;
    if (is_value) fail("InvalidSelector", "index at [" + std::to_string(idx[at]) + "] contains a value");
    if (idx[at] >= items.size()) fail("InvalidSelector", "index at [" + std::to_string(idx[at]) + "] too large");
    return items[idx[at]].get(idx, at + 1);
}
void List::set(const Value& value, const std::vector<size_t>& idx, size_t at) {
    if (is_value) fail("InvalidSelector", "index at [" + std::to_string(at < idx.size() ? idx[at] : 0) + "] contains a value");
    if (at >= idx.size() || idx[at] >= items.size()) fail("InvalidSelector", "invalid index for " + debug());
    if (at + 1 == idx.size()) { List v; v.v = value; items[idx[at]] = v; return; }
    items[idx[at]].set(value, idx, at + 1);
}
std::string List::debug() const {
    if (is_value) return "Algebra(" + v.to_string() + ")";
    std::string s = "List([";
    for (size_t i = 0; i < items.size(); i++) { if (i) s += ", "; s += items[i].debug(); }
    return s + "])";
}

std::string ReturnValue::debug() const {
    return kind == Bool ? std::string("Bool(") + (b ? "true" : "false") + ")" : kind == Algebra ? "Algebra(" + a.to_string() + ")" : "List(" + l.debug() + ")";
}
const Value& ReturnValue::into_algebra() const {
    if (kind != Algebra) fail("InvalidType", "Cannot convert to algebraic value " + debug());
    return a;
}
bool ReturnValue::into_bool() const {
    if (kind != Bool) fail("InvalidType", "Cannot convert to boolean value " + debug());
    return b;
}
const FS& ReturnValue::into_fs() const {
    if (kind != Algebra || a.kind != Value::FieldScalar) fail("InvalidType", "Cannot convert to scalar value " + debug());
    return a.fs;
}
uint64_t ReturnValue::into_u64() const {
    const FS& f = into_fs();
    uint64_t v;
    if (!f.try_to_u64(v)) fail("CannotConvertToU64", f.format(false));
    return v;
}

ScopeValue ScopeValue::from(const ReturnValue& r) {
    ScopeValue s;
    if (r.kind == ReturnValue::Bool) { s.kind = Bool; s.b = r.b; }
    else if (r.kind == ReturnValue::Algebra) { s.kind = Algebra; s.a = r.a; }
    else { s.kind = ListK; s.l = r.l; }
    return s;
}
std::string ScopeValue::debug() const {
    switch (kind) {
        case UndefVar: return "UndefVar";
        case UndefComponent: return "UndefComponent";
        case Bool: return std::string("Bool(") + (b ? "true" : "false") + ")";
        case Algebra: return "Algebra(" + a.to_string() + ")";
        case Function: return "Function";
        case Template: return "Template";
        case Component: return "Component { template: " + rust_debug_str(tmpl) + " }";
        case ListK: return "List(" + l.debug() + ")";
    }
    return "";
}

void Scope::insert(const std::string& k, const ScopeValue& v) {
    if (vars.count(k)) fail("AlreadyExists", k);
    vars.emplace(k, v);
}
ScopeValue* Scope::get(const std::string& k) {
    Scope* it = this;
    while (true) {
        auto f = it->vars.find(k);
        if (f != it->vars.end()) return &f->second;
        if (!it->prev || it->start) return nullptr;
        it = it->prev;
    }
}
void Scope::update(const std::string& k, const ScopeValue& v) {
    ScopeValue* p = get(k);
    if (!p) fail("NotFound", k);
    *p = v;
}

// ---- evaluator ---------------------------------------------------------------------------------------------------------
namespace {
struct Swap {                                       // std::mem::swap on entry, swap back on normal exit (eval.rs:350-358)
    std::string& target;
    std::string saved;
    Swap(std::string& t, const std::string& v) : target(t), saved(t) { target = v; }
    void restore() { target = saved; }
};
std::string parent_dir(const std::string& p) {
    const size_t k = p.find_last_of('/');
    if (k == std::string::npos) return "";
    if (k == 0) return "/";
    return p.substr(0, k);
}
std::string join_path(const std::string& base, const std::string& f) {       // PathBuf::push
    if (!f.empty() && f[0] == '/') return f;
    if (base.empty()) return f;
    return base.back() == '/' ? base + f : base + "/" + f;
}
}  // namespace

void Evaluator::note_error(const Meta& meta) {      // register_error, eval.rs:167-178: the FIRST error keeps its context
    if (last_error.set) return;
    last_error.set = true;
    last_error.file = current_file;
    last_error.component = current_component;
    last_error.function = in_function ? current_function : "";
    last_error.start = meta.start;
    last_error.end = meta.end;
}

void Evaluator::eval_inline(Scope& scope, const std::string& code) {        // eval.rs:99-110
    std::vector<BodyElement> elements;
    try { elements = parse_body(code); }
    catch (FeError& e) { Meta m; m.start = e.meta_start; m.end = e.meta_end; note_error(m); throw; }
    eval_body_elements(scope, elements);
    for (auto& b : elements) collected_asts.push_back(b);
}

void Evaluator::eval_template(Scope& scope, const std::string& template_name) {     // eval.rs:112-122
    ScopeValue* t = scope.get(template_name);
    if (!t || t->kind != ScopeValue::Template) fail("NotFound", "template " + template_name);
    StmtP stmt = t->stmt;
    Scope inner(true, &scope, t->path);
    eval_statement(inner, *stmt);
}

void Evaluator::eval_file(Scope& scope, const std::string& p, const std::string& filename) {      // eval.rs:124-129
    path = p;
    eval_include(scope, filename);
}

void Evaluator::eval_asts(Scope& scope, const std::vector<BodyElement>& asts) {      // eval.rs:131-160
    for (auto& b : asts)
        if (b.kind == BodyKind::FunctionDef || b.kind == BodyKind::TemplateDef) eval_body_element(scope, b);
    for (auto& b : asts)
        if (b.kind == BodyKind::Declaration) eval_statement(scope, *b.stmt);
}

ReturnValue Evaluator::eval_expression(Scope& scope, const Expression& e) {          // eval.rs:206-216
    try {
        switch (e.kind) {
            case ExprKind::FunctionCall: return eval_function_call(e.meta, scope, e.name, e.list);
            case ExprKind::Variable: return eval_variable(scope, *e.var);
            case ExprKind::Number: return ReturnValue::of(Value::of(FS(e.number.mod_field())));      // eval.rs:596-600
            case ExprKind::PrefixOp: {                                                                  // eval.rs:602-618
                ReturnValue r = eval_expression(scope, *e.rhe);
                return ReturnValue::of(eval_prefix(e.op, r.into_algebra()));
            }
            case ExprKind::InfixOp: return eval_infix_op(scope, e);
            case ExprKind::Array: {                                                                      // eval.rs:684-707
                List out;
                out.is_value = false;
                for (auto& x : e.list) {
                    ReturnValue v = eval_expression(scope, *x);
                    if (v.kind == ReturnValue::Algebra) { List l; l.v = v.a; out.items.push_back(l); }
                    else if (v.kind == ReturnValue::ListK) out.items.push_back(v.l);
                    else fail("InvalidType", "a boolean cannot be an array element");      // the reference hits unreachable!()
                }
                return ReturnValue::of(out);
            }
        }
    } catch (FeError&) { note_error(e.meta); throw; }
    fail("Unexpected", "expression");
}

ReturnValue Evaluator::eval_infix_op(Scope& scope, const Expression& e) {            // eval.rs:620-682
    ReturnValue left = eval_expression(scope, *e.lhe);
    ReturnValue right = eval_expression(scope, *e.rhe);
    switch (e.op) {
        case Opcode::Add: case Opcode::Sub: case Opcode::Mul: case Opcode::Div: case Opcode::IntDiv: case Opcode::Mod: case Opcode::ShiftL:
        case Opcode::ShiftR: case Opcode::BitAnd: case Opcode::BitOr: case Opcode::BitXor: case Opcode::Pow: {
            const Value& l = left.into_algebra();
            const Value& r = right.into_algebra();
            return ReturnValue::of(eval_infix(l, e.op, r));
        }
        case Opcode::BoolAnd: { const bool l = left.into_bool(); if (!l) return ReturnValue::of(false); return ReturnValue::of(right.into_bool()); }
        case Opcode::BoolOr: { const bool l = left.into_bool(); if (l) return ReturnValue::of(true); return ReturnValue::of(right.into_bool()); }
        case Opcode::Greater: { const FS& l = left.into_fs(); return ReturnValue::of(cmp(l.n, right.into_fs().n) > 0); }
        case Opcode::GreaterEq: { const FS& l = left.into_fs(); return ReturnValue::of(cmp(l.n, right.into_fs().n) >= 0); }
        case Opcode::Lesser: { const FS& l = left.into_fs(); return ReturnValue::of(cmp(l.n, right.into_fs().n) < 0); }
        case Opcode::LesserEq: { const FS& l = left.into_fs(); return ReturnValue::of(cmp(l.n, right.into_fs().n) <= 0); }
        case Opcode::Eq:
        case Opcode::NotEq: {
            const bool want = e.op == Opcode::Eq;
            if (left.kind == ReturnValue::Bool && right.kind == ReturnValue::Bool) return ReturnValue::of((left.b == right.b) == want);
            if (left.kind == ReturnValue::Algebra && right.kind == ReturnValue::Algebra && left.a.kind == Value::FieldScalar &&
                right.a.kind == Value::FieldScalar)
                return ReturnValue::of((left.a.fs == right.a.fs) == want);
            fail("InvalidType", "Cannot compare " + left.debug() + "==" + right.debug());
        }
        default: fail("NotYetImplemented", std::string("eval_infix_op '") + opcode_text(e.op) + "'");
    }
}

ReturnValue Evaluator::eval_function_call(const Meta& meta, Scope& scope, const std::string& name, const std::vector<ExprP>& params) {   // eval.rs:314-365
    ScopeValue* f = scope.root()->get(name);
    if (!f || f->kind != ScopeValue::Function) fail("NotFound", "function " + name);
    const std::vector<std::string> args = f->args;
    StmtP stmt = f->stmt;
    const std::string fpath = f->path;
    if (args.size() != params.size()) fail("InvalidParameter", name);
    Scope func_scope(true, &scope, current_file + ":" + std::to_string(meta.start));
    for (size_t n = 0; n < args.size(); n++) {
        ReturnValue v = eval_expression(scope, *params[n]);
        func_scope.insert(args[n], ScopeValue::from(v));
    }
    Swap fn(current_function, name), fl(current_file, fpath);
    const bool was_in = in_function;
    in_function = true;
    eval_statement(func_scope, *stmt);
    in_function = was_in;
    fn.restore(); fl.restore();
    ReturnValue out;
    if (!func_scope.take_return(out)) fail("BadFunctionReturn", name);
    return out;
}

void Evaluator::eval_component_decl(Scope& scope, const Variable& name) {            // eval.rs:367-372
    ScopeValue u;
    u.kind = ScopeValue::UndefComponent;
    for (auto& n : generate_selectors(scope, name)) scope.insert(n, u);
}

void Evaluator::eval_component_inst(const Meta& meta, Scope& scope, const std::string& component_name, const Expression& init) {     // eval.rs:374-494
    auto invalid_template = [&]() { fail("InvalidType", "component " + component_name + " only can be initialized with existingtemplate"); };
    if (init.kind != ExprKind::FunctionCall) invalid_template();
    const std::string& template_name = init.name;
    ScopeValue* t = scope.root()->get(template_name);
    if (!t || t->kind != ScopeValue::Template) invalid_template();
    const std::vector<std::string> args = t->args;
    StmtP stmt = t->stmt;
    const std::string tpath = t->path;
    if (args.size() != init.list.size()) fail("InvalidParameter", "Invalid parameter count when instantiating " + template_name);

    ScopeValue comp;
    comp.kind = ScopeValue::Component;
    comp.tmpl = template_name;
    comp.path = tpath;
    Scope template_scope(true, &scope, current_file + ":" + std::to_string(meta.start));
    for (size_t n = 0; n < args.size(); n++) {
        ReturnValue v = eval_expression(scope, *init.list[n]);
        comp.cargs.push_back(v);
        template_scope.insert(args[n], ScopeValue::from(v));
    }
    {
        Swap fl(current_file, tpath), cc(current_component, expand_full_name(component_name));
        if (stmt->kind != StmtKind::Block) fail("Unexpected", "template body is not a block");
        // signal declarations of the template's top level, stable-sorted by kind: outputs, public inputs, private inputs, internal
        std::vector<const Statement*> decls;
        for (auto& s : stmt->stmts)
            if (s->kind == StmtKind::Declaration && s->xtype.kind == VarKind::Signal) decls.push_back(s.get());
        std::stable_sort(decls.begin(), decls.end(), [](const Statement* a, const Statement* b) { return (uint32_t)a->xtype.signal < (uint32_t)b->xtype.signal; });
        for (const Statement* d : decls) {
            std::vector<SignalId> pending = eval_declaration_signals(template_scope, d->xtype.signal, *d->name);
            const bool is_input = d->xtype.signal == SignalType::PublicInput || d->xtype.signal == SignalType::PrivateInput;
            const bool is_not_main_in_genconstraints = !(component_name == "main" && mode == Mode::GenConstraints);
            if (is_input && is_not_main_in_genconstraints) comp.pending_inputs.insert(comp.pending_inputs.end(), pending.begin(), pending.end());
        }
        fl.restore(); cc.restore();
    }
    const bool can_be_expanded_now = comp.pending_inputs.empty();
    ScopeValue* slot = scope.get(component_name);
    if (!slot) fail("NotFound", component_name);
    *slot = comp;
    if (can_be_expanded_now) eval_component_expand(meta, scope, component_name);
}

void Evaluator::eval_component_expand(const Meta& meta, Scope& scope, const std::string& component_name) {      // eval.rs:496-544
    ScopeValue* c = scope.get(component_name);
    if (!c || c->kind != ScopeValue::Component) fail("NotFound", component_name);
    const std::string tmpl = c->tmpl;
    const std::vector<ReturnValue> values = c->cargs;
    ScopeValue* t = scope.root()->get(tmpl);
    if (!t || t->kind != ScopeValue::Template) fail("NotFound", "template " + tmpl);
    const std::vector<std::string> args = t->args;
    StmtP stmt = t->stmt;
    const std::string tpath = t->path;
    Scope template_scope(true, &scope, current_file + ":" + std::to_string(meta.start));
    for (size_t n = 0; n < args.size(); n++) template_scope.insert(args[n], ScopeValue::from(values[n]));
    Swap fl(current_file, tpath), cc(current_component, expand_full_name(component_name));
    eval_statement(template_scope, *stmt);
    fl.restore(); cc.restore();
}

ReturnValue Evaluator::eval_variable(Scope& scope, const Variable& var) {            // eval.rs:546-594
    const std::string name_sel = expand_selectors(scope, var);
    const std::string name_sel_full = expand_full_name(name_sel);
    if (const Signal* s = signals.get_by_name(name_sel_full)) {
        if (s->has_value && s->value.kind == Value::FieldScalar) return ReturnValue::of(s->value);
        return ReturnValue::of(Value::from_signal(s->id));
    }
    ScopeValue* sv = scope.get(var.name);
    if (!sv) fail("NotFound", name_sel);
    switch (sv->kind) {
        case ScopeValue::Algebra: return ReturnValue::of(sv->a);
        case ScopeValue::Bool: return ReturnValue::of(sv->b);
        case ScopeValue::ListK: {
            const List lcopy = sv->l;               // index expressions may touch the scope
            std::vector<size_t> idx;
            for (auto& sel : var.sels) {
                if (sel.is_pin) fail("InvalidSelector", "Invalid selector ." + sel.name);
                idx.push_back((size_t)eval_expression(scope, *sel.pos).into_u64());
            }
            const List& got = lcopy.get(idx);
            if (got.is_value) return ReturnValue::of(got.v);
            return ReturnValue::of(got);
        }
        default:
            fail("InvalidType", "expected valid value from variable '" + name_sel + "' (current is '" + sv->debug() + "') [nameselfull=" + name_sel_full + "]");
    }
}

std::vector<SignalId> Evaluator::eval_declaration_signals(Scope& scope, SignalType xtype, const Variable& var) {      // eval.rs:838-863
    std::vector<SignalId> pending;
    for (auto& name : generate_selectors(scope, var)) {
        const std::string full = expand_full_name(name);
        if (signals.get_by_name(full)) fail("AlreadyExists", "signal " + full);
        auto d = deferred_signal_values.find(full);
        if (d != deferred_signal_values.end()) {
            signals.insert(full, xtype, &d->second);
            deferred_signal_values.erase(d);
        } else {
            pending.push_back(signals.insert(full, xtype, nullptr));
        }
    }
    return pending;
}

void Evaluator::eval_declaration(Scope& scope, const Statement& s) {                 // eval.rs:865-944
    if (skip_eval(s.meta)) return;
    if (current_component.empty() && mode == Mode::Collect) return;
    const Variable& var = *s.name;
    if (scope.contains_key(var.name)) fail("AlreadyExists", var.name);
    if (s.xtype.kind == VarKind::Var && !s.has_init) {
        if (var.sels.empty()) { ScopeValue u; u.kind = ScopeValue::UndefVar; scope.insert(var.name, u); }
        else {
            ScopeValue l;
            l.kind = ScopeValue::ListK;
            l.l = List::make(expand_indexes(scope, var.sels));
            scope.insert(var.name, l);
        }
    } else if (s.xtype.kind == VarKind::Var) {
        ReturnValue v = eval_expression(scope, *s.value);
        if (s.op != Opcode::Assig) fail("InvalidType", "Unsupported type for var '" + var.name + "' declaration");
        scope.insert(var.name, ScopeValue::from(v));
    } else if (s.xtype.kind == VarKind::Component) {
        eval_component_decl(scope, var);
        if (s.has_init) eval_component_inst(s.meta, scope, expand_selectors(scope, var), *s.value);
    } else if (s.xtype.kind == VarKind::Signal && !s.has_init) {
        // declared when the component was instantiated (eval_component_inst)
    } else {
        fail("NotYetImplemented", "eval_declaration " + debug_string(var));
    }
}

void Evaluator::eval_substitution(Scope& scope, const Statement& s) {                // eval.rs:946-1016
    if (skip_eval(s.meta)) return;
    const Variable& var = *s.name;
    const std::string var_sel = expand_selectors(scope, var);
    if (ScopeValue* v = scope.get(var_sel)) {
        if (v->kind == ScopeValue::UndefComponent) { eval_component_inst(s.meta, scope, var_sel, *s.value); return; }
    }
    Value right = eval_expression(scope, *s.value).into_algebra();
    Value value;
    if (s.op == Opcode::Assig) value = right;
    else {
        Value left = eval_variable(scope, var).into_algebra();
        Opcode op;
        switch (s.op) {
            case Opcode::AssigAdd: op = Opcode::Add; break;
            case Opcode::AssigSub: op = Opcode::Sub; break;
            case Opcode::AssigMul: op = Opcode::Mul; break;
            case Opcode::AssigDiv: op = Opcode::Div; break;
            case Opcode::AssigMod: op = Opcode::Mod; break;
            case Opcode::AssigShiftL: op = Opcode::ShiftL; break;
            case Opcode::AssigShiftR: op = Opcode::ShiftR; break;
            case Opcode::AssigBitAnd: op = Opcode::BitAnd; break;
            case Opcode::AssigBitOr: op = Opcode::BitOr; break;
            case Opcode::AssigBitXor: op = Opcode::BitXor; break;
            default: fail("Unexpected", "assignment operator");
        }
        value = eval_infix(left, op, right);
    }
    if (var.sels.empty()) {
        ScopeValue nv;
        nv.kind = ScopeValue::Algebra;
        nv.a = value;
        scope.update(var.name, nv);
    } else if (!var.sels[0].is_pin) {
        const std::vector<size_t> idx = expand_indexes(scope, var.sels);
        ScopeValue* sv = scope.get(var.name);
        if (!sv) fail("NotFound", var.name);
        if (sv->kind != ScopeValue::ListK) fail("InvalidType", var.name);
        sv->l.set(value, idx);
    }
}

void Evaluator::eval_block(Scope& scope, const Statement& s) {                       // eval.rs:1018-1046
    if (skip_eval(s.meta)) return;
    Scope inner(false, &scope, current_file + ":" + std::to_string(s.meta.start));
    for (auto& st : s.stmts) {
        eval_statement(inner, *st);
        if (inner.has_return()) break;
    }
}

void Evaluator::eval_signal_left(const Meta& meta, Scope& scope, const Variable& signal, Opcode op, const Expression& expr) {       // eval.rs:1048-1161
    // generating constraints: the constraint first, then the assignment; generating the witness: the other way round
    Expression as_expr;
    as_expr.kind = ExprKind::Variable;
    as_expr.meta = meta;
    as_expr.var = std::make_shared<const Variable>(signal);
    if (mode == Mode::GenConstraints && op == Opcode::SignalContrainLeft) eval_signal_eq(meta, scope, as_expr, expr);
    if (!skip_eval(meta)) {
        const std::string signal_sel = expand_selectors(scope, signal);
        const std::string signal_full = expand_full_name(signal_sel);
        const Signal* sg = signals.get_by_name(signal_full);
        if (!sg) fail("NotFound", "Signal " + signal_full);
        const SignalId signal_id = sg->id;
        ReturnValue v = eval_expression(scope, expr);
        if (v.kind != ReturnValue::Algebra) fail("InvalidType", "Cannot assign " + v.debug() + " to signal");
        signals.update(signal_id, v.a);
        std::string component_name;
        if (signal_component(scope, signal, component_name)) {
            ScopeValue* comp = scope.get(component_name);
            if (!comp || comp->kind != ScopeValue::Component)
                fail("NotFound", "signal not found '" + signal.name + "' in scope Meta { start: " + std::to_string(meta.start) + ", end: " + std::to_string(meta.end) + " }");
            bool needs_expansion = false;
            if (!comp->pending_inputs.empty()) {
                auto& p = comp->pending_inputs;
                p.erase(std::remove(p.begin(), p.end(), signal_id), p.end());
                needs_expansion = p.empty();
            }
            if (needs_expansion) eval_component_expand(meta, scope, component_name);      // all inputs are set: run the template now
        }
    }
    if (mode == Mode::GenWitness && op == Opcode::SignalContrainLeft) eval_signal_eq(meta, scope, as_expr, expr);
}

void Evaluator::eval_signal_eq(const Meta& meta, Scope& scope, const Expression& lhe, const Expression& rhe) {         // eval.rs:1186-1250
    (void)meta;
    Value left = eval_expression(scope, lhe).into_algebra();
    Value right = eval_expression(scope, rhe).into_algebra();
    Value constrain = eval_infix(left, Opcode::Sub, right);
    if (mode == Mode::GenWitness) {
        if (!(constrain.kind == Value::FieldScalar && constrain.fs.is_zero()))
            fail("CannotTestConstrain", debug_string(lhe) + "===" + debug_string(rhe) + " => " + signals.format(left) + "===" + signals.format(right));
    } else if (mode == Mode::GenConstraints) {
        if (constrain.kind == Value::FieldScalar) fail("CannotGenerateConstrain", signals.format(left) + "===" + signals.format(right));
        constraints.push(constrain.into_qeq(), debug ? current_file + ":" + std::to_string(meta.start) : std::string());
    }
}

void Evaluator::eval_include(Scope& scope, const std::string& filename) {            // eval.rs:1252-1304
    const std::string full_path = join_path(path, filename);
    std::string code;
    {
        std::ifstream f(full_path, std::ios::binary);
        if (!f) { FeError e("Io", rust_debug_str(full_path) + ", " + rust_debug_str("No such file or directory (os error 2)")); throw e; }
        std::stringstream ss;
        ss << f.rdbuf();
        code = ss.str();
    }
    if (processed_files.count(code)) return;
    processed_files.insert(code);
    Swap fl(current_file, full_path), pt(path, parent_dir(full_path));
    std::vector<BodyElement> elements;
    try { elements = parse_body(code); }
    catch (FeError& e) { Meta m; m.start = e.meta_start; m.end = e.meta_end; note_error(m); throw; }
    eval_body_elements(scope, elements);
    for (auto& b : elements) collected_asts.push_back(b);
    pt.restore(); fl.restore();
}

void Evaluator::eval_body_element(Scope& scope, const BodyElement& b) {              // eval.rs:265-283, 1306-1354
    try {
        switch (b.kind) {
            case BodyKind::Include: eval_include(scope, b.path); break;
            case BodyKind::FunctionDef: {
                ScopeValue f;
                f.kind = ScopeValue::Function;
                f.args = b.args; f.stmt = b.stmt; f.path = current_file;
                scope.insert(b.name, f);
                break;
            }
            case BodyKind::TemplateDef: {
                ScopeValue t;
                t.kind = ScopeValue::Template;
                t.attrs = b.meta.attrs; t.args = b.args; t.stmt = b.stmt; t.path = current_file;
                scope.insert(b.name, t);
                break;
            }
            case BodyKind::Declaration: eval_statement(scope, *b.stmt); break;
        }
    } catch (FeError&) { note_error(b.meta); throw; }
}
void Evaluator::eval_body_elements(Scope& scope, const std::vector<BodyElement>& bes) { for (auto& b : bes) eval_body_element(scope, b); }

void Evaluator::eval_internal_call(Scope& scope, const Statement& s) {               // eval.rs:285-312: dbg!() / dbg_signals!()
    if (s.call_name == "dbg_signals") {
        for (size_t n = 0; n < signals.len(); n++) printf("%s\n", signals.to_string(n).c_str());
        return;
    }
    if (s.call_name == "dbg") {
        printf("DBG ");
        for (auto& p : s.args) {
            if (p->kind == ExprKind::Variable && p->var->sels.empty()) {
                const std::string& n = p->var->name;
                if (n == "CTX") { printf("CTX => %s %s:%llu\n", current_component.c_str(), current_file.c_str(), (unsigned long long)p->meta.start); continue; }
                if (n == "SCOPE") { for (auto& kv : scope.vars) printf("%s=%s ", kv.first.c_str(), kv.second.debug().c_str()); continue; }
                if (n == "TRACEON") { debug = true; continue; }
                if (n == "TRACEOFF") { debug = false; continue; }
            }
            ReturnValue v = eval_expression(scope, *p);
            printf("%s => %s ", debug_string(*p).c_str(), v.kind == ReturnValue::Algebra ? signals.format(v.a).c_str() : v.debug().c_str());
        }
        printf("\n");
        return;
    }
    fail("NotFound", "internal funcion " + s.call_name + "!");
}

void Evaluator::eval_statement(Scope& scope, const Statement& s) {                   // eval.rs:218-263
    try {
        switch (s.kind) {
            case StmtKind::IfThenElse: {                                              // eval.rs:709-735
                if (skip_eval(s.meta)) return;
                ReturnValue c = eval_expression(scope, *s.cond);
                if (c.kind != ReturnValue::Bool) fail("InvalidType", "if condition is not boolean");
                if (c.b) eval_statement(scope, *s.xthen);
                else if (s.xelse) eval_statement(scope, *s.xelse);
                return;
            }
            case StmtKind::For: {                                                     // eval.rs:737-783
                if (skip_eval(s.meta)) return;
                Scope inner(false, &scope, current_file + ":" + std::to_string(s.meta.start));
                eval_statement(inner, *s.init);
                while (true) {
                    ReturnValue c = eval_expression(inner, *s.cond);
                    if (c.kind != ReturnValue::Bool) fail("InvalidType", "for loop condition is not boolean");
                    if (!c.b) break;
                    eval_statement(inner, *s.body);
                    if (inner.has_return()) break;
                    eval_statement(inner, *s.step);
                }
                return;
            }
            case StmtKind::While: {                                                   // eval.rs:785-822
                if (skip_eval(s.meta)) return;
                Scope inner(false, &scope, current_file + ":" + std::to_string(s.meta.start));
                while (true) {
                    ReturnValue c = eval_expression(inner, *s.cond);
                    if (c.kind != ReturnValue::Bool) fail("InvalidType", "while loop condition is not boolean");
                    if (!c.b) break;
                    eval_statement(inner, *s.body);
                    if (inner.has_return()) break;
                }
                return;
            }
            case StmtKind::Return:                                                    // eval.rs:824-836
                if (skip_eval(s.meta)) return;
                scope.set_return(eval_expression(scope, *s.value));
                return;
            case StmtKind::Declaration: eval_declaration(scope, s); return;
            case StmtKind::Substitution: eval_substitution(scope, s); return;
            case StmtKind::Block: eval_block(scope, s); return;
            case StmtKind::SignalLeft: eval_signal_left(s.meta, scope, *s.name, s.op, *s.value); return;
            case StmtKind::SignalRight:                                               // eval.rs:1163-1184
                if (skip_eval(s.meta)) return;
                eval_signal_left(s.meta, scope, *s.name, s.op == Opcode::SignalContrainRight ? Opcode::SignalContrainLeft : Opcode::SignalWireLeft, *s.value);
                return;
            case StmtKind::SignalEq: eval_signal_eq(s.meta, scope, *s.lhe, *s.value); return;
            case StmtKind::InternalCall: eval_internal_call(scope, s); return;
        }
    } catch (FeError&) { note_error(s.meta); throw; }
}

std::vector<std::string> Evaluator::generate_selectors(Scope& scope, const Variable& var) {         // eval.rs:1381-1418
    std::vector<uint64_t> sizes;
    for (auto& sel : var.sels) {
        if (sel.is_pin) fail("InvalidType", "selectors for " + var.name);
        sizes.push_back(eval_expression(scope, *sel.pos).into_u64());
    }
    std::vector<std::string> out;
    std::vector<uint64_t> stack(sizes.size(), 0);
    uint64_t total = 1;
    for (uint64_t s : sizes) {
        if (s && total > ((uint64_t)1 << 28) / s) fail("InvalidType", "array " + var.name + " is too large");
        total *= s;
    }
    if (total == 0) return out;
    for (uint64_t k = 0; k < total; k++) {
        std::string name = var.name;
        for (uint64_t i : stack) name += "[" + std::to_string(i) + "]";
        out.push_back(name);
        for (size_t d = sizes.size(); d-- > 0;) { if (++stack[d] < sizes[d]) break; stack[d] = 0; }
    }
    return out;
}

std::string Evaluator::expand_selectors(Scope& scope, const Variable& v, int limit) {                // eval.rs:1420-1446
    std::string out = v.name;
    for (size_t i = 0; i < v.sels.size(); i++) {
        if (limit >= 0 && (int)i == limit) return out;
        const Selector& sel = v.sels[i];
        if (sel.is_pin) out += "." + sel.name;
        else out += "[" + std::to_string(eval_expression(scope, *sel.pos).into_u64()) + "]";
    }
    return out;
}

std::vector<size_t> Evaluator::expand_indexes(Scope& scope, const std::vector<Selector>& sels) {    // eval.rs:1448-1464
    std::vector<size_t> idx;
    for (auto& sel : sels) {
        if (sel.is_pin) fail("InvalidSelector", "Invalid selector ." + sel.name);
        idx.push_back((size_t)eval_expression(scope, *sel.pos).into_u64());
    }
    return idx;
}

bool Evaluator::signal_component(Scope& scope, const Variable& signal, std::string& out) {           // eval.rs:1466-1494: a[1].b[1].c -> a[1].b[1]
    size_t last_pin = signal.sels.size();
    bool found = false;
    while (!found && last_pin > 0) {
        if (signal.sels[last_pin - 1].is_pin) found = true;
        else last_pin--;
    }
    if (!found) return false;
    out = expand_selectors(scope, signal, (int)last_pin - 1);
    return true;
}

// ---- optimiser (optimizer/mod.rs:14-179) ----------------------------------------------------------------------------------
void optimize(const Constraints& in, const std::vector<SignalId>& irreductible, Constraints& out, std::vector<SignalId>& removed) {
    struct Change { SignalId s; FS f; };
    std::map<SignalId, Change> replaces;
    std::vector<size_t> rm;
    std::set<SignalId> irr(irreductible.begin(), irreductible.end());
    for (size_t n = 0; n < in.len(); n++) {
        QEQ c = in.q[n];
        // [k one][b] + [c] and [a][k one] + [c] are linear: fold them into the c part (only to DETECT aliases; the stored
        // constraint is untouched, as in the reference)
        if (c.a.t.size() == 1 && c.a.t[0].first == SIGNAL_ONE) { QEQ r; r.c = c.c.add_lc(c.b.mul_fs(c.a.t[0].second)); c = r; }
        else if (c.b.t.size() == 1 && c.b.t[0].first == SIGNAL_ONE) { QEQ r; r.c = c.c.add_lc(c.a.mul_fs(c.b.t[0].second)); c = r; }
        if (!(c.a.t.empty() && c.b.t.empty() && c.c.t.size() == 2)) continue;
        const auto& first = c.c.t[0];
        const auto& second = c.c.t[1];
        const bool fi = irr.count(first.first) != 0, si = irr.count(second.first) != 0;
        const std::pair<SignalId, FS>*search, *replace;
        if (!fi && si) { search = &first; replace = &second; }
        else if (fi && !si) { search = &second; replace = &first; }
        else if (!fi && !si) { if (first.first > second.first) { search = &first; replace = &second; } else { search = &second; replace = &first; } }
        else continue;
        const SignalId search_s = search->first;
        SignalId replace_s = replace->first;
        FS replace_f = replace->second.div(search->second).neg();
        if (replaces.count(search_s)) continue;
        for (auto it = replaces.find(replace_s); it != replaces.end(); it = replaces.find(replace_s)) {
            replace_s = it->second.s;
            replace_f = replace_f.mul(it->second.f);
        }
        replaces[search_s] = Change{replace_s, replace_f};
        rm.push_back(n);
    }
    // [s] -> f1 [r] with [r] -> f2 [r2] becomes [s] -> f1 f2 [r2], until nothing changes
    bool any = true;
    while (any) {
        any = false;
        for (auto& kv : replaces) {
            auto it = replaces.find(kv.second.s);
            if (it == replaces.end()) continue;
            const Change r2 = it->second;
            kv.second = Change{r2.s, kv.second.f.mul(r2.f)};
            any = true;
        }
    }
    out = Constraints();
    size_t ri = 0;
    auto apply = [&](LC& lc) {
        for (auto& e : lc.t) {
            auto it = replaces.find(e.first);
            if (it != replaces.end()) e = std::make_pair(it->second.s, e.second.mul(it->second.f));
        }
    };
    for (size_t n = 0; n < in.len(); n++) {
        if (ri < rm.size() && rm[ri] == n) { ri++; continue; }
        QEQ c = in.q[n];
        apply(c.a); apply(c.b); apply(c.c);
        out.push(c);
    }
    removed.clear();
    for (auto& kv : replaces) removed.push_back(kv.first);      // std::map iterates in ascending order = the reference's sort()
}

}  // namespace zafe
