// za's front-end as a C++ library: evaluator (constraint generation and witness generation), signal table, constraint
// list and the single-pass optimiser.  References: /root/reference/compiler/src/evaluator/{eval.rs, scope.rs, types.rs},
// /root/reference/compiler/src/types/{signal.rs, constraint.rs}, /root/reference/compiler/src/optimizer/mod.rs.
#pragma once
#include <map>
#include <set>
#include <unordered_map>
#include "ast.hpp"

namespace zafe {

struct Signal {                                     // types/signal.rs:50-56
    SignalId id = 0;
    SignalType xtype = SignalType::Internal;
    std::string full_name;
    bool has_value = false;
    Value value;
    static size_t dots(const std::string& s) { return (size_t)std::count(s.begin(), s.end(), '.'); }
    bool is_main_public_input() const {             // signal.rs:59-63
        return dots(full_name) == 1 && (xtype == SignalType::Output || xtype == SignalType::PublicInput);
    }
    bool is_main_input() const {                    // signal.rs:64-70
        return dots(full_name) == 1 && (xtype == SignalType::Output || xtype == SignalType::PublicInput || xtype == SignalType::PrivateInput);
    }
};

struct Signals {                                    // signal.rs:73-190; signal 0 is "one"
    std::vector<Signal> ids;
    std::unordered_map<std::string, SignalId> names;
    Signals() { insert("one", SignalType::PublicInput, nullptr); }
    size_t len() const { return ids.size(); }
    const Signal* get_by_name(const std::string& n) const { auto it = names.find(n); return it == names.end() ? nullptr : &ids[it->second]; }
    SignalId insert(const std::string& full_name, SignalType t, const Value* v) {
        Signal s;
        s.id = ids.size(); s.xtype = t; s.full_name = full_name;
        if (v) { s.has_value = true; s.value = *v; }
        ids.push_back(s);
        names[full_name] = s.id;
        return s.id;
    }
    void update(SignalId id, const Value& v) { ids[id].has_value = true; ids[id].value = v; }
    std::vector<std::string> main_public_input_names() const {
        std::vector<std::string> v;
        for (size_t i = 1; i < ids.size(); i++) if (ids[i].is_main_public_input()) v.push_back(ids[i].full_name);
        return v;
    }
    std::vector<SignalId> main_input_ids() const {
        std::vector<SignalId> v;
        for (size_t i = 1; i < ids.size(); i++) if (ids[i].is_main_input()) v.push_back(i);
        return v;
    }
    std::string name_of(SignalId id) const { return id < ids.size() ? ids[id].full_name : std::string("unwnown"); }
    std::string format(const Value& a) const {      // signal.rs:167-178
        auto nm = [this](SignalId id) { return name_of(id); };
        return a.kind == Value::FieldScalar ? a.fs.to_string() : a.kind == Value::LinearCombination ? a.lc.format(nm) : a.qeq.format(nm);
    }
    std::string to_string(SignalId id) const;       // "{name}:{type}:{value}", signal.rs:162-165
};

struct Constraints {                                // types/constraint.rs
    std::vector<QEQ> q;
    std::vector<std::string> debug;
    size_t len() const { return q.size(); }
    size_t push(const QEQ& e, const std::string& dbg = std::string()) { q.push_back(e); debug.push_back(dbg); return q.size() - 1; }
    // constraint.rs:29-67; returns "" or the reference's message
    std::string satisfies_with_signals(const Signals& s) const;
};

// optimizer/mod.rs:14-179
void optimize(const Constraints& in, const std::vector<SignalId>& irreductible, Constraints& out, std::vector<SignalId>& removed);

struct List {                                       // evaluator/types.rs:7-69
    bool is_value = true;
    Value v;
    std::vector<List> items;
    static List make(const std::vector<size_t>& sizes, size_t at = 0);
    const List& get(const std::vector<size_t>& idx, size_t at = 0) const;
    void set(const Value& value, const std::vector<size_t>& idx, size_t at = 0);
    std::string debug() const;
};

struct ReturnValue {                                // evaluator/types.rs:71-76
    enum Kind { Bool, Algebra, ListK } kind = Algebra;
    bool b = false;
    Value a;
    List l;
    static ReturnValue of(bool x) { ReturnValue r; r.kind = Bool; r.b = x; return r; }
    static ReturnValue of(const Value& x) { ReturnValue r; r.kind = Algebra; r.a = x; return r; }
    static ReturnValue of(const List& x) { ReturnValue r; r.kind = ListK; r.l = x; return r; }
    std::string debug() const;
    const Value& into_algebra() const;
    bool into_bool() const;
    const FS& into_fs() const;
    uint64_t into_u64() const;
};

struct ScopeValue {                                 // evaluator/scope.rs:13-41
    enum Kind { UndefVar, UndefComponent, Bool, Algebra, Function, Template, Component, ListK } kind = UndefVar;
    bool b = false;
    Value a;
    List l;
    std::vector<std::string> args;                  // Function, Template
    StmtP stmt;
    std::string path;                               // Function, Template, Component
    std::vector<std::string> attrs;                 // Template
    std::string tmpl;                               // Component: template name
    std::vector<ReturnValue> cargs;                 // Component: evaluated arguments
    std::vector<SignalId> pending_inputs;
    static ScopeValue from(const ReturnValue& r);
    std::string debug() const;
};

struct Scope {                                      // evaluator/scope.rs:59-200
    bool start;
    Scope* prev;
    std::string pos;
    bool has_ret = false;
    ReturnValue ret;
    std::unordered_map<std::string, ScopeValue> vars;
    Scope(bool start_, Scope* prev_, const std::string& pos_) : start(start_), prev(prev_), pos(pos_) {}
    Scope* root() { Scope* it = this; while (it->prev) it = it->prev; return it; }
    Scope* start_scope() { Scope* it = this; while (!it->start) it = it->prev; return it; }
    void insert(const std::string& k, const ScopeValue& v);
    ScopeValue* get(const std::string& k);
    bool contains_key(const std::string& k) { return get(k) != nullptr; }
    void update(const std::string& k, const ScopeValue& v);
    void set_return(const ReturnValue& v) { Scope* s = start_scope(); s->has_ret = true; s->ret = v; }
    bool take_return(ReturnValue& out) { Scope* s = start_scope(); if (!s->has_ret) return false; out = s->ret; s->has_ret = false; return true; }
    bool has_return() { return start_scope()->has_ret; }
};

enum class Mode { Collect, GenConstraints, GenWitness };     // eval.rs:34-38

struct ErrorContext { std::string file, component, function; uint64_t start = 0, end = 0; bool set = false; };

struct Evaluator {                                  // eval.rs:50-80
    Mode mode;
    Signals signals;
    Constraints constraints;
    std::string current_file, current_component, current_function;
    bool in_function = false;
    std::set<std::string> processed_files;          // the reference keys on a Blake2b digest of the text; the text itself is the same test
    std::vector<BodyElement> collected_asts;
    std::string path = ".";
    std::unordered_map<std::string, Value> deferred_signal_values;
    bool debug = false;
    ErrorContext last_error;
    explicit Evaluator(Mode m) : mode(m) {}

    void eval_inline(Scope& scope, const std::string& code);
    void eval_template(Scope& scope, const std::string& template_name);
    void eval_file(Scope& scope, const std::string& path, const std::string& filename);
    void eval_asts(Scope& scope, const std::vector<BodyElement>& asts);
    void set_deferred_value(const std::string& full_name, const Value& v) { deferred_signal_values[full_name] = v; }

   private:
    bool skip_eval(const Meta& m) const { return mode == Mode::GenConstraints && m.has_tag("w"); }
    ReturnValue eval_expression(Scope& scope, const Expression& e);
    void eval_statement(Scope& scope, const Statement& s);
    void eval_body_element(Scope& scope, const BodyElement& b);
    void eval_body_elements(Scope& scope, const std::vector<BodyElement>& bes);
    ReturnValue eval_function_call(const Meta& meta, Scope& scope, const std::string& name, const std::vector<ExprP>& params);
    void eval_component_decl(Scope& scope, const Variable& name);
    void eval_component_inst(const Meta& meta, Scope& scope, const std::string& component_name, const Expression& init);
    void eval_component_expand(const Meta& meta, Scope& scope, const std::string& component_name);
    ReturnValue eval_variable(Scope& scope, const Variable& var);
    ReturnValue eval_infix_op(Scope& scope, const Expression& e);
    void eval_declaration(Scope& scope, const Statement& s);
    std::vector<SignalId> eval_declaration_signals(Scope& scope, SignalType xtype, const Variable& var);
    void eval_substitution(Scope& scope, const Statement& s);
    void eval_block(Scope& scope, const Statement& s);
    void eval_signal_left(const Meta& meta, Scope& scope, const Variable& signal, Opcode op, const Expression& expr);
    void eval_signal_eq(const Meta& meta, Scope& scope, const Expression& lhe, const Expression& rhe);
    void eval_include(Scope& scope, const std::string& filename);
    void eval_internal_call(Scope& scope, const Statement& s);
    std::vector<std::string> generate_selectors(Scope& scope, const Variable& var);
    std::string expand_selectors(Scope& scope, const Variable& v, int limit = -1);
    std::vector<size_t> expand_indexes(Scope& scope, const std::vector<Selector>& sels);
    bool signal_component(Scope& scope, const Variable& signal, std::string& out);
    std::string expand_full_name(const std::string& s) const { return current_component.empty() ? s : current_component + "." + s; }
    void note_error(const Meta& meta);
};

// Rust's {:?} of a str
std::string rust_debug_str(const std::string& s);
// Debug text of an error as the bindings return it (format!("{:?}", err) of prover::groth16::error::Error)
std::string error_debug(const FeError& e);

}  // namespace zafe
