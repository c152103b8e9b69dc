// The syntax tree of za's circuit language: /root/reference/parser/src/ast.rs.  Variant and field ORDER follows the
// reference because the tree is stored inside every proving.key as bincode(Vec<BodyElementP>) (format.rs:231-234) and
// read back by `prove` to regenerate the witness (helper.rs:91-108).
#pragma once
#include <memory>
#include "algebra.hpp"

namespace zafe {

struct Meta {                                   // ast.rs:19-24: byte offsets into the preprocessed text + attribute tags
    uint64_t start = 0, end = 0;
    std::vector<std::string> attrs;
    bool has_tag(const char* t) const { for (auto& a : attrs) if (a == t) return true; return false; }
};

struct Expression;
struct Statement;
typedef std::shared_ptr<const Expression> ExprP;
typedef std::shared_ptr<const Statement> StmtP;

struct Selector {                               // ast.rs:45-48
    bool is_pin = false;
    Meta meta;
    std::string name;                           // Pin
    ExprP pos;                                  // Index
};
struct Variable {                               // ast.rs:51-55
    Meta meta;
    std::string name;
    std::vector<Selector> sels;
};
typedef std::shared_ptr<const Variable> VarP;

enum class ExprKind : uint32_t { FunctionCall, Variable, Number, PrefixOp, InfixOp, Array };     // ast.rs:58-89
struct Expression {
    ExprKind kind;
    Meta meta;
    std::string name;                           // FunctionCall
    std::vector<ExprP> list;                    // FunctionCall args, Array values
    VarP var;                                   // Variable
    BigDigits number;                           // Number (never negative: the grammar has no signed literal)
    Opcode op = Opcode::Add;                    // PrefixOp, InfixOp
    ExprP lhe, rhe;
};

enum class SignalType : uint32_t { Output, PublicInput, PrivateInput, Internal };                 // ast.rs:184-190
enum class VarKind : uint32_t { Empty, Var, Signal, Component };                                  // ast.rs:192-198
struct VariableType { VarKind kind = VarKind::Empty; SignalType signal = SignalType::Internal; };

enum class StmtKind : uint32_t { IfThenElse, For, While, Return, Declaration, Substitution, Block, SignalLeft, SignalRight, SignalEq, InternalCall };
struct Statement {                              // ast.rs:92-158
    StmtKind kind;
    Meta meta;
    ExprP cond;                                 // IfThenElse xif, For cond, While cond
    StmtP xthen, xelse;                         // IfThenElse
    StmtP init, step, body;                     // For (body also While)
    ExprP value;                                // Return, Substitution, SignalLeft, SignalRight; Declaration init value; SignalEq rhe
    ExprP lhe;                                  // SignalEq
    VariableType xtype;                         // Declaration
    VarP name;                                  // Declaration, Substitution, SignalLeft, SignalRight
    bool has_init = false;                      // Declaration
    Opcode op = Opcode::Assig;                  // Declaration init op, Substitution, SignalLeft/Right/Eq
    std::vector<StmtP> stmts;                   // Block
    std::string call_name;                      // InternalCall
    std::vector<ExprP> args;                    // InternalCall
};

enum class BodyKind : uint32_t { Include, FunctionDef, TemplateDef, Declaration };                // ast.rs:161-181
struct BodyElement {
    BodyKind kind;
    Meta meta;
    std::string path;                           // Include
    std::string name;                           // FunctionDef, TemplateDef
    std::vector<std::string> args;
    StmtP stmt;                                 // FunctionDef / TemplateDef body, Declaration
};

// parser.cpp — parser/src/parse.rs + lang.lalrpop
std::string preprocess(const std::string& text);
std::vector<BodyElement> parse_body(const std::string& text);
StmtP parse_statement(const std::string& text);
ExprP parse_expression(const std::string& text);
BodyElement parse_body_element(const std::string& text);
// display.rs (Debug impls)
std::string debug_string(const Expression& e);
std::string debug_string(const Statement& s);
std::string debug_string(const Variable& v);
std::string debug_string(const BodyElement& b);
// bincode.cpp — bincode 1.x image of Vec<BodyElementP> as serde derives it
std::vector<uint8_t> ast_serialize(const std::vector<BodyElement>& body);
std::vector<BodyElement> ast_deserialize(const uint8_t* data, size_t len);

}  // namespace zafe
