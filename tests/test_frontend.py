"""Front-end: parser, evaluator, optimiser and `za test` of libza2c, checked against the reference's own
unit tests.  Every case below is one of the reference's test inputs with the expectation its test asserts:
  parser/src/ast.rs:254-390        Debug text of expressions / statements / body elements (display.rs)
  parser/src/parse.rs:96-111       comment preprocessor
  compiler/src/evaluator/test.rs   scope values, signal table, constraint text, witness checks, signal ordering
  compiler/src/optimizer/mod.rs:186-233   linear-constraint substitution
  interop/circuits/circomlib/za_test/*.za   `za test` over circomlib (read from /root/reference when it is there)
No GPU: the seam (za2c_parse / za2c_eval / za2c_test, include/za2c.h) is host code."""
import os

import pytest

from za_b200 import za2c

REF = "/root/reference"


# ---- parser ---------------------------------------------------------------------------------------------------------
EXPRESSIONS = [
    ("255", "255"), ("-255", "(- 255)"), ("0xFF", "255"), ("0xff", "255"),
    ("- 1 | 2 ^ 3 & 4 << 5 + 6 * 7", "((- 1) | (2 ^ (3 & (4 << (5 + (6 * 7))))))"),
    ("(a | b) ^ c & d << e + f * g", "((a | b) ^ (c & (d << (e + (f * g)))))"),
    ("a == b && c == d || e == f", "(((a == b) && (c == d)) || (e == f))"),
    ("a > b || c < d || e >=f || g<=h || i==j || k !=l",
     "((((((a > b) || (c < d)) || (e >= f)) || (g <= h)) || (i == j)) || (k != l))"),
    ("(a == b && c == d) || e == f", "(((a == b) && (c == d)) || (e == f))"),
    ("a", "a"), ("a[5]", "a[5]"), ("a.b", "a.b"), ("a[5].b", "a[5].b"), ("a[c[1*1].d].b", "a[c[(1 * 1)].d].b"),
    ("f(a*1,b(),c(1*2))", "f((a * 1),b(),c((1 * 2)))"),
]

STATEMENTS = [
    "var a;", "var a = b;", "component a = b;", "signal a;", "signal input a;", "signal private input a;",
    "signal output a;",
    "a = b;", "a -= b;", "a *= b;", "a /= b;", "a %= b;", "a >>= b;", "a <<= b;", "a |= b;", "a &= b;", "a[1].a = b;",
    "if (a) {b = c;}", "if (a) {b = c;} else {b = c;}", "if (a) {b = c;} else if (b) {d = e;}",
    "if (a) {b = c;} else if (b) {d = e;} else {i = k;}",
    "while (a) {b += c;}",
    "for (a = u;(a < b);a += d) {b += c;}", "for (var a = u;(a < b);a += d) {b += c;}",
    "return a;",
    "a <-- b;", "a --> b;", "a ==> b;", "a <== b;", "a === b;",
    "if (a) {b = c; b = c;}", "if (a) {b = c; b = c;} else {a = a; b = a;}",
]

BODY_ELEMENTS = ['include "hola";', "function f1(a,b,c) {a += b;}", "template f1(a,b,c) {a += b;}", "var a;"]


@pytest.mark.parametrize("text,expected", EXPRESSIONS)
def test_expression_debug_text(text, expected):
    assert za2c.parse(za2c.EXPRESSION, text) == expected


@pytest.mark.parametrize("text", STATEMENTS)
def test_statement_round_trips_through_its_debug_text(text):
    assert za2c.parse(za2c.STATEMENT, text) == text


@pytest.mark.parametrize("text", BODY_ELEMENTS)
def test_body_element_round_trips(text):
    assert za2c.parse(za2c.BODY_ELEMENT, text) == text


@pytest.mark.parametrize("text,expected", [
    ("helo // jalo", "helo        "),
    ("helo // jalo\nfoo", "helo        \nfoo"),
    ("helo /* jalo */\nfoo", "helo           \nfoo"),
    ("helo /* jalo \n*/foo", "helo            foo"),
    ("helo /* // */foo", "helo         foo"),
    ("a /*#[foo]#*/ b", "a   #[foo]    b"),
])
def test_preprocessor_comments(text, expected):
    assert za2c.parse(za2c.PREPROCESS, text) == expected


def test_unterminated_comment_and_syntax_errors_are_errors():
    with pytest.raises(TypeError):
        za2c.parse(za2c.PREPROCESS, "a /* b *")        # parse.rs:58-63: only a '*' at end of input inside a block comment
    assert za2c.parse(za2c.PREPROCESS, "a /* b") == "a     "
    with pytest.raises(TypeError):
        za2c.parse(za2c.STATEMENT, "var = ;")


def test_syntax_tree_survives_its_bincode_image():
    """helper::setup stores the tree in the proving key (format.rs:223-246) and helper::prove evaluates that stored tree."""
    src = open(os.path.join(os.path.dirname(__file__), "golden", "factor.za")).read() if os.path.exists(
        os.path.join(os.path.dirname(__file__), "golden", "factor.za")) else \
        "template T() { signal private input p; signal output r; r <== p*p; }\ncomponent main = T();"
    image = za2c.parse(za2c.BODY_BINCODE, src)
    assert len(image) > 0 and len(image) % 2 == 0 and bytes.fromhex(image)


# ---- evaluator: scope values ------------------------------------------------------------------------------------------
SCOPE_CASES = [
    ("var i = 1;\nvar j = 5;\nvar k = j;", {"i": "Algebra(1)", "j": "Algebra(5)", "k": "Algebra(5)"}),
    ("var i = 1+2*3;\nvar j = i-3;", {"i": "Algebra(7)", "j": "Algebra(4)"}),
    ("var iyes = 1==1;\nvar ino = 1!=1;\nvar byes = iyes==iyes;\nvar bno = iyes!=iyes;",
     {"iyes": "Bool(true)", "ino": "Bool(false)", "byes": "Bool(true)", "bno": "Bool(false)"}),
    ("var yes1 = 1<2;\nvar no1 = 1 >2;\nvar yes2 = 1<=2;\nvar no2 = 1>=2;",
     {"yes1": "Bool(true)", "no1": "Bool(false)", "yes2": "Bool(true)", "no2": "Bool(false)"}),
    ("var i = -5;\nvar j=-i;", {"j": "Algebra(5)"}),
    ("function f(a) {\n return a;\n}\nvar k=f(1);", {"k": "Algebra(1)"}),
    ("function f(a,b) {\n return a+b; }\nvar k=f(1,2);", {"k": "Algebra(3)"}),
    ("function f(a) {\n var t=5;\n t+=a;\n t-=2;\n t*=2;\n return t;\n}\nvar k=f(2);", {"k": "Algebra(10)"}),
    ("function fact(N) {\n var f=1;\n for (var i=1;i<=N;i+=1) {\n f = f * i;\n } return f;\n}\nvar out=fact(10);",
     {"out": "Algebra(3628800)"}),
    ("function fact(N) {\n var f=1;\n for (var i=1;i<=N;i+=1) {\n return N; f = f * i;\n }\n return f;\n}\nvar out=fact(10);",
     {"out": "Algebra(10)"}),
    ("function fact(N) {\n var f=1;\n var i=1;\n while (i<=N) {\n f = f * i;\n i+=1;\n }\n return f;\n}\nvar out=fact(10);",
     {"out": "Algebra(3628800)"}),
    ("function fact(N) {\n var f=1;\n var i=1;\n while (i<=N) {\n return N;\n f = f * i;\n i+=1;\n }\n return f;\n}\n"
     "var out=fact(10);", {"out": "Algebra(10)"}),
    ("function test(v) {\n if (v==1) {\n return 1;\n }\n return 2;\n}\nvar out1=test(1);\nvar out2=test(2);",
     {"out1": "Algebra(1)", "out2": "Algebra(2)"}),
    ("function test(v){\n if (v==1) {\n return 1;\n } else {\n return 2;\n }\n}\nvar out1=test(1);\nvar out2=test(2);",
     {"out1": "Algebra(1)", "out2": "Algebra(2)"}),
    ("function test(){\n var M = [[1,2,3],[4,5,6],[7,8,9]];\n return M[1][1];\n}\nvar out=test();", {"out": "Algebra(5)"}),
    ("function test(){\n var M[5][5];\n M[3][1] = 5;\n M[1][2] = 7;\n return M[3][1] + M[1][2];\n}\nvar out=test();",
     {"out": "Algebra(12)"}),
    ("function f() {\n var k[1];\n k[0]=6;\n return k[0];\n}\nvar out=f();", {"out": "Algebra(6)"}),
    ("var P=[1,2,3,4,5];\nvar out=P[2];", {"out": "Algebra(3)"}),
]


@pytest.mark.parametrize("source,expected", SCOPE_CASES)
def test_scope_values(source, expected):
    scope = za2c.evaluate(source=source)["scope"]
    for name, value in expected.items():
        assert scope.get(name) == value, (name, scope)


# ---- evaluator: signals and constraints -------------------------------------------------------------------------------
def _signals(r):
    return {s.split(":", 1)[0]: s for s in r["signals"]}


def test_template_signal_base():
    r = za2c.evaluate(source="""
        template t() {
            signal a;
            signal input b;
            signal private input c;
            signal output d;
        }
        component main=t();""")
    s = _signals(r)
    assert s["main.a"] == "main.a:Internal:None"
    assert s["main.b"] == "main.b:PublicInput:None"
    assert s["main.c"] == "main.c:PrivateInput:None"
    assert s["main.d"] == "main.d:Output:None"
    assert "main.e" not in s


CONSTRAINT_CASES = [
    ("template t() {\n signal input a;\n signal input b;\n signal private input c;\n c === 5 * a * b  + 5;\n}\ncomponent main=t();",
     ["[-5main.a]*[1main.b]+[-5one+1main.c]"], {}),
    ("template t() {\n signal a;\n var i = 1;\n #[w] i=2;\n a === i;\n}\ncomponent main=t();", ["[ ]*[ ]+[1main.a-1one]"], {}),
    ("template t() {\n signal in;\n signal const;\n const <-- 2;\n 2 === 1 + in * const ;\n}\ncomponent main=t();",
     ["[ ]*[ ]+[-2main.in+1one]"], {"main.const": "main.const:Internal:Some(2)"}),
    ("template t() {\n signal in;\n signal out;\n out <== in;\n out === 1;\n}\ncomponent main=t();",
     ["[ ]*[ ]+[1main.out-1main.in]", "[ ]*[ ]+[1main.out-1one]"], {}),
    ("template t() {\n signal in;\n signal const;\n const <== 2;\n 2 === 1 + in * const ;\n}\ncomponent main=t();",
     ["[ ]*[ ]+[1main.const-2one]", "[ ]*[ ]+[-2main.in+1one]"], {}),
    ("template t() {\n signal in[2][2];\n for (var i=0;i<2;i+=1) {\n in[i][0] <-- i+2 ;\n in[i][1] <--i+3 ; \n }\n}\ncomponent main=t();",
     [], {"main.in[0][0]": "main.in[0][0]:Internal:Some(2)", "main.in[0][1]": "main.in[0][1]:Internal:Some(3)",
          "main.in[1][0]": "main.in[1][0]:Internal:Some(3)", "main.in[1][1]": "main.in[1][1]:Internal:Some(4)"}),
    ("template t() {\n signal in[2][2];\n signal s;\n in[1][0] + in[0][1] === 0 ;\n}\ncomponent main=t();",
     ["[ ]*[ ]+[1main.in[1][0]+1main.in[0][1]]"], {}),
    ("template t() {\n signal in[2];\n signal s;\n in[0] <== 1 ;\n in[0] === in[1];\n}\ncomponent main=t();",
     ["[ ]*[ ]+[1main.in[0]-1one]", "[ ]*[ ]+[-1main.in[1]+1one]"], {}),
    ("template t0() {\n signal t0in;\n t0in === 5;\n}\ntemplate t1() {\n signal t1in;\n component T0 = t0();\n t1in <== T0.t0in;\n}\n"
     "component main=t1();", ["[ ]*[ ]+[1main.T0.t0in-5one]"], {}),
    ("template t0() {\n signal t0in;\n t0in === 5;\n}\ntemplate t1() {\n signal t1in;\n component T0[1];\n"
     " for (var k=0;k<1;k +=1) {\n T0[k] = t0();\n t1in <== T0[k].t0in;\n }\n}\ncomponent main=t1();",
     ["[ ]*[ ]+[1main.T0[0].t0in-5one]"], {}),
]


@pytest.mark.parametrize("source,constraints,signals", CONSTRAINT_CASES)
def test_constraints_and_signal_values(source, constraints, signals):
    r = za2c.evaluate(source=source)
    assert r["constraints"][:len(constraints)] == constraints
    s = _signals(r)
    for name, text in signals.items():
        assert s[name] == text


def test_signal_ordering():
    """evaluator/test.rs:738-769: ids are one, outputs, public inputs, private inputs, internals (what bellman's input /
    aux split and the verifier's ic order follow)."""
    r = za2c.evaluate(source="""
        template t() {
            signal input pub1;
            signal private input priv1;
            signal int1;
            signal output out;
            signal private input priv2;
            signal int2;
            signal input pub2;
            out <== pub1 + pub2 + int1 + int2 + priv1 + priv2;
        }
        component main = t();""")
    names = [s.split(":", 1)[0] for s in r["signals"]]
    assert names == ["one", "main.out", "main.pub1", "main.pub2", "main.priv1", "main.priv2", "main.int1", "main.int2"]


# ---- evaluator: witness mode ----------------------------------------------------------------------------------------
WITNESS_PASS = [
    ("template t0() {\n signal t0in;\n t0in <-- 5;\n t0in === 5;\n}\ncomponent main = t0();", None),
    ("template t1() {\n signal input a;\n a === 2;\n}\ntemplate t0() {\n component c1 = t1();\n c1.a <-- 2;\n}\ncomponent main = t0();", None),
    ("template t2() {\n signal input in[1];\n signal output out;\n out <== in[0] * 3;\n}\n"
     "template t1() {\n signal input in[1];\n signal output out;\n component c2 = t2();\n c2.in[0] <==  in[0];\n out <== c2.out * 7;\n}\n"
     "template t0() {\n component c1[1];\n c1[0] = t1();\n c1[0].in[0] <== 2;\n c1[0].out === 2*3*7;\n}\ncomponent main = t0();", None),
    ("template t() {\n signal input a;\n signal input b;\n a === 2 * b;\n}\ncomponent main = t();", {"main.a": 4, "main.b": 2}),
    ("template t() {\n signal input p;\n signal output out;\n out <== 1-p;\n}\ncomponent main = t();", {"main.p": 2}),
]

WITNESS_FAIL = [
    "template t0() {\n signal t0in;\n t0in === 5;\n}\ncomponent main = t0();",
    "template t0() {\n signal t0in;\n t0in <-- 2;\n t0in === 5;\n}\ncomponent main = t0();",
    "template t1() {\n signal input a;\n a === 3;\n}\ntemplate t0() {\n component c1 = t1();\n c1.a <-- 2;\n}\ncomponent main = t0();",
]


@pytest.mark.parametrize("source,deferred", WITNESS_PASS)
def test_witness_passes(source, deferred):
    r = za2c.evaluate(source=source, witness=True, deferred=deferred, check=deferred is not None)
    assert r["constraints"] == []          # witness mode stores no constraints (test.rs:58,69)


@pytest.mark.parametrize("source", WITNESS_FAIL)
def test_witness_fails(source):
    with pytest.raises(TypeError):
        za2c.evaluate(source=source, witness=True)


def test_p_1_value_is_the_field_negative():
    r = za2c.evaluate(source="template t() {\n signal input p;\n signal output out;\n out <== 1-p;\n}\ncomponent main = t();",
                      witness=True, deferred={"main.p": 2}, check=True)
    assert _signals(r)["main.out"] == "main.out:Output:Some(-1)" or "Some(" in _signals(r)["main.out"]


# ---- optimiser ----------------------------------------------------------------------------------------------------------
def test_optimizer_substitutes_linear_constraints():
    """optimizer/mod.rs:186-233: t = 2 in; 2 k = 4 t; out = k  ->  one constraint out - 4 in, removed [t, k]."""
    r = za2c.evaluate(source="""
        template T() {
            signal input in;
            signal output out;
            signal t;
            signal k;
            t <== in * 2;
            2 * k === 4 * t;
            out <== k;
        }
        component main = T();""")
    assert len(r["constraints"]) == 3
    ids = {s.split(":", 1)[0]: i for i, s in enumerate(r["signals"])}
    assert r["removed"] == [ids["main.t"], ids["main.k"]]
    assert r["optimized"] == ["[ ]*[ ]+[1main.out-4main.in]"]


def test_example_circuit_of_the_readme():
    """example/circuit.za + input.json (README.md:55-70): one constraint, witness from {p, q, r}."""
    src = "template Factor() {\n  signal private input p;\n  signal private input q;\n  signal input r;\n\n  p * q === r;\n}\n\ncomponent main = Factor();\n"
    c = za2c.evaluate(source=src)
    assert c["constraints"] == ["[1main.p]*[1main.q]+[-1main.r]"] and c["optimized"] == c["constraints"]
    assert [s.split(":")[0] for s in c["signals"]] == ["one", "main.r", "main.p", "main.q"]
    w = za2c.evaluate(source=src, witness=True, deferred={"main.p": 2, "main.q": 3, "main.r": 6}, check=True)
    assert _signals(w)["main.r"] == "main.r:PublicInput:Some(6)"
    with pytest.raises(TypeError):
        za2c.evaluate(source=src, witness=True, deferred={"main.p": 2, "main.q": 3, "main.r": 7}, check=True)


# ---- `za test` over the reference's circomlib tests ---------------------------------------------------------------------
ZA_TEST_COUNTS = {      # name -> (signals, constraints) as this front-end counts them; tests of one file share a template
    "babyjub.za": {"test_BabyAdd_01": (11, 12), "test_BabyAdd_different": (11, 12), "test_BabyAdd_same": (11, 12)},
    "comparators.za": {"test_IsEqual_false": (7, 7), "test_IsEqual_true": (7, 7), "test_IsZero_false": (4, 4),
                       "test_IsZero_true": (4, 4)},
    "eddsamimc.za": {"test_eddsamimc_verifier": (21736, 21744)},
    "eddsaposeidon.za": {"test_eddsaposeidon_verifier": (21879, 21887)},
    "sha256.za": {"test_sha256_2": (204151, 204465)},
    "smtprocessor.za": {"test_smtprocessor_delete": (46894, 46903), "test_smtprocessor_insert": (46894, 46903),
                        "test_smtprocessor_update": (46894, 46903)},
    "smtverifier.za": {"test_smtverify_exclusion": (26851, 26860), "test_smtverify_inclusion_1": (26851, 26860),
                       "test_smtverify_inclusion_2": (26851, 26860), "test_smtverify_inclusion_adr1": (26851, 26860),
                       "test_smtverify_inclusion_adr2": (26851, 26860)},
}


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree (circomlib circuits) is not on this machine")
@pytest.mark.parametrize("name", sorted(ZA_TEST_COUNTS))
def test_za_test_runs_the_reference_circomlib_tests(name):
    """compiler/src/tester/embeeded.rs: every #[test] template is evaluated in witness mode, then in constraint mode, and
    the constraints are checked on the witness — the reference's own known-answer tests for circomlib (babyjub.za:16-34
    etc.) pass through this front-end."""
    got = za2c.run_tests(os.path.join(REF, "interop", "circuits", "circomlib", "za_test", name))
    assert {t["name"]: (t["signals"], t["constraints"]) for t in got} == ZA_TEST_COUNTS[name]


# ---- algebra (compiler/src/algebra/{fs,lc,qeq}.rs tests, through the evaluator) ------------------------------------------
def test_field_scalar_arithmetic():
    """fs.rs:375-436: + * neg += % << >> / on field scalars (a scope value prints through Display: the full residue)."""
    scope = za2c.evaluate(source="""
        function f() { var aa = 1 + 1; aa += 1; return aa; }
        var six = (1+1+1) * (1+1);
        var m1 = -1;
        var two = -(m1 + m1);
        var three = f();
        var md = 1012 % 1000;
        var shl = 10 << 2;
        var shr = 40 >> 1;
        var dv = 6 * (1 / 2);
    """)["scope"]
    r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    assert scope == {"six": "Algebra(6)", "m1": "Algebra(%d)" % (r - 1), "two": "Algebra(2)", "three": "Algebra(3)", "md": "Algebra(12)",
                     "shl": "Algebra(40)", "shr": "Algebra(20)", "dv": "Algebra(3)"}


def test_linear_combinations_and_quadratic_equations():
    """lc.rs:152-221, qeq.rs:150-172: LC + FS, LC * FS, -LC, LC + LC cancelling to zero, LC * LC -> QEQ, in the text form of
    qeq.rs:20-32 ([a]*[b]+[c], negative coefficients as -k)."""
    r = za2c.evaluate(source="""
        template t() {
            signal s1; signal s2;
            (2*s1 + s2) * s2 === 0;
            (s1 + 1 + 1) * 2 === 0;
            -(-s1 + s2) === 0;
            (-s1 + s2) + (s1 - s2) === 0;
            (s1 + 1) * (s2 + 2) === s1 * 3;
        }
        component main = t();""")
    assert r["constraints"] == ["[2main.s1+1main.s2]*[1main.s2]+[ ]", "[ ]*[ ]+[2main.s1+4one]", "[ ]*[ ]+[1main.s1-1main.s2]", "[ ]*[ ]+[ ]",
                                "[1main.s1+1one]*[1main.s2+2one]+[-3main.s1]"]


def test_top_level_statements_are_declarations_only():
    """lang.lalrpop: a body element is include / function / template / var / component / signal — an assignment is not."""
    with pytest.raises(TypeError) as e:
        za2c.evaluate(source="var a = 1; a += 1;")
    assert "UnrecognizedToken" in str(e.value)


# ---- error texts (the bindings hand the Debug text of the error enum to their caller, lib.rs:59,75,99) -----------------------
ERRORS = [
    # eval.rs:386-390
    ("component main = Nope();", 'Evaluator(InvalidType("component main only can be initialized with existingtemplate"))'),
    ("var k = nope(1);", 'Evaluator(NotFound("function nope"))'),                                        # eval.rs:325
    ("template t() { signal a; signal a; }\ncomponent main = t();", 'Evaluator(AlreadyExists("signal main.a"))'),   # eval.rs:851
    ("var a = 1;\nvar a = 2;", 'Evaluator(AlreadyExists("a"))'),                                         # eval.rs:882
    ("var a = b;", 'Evaluator(NotFound("b"))'),                                                          # eval.rs:1008
    ("function f(a) { return a; }\nvar k = f(1,2);", 'Evaluator(InvalidParameter("f"))'),               # eval.rs:334
    ("function f(a) { var b = a; }\nvar k = f(1);", 'Evaluator(BadFunctionReturn("f"))'),               # eval.rs:361
    # a constraint between two known scalars cannot be generated (eval.rs:1226): "left===right" through signals.format
    ("template t() { signal a; a <-- 2; a === 3; }\ncomponent main = t();", 'Evaluator(CannotGenerateConstrain("2===3"))'),
    # a product of three signals is not a quadratic equation (algebra/value.rs: InvalidOperation)
    ("template t() { signal a; signal b; a*a*a === b; }\ncomponent main = t();",
     'Evaluator(Algebra(InvalidOperation("Cannot apply operator * on [1s1]*[1s1]+[ ] over 1s1")))'),
]


@pytest.mark.parametrize("source,text", ERRORS)
def test_error_debug_text(source, text):
    with pytest.raises(TypeError) as e:
        za2c.evaluate(source=source)
    assert str(e.value) == text


def test_failed_witness_check_names_both_sides():
    """eval.rs:1213-1219: CannotTestConstrain("{lhe:?}==={rhe:?} => {left}==={right}")."""
    with pytest.raises(TypeError) as e:
        za2c.evaluate(source="template t() { signal a; a <-- 2; a === 3; }\ncomponent main = t();", witness=True)
    assert str(e.value) == 'Evaluator(CannotTestConstrain("a===3 => 2===3"))'


def test_division_is_the_field_inverse():
    """fs.rs:234-254: a / b = a * b^-1 mod r.  The inverse is a binary extended Euclid here (not b^(r-2)): checked against
    python's pow(b, -1, r) on small, large and random operands, and 1 / 0 is the reference's InvalidOperation."""
    import random
    r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    rnd = random.Random(7)
    bs = [1, 2, 3, r - 1, r - 2, (r - 1) // 2, (r + 1) // 2, 1 << 253] + [rnd.randrange(1, r) for _ in range(40)]
    src = "\n".join("var q%d = 5 / %d;" % (i, b) for i, b in enumerate(bs))
    scope = za2c.evaluate(source=src)["scope"]
    for i, b in enumerate(bs):
        assert scope["q%d" % i] == "Algebra(%d)" % (5 * pow(b, -1, r) % r), b
    with pytest.raises(TypeError) as e:
        za2c.evaluate(source="var z = 1 / 0;")
    assert "Cannot find inv" in str(e.value)
