"""The outer ABI of za's bindings on a B200: libza2c `setup` -> `prove` -> `verify` from circuit text (SURVEY.md §8f N4,
§8b2), the way /root/reference/binding/python3/test/test.py:1-27, binding/go/test/test.go and `za setup / prove / verify`
(example/Makefile) drive the reference.  The GPU stages behind these calls (generate_parameters, Parameters::read(checked),
constraint check, create_proof) are the ones the parity tests of tests/test_gpu_parity.py pin; here the checks are
end-to-end: proofs verify, wrong statements do not, and the files / JSON the calls leave behind are readable by the
independent Python readers of za_b200/format.py."""
import json
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BINDING_TEST_CIRCUIT = """
template T() {
        signal private input p;
        signal private input q;
        signal output r;

        r <== p*q;
}
component main = T();
"""

EXAMPLE_CIRCUIT = """template Factor() {
  signal private input p;
  signal private input q;
  signal input r;

  p * q === r;
}

component main = Factor();
"""


@pytest.fixture()
def circom():
    """`import libza2py as circom` (binding/python3/test/test.py:1) with bindings/python3 on the path."""
    sys.path.insert(0, os.path.join(ROOT, "bindings", "python3"))
    try:
        import libza2py
        yield libza2py
    finally:
        sys.path.pop(0)
        from za_b200 import za2c
        za2c.release()


def test_python3_binding_test_flow(circom, tmp_path):
    """binding/python3/test/test.py:4-27, statement by statement."""
    circuit_path, pk_path = str(tmp_path / "circuit.circom"), str(tmp_path / "proving.key")
    circom.verbose(True)
    with open(circuit_path, "w") as f:
        f.write(BINDING_TEST_CIRCUIT)
    verifying_key = circom.setup(circuit_path, pk_path, "json")
    all_inputs = {"p": "2", "q": "3"}
    proof_and_public_inputs = circom.prove(pk_path, json.dumps(all_inputs))
    success = circom.verify(verifying_key, proof_and_public_inputs)
    assert success is True
    circom.verbose(False)
    # what the strings hold (format.rs:30-99): the public output r = 6 and main.r named in the verifying key
    assert json.loads(proof_and_public_inputs)["public_inputs"] == ["6"]
    assert json.loads(verifying_key)["input_names"] == ["main.r"]
    tampered = json.loads(proof_and_public_inputs); tampered["public_inputs"] = ["7"]
    assert circom.verify(verifying_key, json.dumps(tampered)) is False
    # a second proof from the loaded key: fresh r, s -> different bytes, still valid
    again = circom.prove(pk_path, json.dumps({"p": "5", "q": "7"}))
    assert json.loads(again)["public_inputs"] == ["35"] and circom.verify(verifying_key, again) is True
    with pytest.raises(TypeError):
        circom.setup(circuit_path, pk_path, "yaml")


def test_example_circuit_setup_prove_verify_and_the_key_file(circom, tmp_path):
    """example/Makefile: za setup; za prove; za verify on example/circuit.za + input.json; the proving.key container is
    the reference's (format.rs:223-293) as read by za_b200/format.py."""
    from za_b200 import format as F
    circuit_path, pk_path = str(tmp_path / "circuit.za"), str(tmp_path / "proving.key")
    open(circuit_path, "w").write(EXAMPLE_CIRCUIT)
    vk_json = circom.setup(circuit_path, pk_path, "json")
    pk = F.read_pk(open(pk_path, "rb").read())
    assert pk.num_constraints == 1 and len(pk.params) > 0
    proof = circom.prove(pk_path, '{ "p" : "2", "q":"3", "r":"6" }')
    assert json.loads(proof)["public_inputs"] == ["6"]
    assert circom.verify(vk_json, proof) is True
    with pytest.raises(TypeError) as e:                 # p*q != r: the witness evaluator's own `===` check fails first
        circom.prove(pk_path, '{ "p" : "2", "q":"3", "r":"7" }')
    assert len(str(e.value)) > 0
    with pytest.raises(TypeError):                      # missing input: "signal 'main.q' value is not defined" class
        circom.prove(pk_path, '{ "p" : "2", "r":"6" }')
    # the key stays usable after the failed calls
    assert circom.verify(vk_json, circom.prove(pk_path, '{ "p" : "3", "q":"4", "r":"12" }')) is True
    sol = circom.setup(circuit_path, pk_path, "solidity")
    assert "pragma solidity" in sol and "verifyTx" in sol and "Pairing" not in sol[:0]
    # a Solidity-type setup re-keys the file: the old verifying key no longer matches new proofs
    assert circom.verify(vk_json, circom.prove(pk_path, '{ "p" : "2", "q":"3", "r":"6" }')) is False


def _mimc_chain_source(n_links, constants):
    rounds = len(constants)
    return """
template Pow7Round(c) {
    signal input x;
    signal input k;
    signal output out;
    signal t2;
    signal t4;
    signal t6;
    var t = k + x + c;
    t2 <== t * t;
    t4 <== t2 * t2;
    t6 <== t4 * t2;
    out <== t6 * t;
}
template Mimc7(nrounds) {
    signal input x_in;
    signal input k;
    signal output out;
    var c = [%s];
    component r[nrounds];
    for (var i = 0; i < nrounds; i += 1) {
        r[i] = Pow7Round(c[i]);
        r[i].k <== k;
        if (i == 0) { r[i].x <== x_in; } else { r[i].x <== r[i-1].out; }
    }
    out <== r[nrounds-1].out + k;
}
template Chain(n) {
    signal private input seed;
    signal input key;
    signal output digest;
    component h[n];
    for (var i = 0; i < n; i += 1) {
        h[i] = Mimc7(%d);
        h[i].k <== key;
        if (i == 0) { h[i].x_in <== seed; } else { h[i].x_in <== h[i-1].out; }
    }
    digest <== h[n-1].out;
}
component main = Chain(%d);
""" % (",".join(str(c) for c in constants), rounds, n_links)


def test_compile_witness_prove_a_mimc7_chain(circom, tmp_path):
    """Config 3's shape from source text: a few thousand constraints compiled by the front-end, witness computed by the
    evaluator from two inputs, proved and verified on the GPU; the public output is checked against plain-integer MiMC7
    (tests/eddsa_circuit.py, circomlib's semantics on the reference's constants)."""
    from tests import eddsa_circuit as E
    from za_b200 import format as F
    c = E.mimc7_constants()
    n_links = 8
    circuit_path, pk_path = str(tmp_path / "chain.za"), str(tmp_path / "proving.key")
    open(circuit_path, "w").write(_mimc_chain_source(n_links, c))
    vk_json = circom.setup(circuit_path, pk_path, "json")
    pk = F.read_pk(open(pk_path, "rb").read())
    assert pk.num_constraints >= 4 * len(c) * n_links              # 4 products per round survive the optimiser
    seed, key = 1234567, 89
    x = seed
    for _ in range(n_links):
        x = E.mimc7(x, key, c)
    proof = circom.prove(pk_path, json.dumps({"seed": str(seed), "key": str(key)}))
    assert json.loads(proof)["public_inputs"] == [str(x), str(key)]     # outputs first, then public inputs (test.rs:738-769)
    assert circom.verify(vk_json, proof) is True
    assert json.loads(vk_json)["input_names"] == ["main.digest", "main.key"]
    forged = json.loads(proof); forged["public_inputs"][0] = str((x + 1) % E.R)
    assert circom.verify(vk_json, json.dumps(forged)) is False
