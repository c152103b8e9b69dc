"""CPU: the product's host-side verifier and JSON formats (no GPU needed) against the committed golden proof
(closed form, python integers) and against the oracle's prover / verifier."""
import json
import os

import numpy as np
import pytest

import za_b200
from tests import oracle as O
from tests import pyref as P

G = os.path.join(os.path.dirname(__file__), "golden")


def vk_from_params_bytes(blob):
    """bellman VerifyingKey::write layout (big-endian, G2 as c1 || c0) -> the za_pk_vk interchange layout."""
    def g1(b): return b[31::-1] + b[63:31:-1]
    def g2(b): return b[63:31:-1] + b[31::-1] + b[127:95:-1] + b[95:63:-1]
    o = 0
    alpha = g1(blob[o:o + 64]); o += 64
    beta1 = g1(blob[o:o + 64]); o += 64
    beta2 = g2(blob[o:o + 128]); o += 128
    gamma = g2(blob[o:o + 128]); o += 128
    delta1 = g1(blob[o:o + 64]); o += 64
    delta2 = g2(blob[o:o + 128]); o += 128
    n = int.from_bytes(blob[o:o + 4], "big"); o += 4
    ic = [g1(blob[o + 64 * i:o + 64 * i + 64]) for i in range(n)]
    return dict(alpha_g1=alpha, beta_g1=beta1, beta_g2=beta2, gamma_g2=gamma, delta_g1=delta1, delta_g2=delta2, ic=ic)


def golden():
    return json.load(open(os.path.join(G, "groth16_example.json")))


def test_golden_proof_verifies_and_wrong_input_fails():
    g = golden()
    vk = {k: bytes.fromhex(v) for k, v in g["vk"].items() if k != "ic"}
    vk["ic"] = [bytes.fromhex(x) for x in g["vk"]["ic"]]
    proof = bytes.fromhex(g["proof_hex"])
    assert za_b200.verify_proof(vk, proof, [6]) is True            # prover.rs:294-298
    assert za_b200.verify_proof(vk, proof, [7]) is False           # prover.rs:300-305
    with pytest.raises(za_b200.ZaError):                           # MalformedVerifyingKey
        za_b200.verify_proof(vk, proof, [])
    bad = bytearray(proof); bad[0] ^= 1
    with pytest.raises(za_b200.ZaError):                           # point off the curve: "bad coordinates"
        za_b200.verify_proof(vk, bytes(bad), [6])


def test_json_round_trip_like_helper_verify():
    g = golden()
    vk = {k: bytes.fromhex(v) for k, v in g["vk"].items() if k != "ic"}
    vk["ic"] = [bytes.fromhex(x) for x in g["vk"]["ic"]]
    vk_json = za_b200.vk_to_json(vk, ["main.r"])
    d = json.loads(vk_json)
    assert list(d) == ["alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2", "gamma_g2", "ic", "input_names"]   # format.rs:131-140
    assert d["input_names"] == ["main.r"] and len(d["ic"]) == 2 and " " not in vk_json
    assert d["alpha_g1"][0] == "0x" + vk["alpha_g1"][31::-1].hex()
    assert za_b200.verify(vk_json, g["proof_json"]) is True         # helper.rs:149-158
    tampered = json.loads(g["proof_json"]); tampered["public_inputs"] = ["7"]
    assert za_b200.verify(vk_json, json.dumps(tampered)) is False
    # decimal coordinates are accepted too (str_to_fq goes through FS::parse, format.rs:33-36)
    dec = json.loads(g["proof_json"]); dec["a"] = [str(int(x, 16)) for x in dec["a"]]
    assert za_b200.verify(vk_json, json.dumps(dec)) is True
    for broken in ('{"a":1}', "[1,2", '{"a":["0x1","0x2"],"b":[],"c":[],"public_inputs":[]}'):
        with pytest.raises(za_b200.ZaError):
            za_b200.verify(vk_json, broken)
    lead = json.loads(g["proof_json"]); lead["public_inputs"] = ["06"]      # Fr::from_str rejects leading zeros
    with pytest.raises(za_b200.ZaError):
        za_b200.verify(vk_json, json.dumps(lead))


def test_agrees_with_oracle_on_random_circuits():
    from tests import circuits
    rng = np.random.default_rng(7)
    for nc in (3, 17):
        ni, na, ptr, var, coeff, inputs, aux = circuits.mul_chain(nc, x0=int(rng.integers(2, 1000)))
        ocs = O.CS(ni, na, ptr, var, coeff)
        prm = O.Params.generate(ocs, [int(x) for x in rng.integers(2, 2 ** 62, 5)])
        rc, proof = prm.create_proof(ocs, inputs, aux, int(rng.integers(1, 2 ** 62)), int(rng.integers(1, 2 ** 62)))
        assert rc == 0
        vk = vk_from_params_bytes(prm.write())
        out = O.fr_int(inputs[1])
        assert prm.verify(proof, [out]) == 1 and za_b200.verify_proof(vk, proof, [out]) is True
        assert prm.verify(proof, [out + 1]) == 0 and za_b200.verify_proof(vk, proof, [out + 1]) is False
        js = za_b200.proof_to_json(proof, [out])
        assert za_b200.verify(za_b200.vk_to_json(vk), js) is True


def test_pairing_constants_rederived():
    q, r = P.Q_MOD, P.R_MOD
    u = 4965661367192848881
    assert 36 * u ** 4 + 36 * u ** 3 + 24 * u ** 2 + 6 * u + 1 == q and 36 * u ** 4 + 36 * u ** 3 + 18 * u ** 2 + 6 * u + 1 == r
    assert 6 * u + 2 == (1 << 64) + 0x9d797039be763ba8
    assert (q ** 12 - 1) % r == 0 and ((q ** 12 - 1) // r) >> (43 * 64) == 0x2f4b6dc970


# ---- Solidity verifier text (ethereum.rs:216-261)
def _py_solidity(vk, names, template):
    """Independent restatement of generate_solidity's substitutions in python (str.replace, python ints)."""
    def c(b): return "0x%064x" % int.from_bytes(b, "little")
    def g1(p): return c(p[0:32]) + "," + c(p[32:64])
    def g2(p): return "[" + c(p[32:64]) + "," + c(p[0:32]) + "],[" + c(p[96:128]) + "," + c(p[64:96]) + "]"
    def dbg(s): return '"' + s.replace("\\", "\\\\").replace('"', '\\"') + '"'
    t = template
    t = t.replace("<%vk_a%>", g1(vk["alpha_g1"])).replace("<%vk_b%>", g2(vk["beta_g2"])).replace("<%vk_gamma%>", g2(vk["gamma_g2"]))
    t = t.replace("<%vk_delta%>", g2(vk["delta_g2"])).replace("<%vk_inputs_length%>", str(len(names)))
    t = t.replace("<%vk_inputs%>", "[" + ", ".join(dbg(n) for n in names) + "]").replace("<%vk_gammaABC_length%>", str(len(vk["ic"])))
    return t.replace("<%vk_gammaABC_pts%>", "\n".join("vk.gammaABC[%d] = Pairing.G1Point(%s);" % (i, g1(p)) for i, p in enumerate(vk["ic"])))


def _golden_vk():
    g = golden()
    vk = {k: bytes.fromhex(v) for k, v in g["vk"].items() if k != "ic"}
    vk["ic"] = [bytes.fromhex(x) for x in g["vk"]["ic"]]
    return vk


def test_solidity_text_against_python_expectation():
    """Config 1's verifying key through za_vk_to_solidity with a caller-supplied template that uses every placeholder
    (twice, to pin replace-all), against the python restatement of ethereum.rs:216-261."""
    vk = _golden_vk()
    tmpl = ("A(<%vk_a%>) B(<%vk_b%>) G(<%vk_gamma%>) D(<%vk_delta%>) n=<%vk_inputs_length%> // <%vk_inputs%>\n"
            "len <%vk_gammaABC_length%>\n<%vk_gammaABC_pts%>\nagain <%vk_a%> <%vk_inputs_length%>\n")
    for names in (["main.r"], ["main.a", 'we"ird\\name'], []):
        vk2 = dict(vk)
        if len(names) + 1 != len(vk["ic"]):
            vk2["ic"] = (vk["ic"] * 3)[:len(names) + 1]
        assert za_b200.vk_to_solidity(vk2, names, tmpl) == _py_solidity(vk2, names, tmpl)
    # G2 is printed imaginary part first (ethereum.rs:227-238)
    text = za_b200.vk_to_solidity(vk, ["main.r"], "<%vk_b%>")
    assert text.startswith("[0x%064x,0x%064x]" % (int.from_bytes(vk["beta_g2"][32:64], "little"), int.from_bytes(vk["beta_g2"][0:32], "little")))


def test_solidity_builtin_template_is_a_complete_contract():
    vk = _golden_vk()
    text = za_b200.vk_to_solidity(vk, ["main.r"])
    assert "<%" not in text and "pragma solidity" in text and "function verifyTx(" in text
    assert "uint[1] memory input" in text and 'public inputs, in order: ["main.r"]' in text
    assert text.count("] = Bn128.G1(0x") == 2 and "new Bn128.G1[](2)" in text
    assert "0x%064x" % int.from_bytes(vk["alpha_g1"][0:32], "little") in text
    # the moduli the reference's template embeds (ethereum.rs:37, :173)
    assert str(P.Q_MOD) in text and str(P.R_MOD) in text


def test_solidity_with_the_reference_template_when_present():
    """With the reference tree at hand (this container only), its CONTRACT_TEMPLATE through the library equals the python
    restatement — the reference's exact contract for config 1's key."""
    path = "/root/reference/prover/src/groth16/ethereum.rs"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    src = open(path).read()
    start = src.index('r#"') + 3
    tmpl = src[start:src.index('"#;', start)]
    vk = _golden_vk()
    got = za_b200.vk_to_solidity(vk, ["main.r"], tmpl)
    assert got == _py_solidity(vk, ["main.r"], tmpl)
    assert "<%" not in got and "vk.gammaABC[1] = Pairing.G1Point(0x" in got
    vk_inf = dict(vk); vk_inf["alpha_g1"] = bytes(64)
    with pytest.raises(za_b200.ZaError):                           # "non-infinite point expected" (ethereum.rs:224)
        za_b200.vk_to_solidity(vk_inf, ["main.r"], tmpl)


def test_untrusted_json_cannot_exhaust_the_stack_and_trailing_bytes_are_refused():
    """Round-1 review: proof / vk JSON comes from untrusted parties.  serde_json (what the reference parses with) stops at 128
    levels of nesting and refuses trailing characters; so does this parser — an error, never a crash."""
    import json
    import za_b200
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "groth16_example.json")))
    vk = {k: bytes.fromhex(v) for k, v in g["vk"].items() if k != "ic"}
    vk["ic"] = [bytes.fromhex(x) for x in g["vk"]["ic"]]
    vk_json = za_b200.vk_to_json(vk, ["main.r"])
    assert za_b200.verify(vk_json, g["proof_json"]) is True
    for bad in ("[" * 100000, '{"a":' * 5000 + "1" + "}" * 5000):
        with pytest.raises(za_b200.ZaError) as e:
            za_b200.verify(vk_json, bad)
        assert "recursion limit" in str(e.value)
        with pytest.raises(za_b200.ZaError):
            za_b200.verify(bad, g["proof_json"])
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.verify(vk_json, g["proof_json"] + " x")
    assert "trailing" in str(e.value)
    assert za_b200.verify(vk_json, g["proof_json"] + "  \n") is True          # trailing whitespace is fine


def test_fast_final_exponentiation_chain_is_a_valid_exponent():
    """verify.cu `final_exponentiation_fast`: the addition chain in u (three exponentiations by -u, squarings, products,
    Frobenius maps, conjugations) restated over exponents modulo the order q^4 - q^2 + 1 of the cyclotomic subgroup — the
    total is a multiple of (q^4 - q^2 + 1) / r by a factor coprime to r, so `== 1` decides the same as the plain power
    (q^12 - 1) / r; and the (q^k - 1) / 6 constants of the Frobenius maps are what the source holds."""
    import re
    from math import gcd
    p, r, u = P.Q_MOD, P.R_MOD, 4965661367192848881
    n = p ** 4 - p ** 2 + 1
    assert n % r == 0 and (p ** 12 - 1) % ((p ** 6 - 1) * (p ** 2 + 1) * n) == 0
    neg_u = lambda e: (-u * e) % n
    sq = lambda e: (2 * e) % n
    frob = lambda e, k: (p ** k * e) % n
    conj = lambda e: (-e) % n
    f = 1
    y0 = neg_u(f); y1 = sq(y0); y2 = sq(y1); y3 = (y2 + y1) % n
    y4 = neg_u(y3); y5 = sq(y4); y6 = neg_u(y5)
    y3, y6 = conj(y3), conj(y6)
    y7 = (y6 + y4) % n; y8 = (y7 + y3) % n; y9 = (y8 + y1) % n; y10 = (y8 + y4) % n; y11 = (y10 + f) % n
    y13 = (frob(y9, 1) + y11) % n
    y14 = (frob(y8, 2) + y13) % n
    y15 = frob((conj(f) + y9) % n, 3)
    e = (y15 + y14) % n
    h = n // r
    assert e % h == 0 and gcd(e // h, r) == 1
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "za_b200", "csrc", "verify.cu")).read()
    for k in (1, 2, 3):
        m = re.search(r"EXP_Q%dM1_6\[\d+\] = \{([^}]*)\}" % k, src)
        limbs = [int(x.strip().rstrip("UL"), 16) for x in m.group(1).split(",")]
        assert sum(v << (64 * i) for i, v in enumerate(limbs)) == (p ** k - 1) // 6 and (p ** k - 1) % 6 == 0


def test_fast_and_plain_final_exponentiation_give_the_same_verdicts():
    """The plain square-and-multiply by (q^12 - 1) / r (ZA_VERIFY_PLAIN_EXP=1, read once per process) against the default."""
    import subprocess
    import sys
    code = r"""
import json, os, sys
sys.path.insert(0, %r)
import za_b200
g = json.load(open(%r))
vk = {k: bytes.fromhex(v) for k, v in g["vk"].items() if k != "ic"}
vk["ic"] = [bytes.fromhex(x) for x in g["vk"]["ic"]]
vkj = za_b200.vk_to_json(vk, ["main.r"])
bad = json.loads(g["proof_json"]); bad["public_inputs"] = ["7"]
swap = json.loads(g["proof_json"]); swap["a"], swap["c"] = swap["c"], swap["a"]
print(za_b200.verify(vkj, g["proof_json"]), za_b200.verify(vkj, json.dumps(bad)), za_b200.verify(vkj, json.dumps(swap)))
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "groth16_example.json"))
    outs = []
    for plain in ("0", "1"):
        env = dict(os.environ, ZA_VERIFY_PLAIN_EXP=plain)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip())
    assert outs[0] == outs[1] == "True False False"
