"""CPU: proving.key container (format.rs:223-293) and CircomCircuit::synthesize (prover.rs:45-103) mirrors,
against an independent python bincode writer and the reference's documented layout."""
import struct

import numpy as np
import pytest

import za_b200
from za_b200 import format as F
from tests import oracle as O
from tests import pyref as P

R = P.R_MOD


def bincode_fs(x):
    digits = []
    while x:
        digits.append(x & 0xFFFFFFFF); x >>= 32
    return struct.pack("<Q", len(digits)) + b"".join(struct.pack("<I", d) for d in digits)


def bincode_lc(terms):
    return struct.pack("<Q", len(terms)) + b"".join(struct.pack("<Q", s) + bincode_fs(c) for s, c in terms)


def container(ast, qeqs, ignore, params):
    out = struct.pack(">I", len(ast)) + ast + struct.pack(">I", len(qeqs))
    for a, b, c in qeqs:
        q = bincode_lc(a) + bincode_lc(b) + bincode_lc(c)
        out += struct.pack(">I", len(q)) + q
    out += struct.pack(">I", len(ignore)) + b"".join(struct.pack(">I", s) for s in ignore)
    return out + params


# example/circuit.za after za's evaluator: signals one, main.r (public input), main.p, main.q; p*q === r is a*b + c = 0
EXAMPLE_QEQS = [([(2, 1)], [(3, 1)], [(1, R - 1)])]


def test_read_pk_parses_an_independently_written_container():
    ast = b"\x02\x00\x00\x00\x00\x00\x00\x00" + b"opaque-ast-bytes"
    qeqs = EXAMPLE_QEQS + [([(0, 5), (2, R - 3)], [], [(3, 2 ** 200 + 7), (1, 1)])]
    params = b"PARAMS-BYTES" * 3
    blob = container(ast, qeqs, [4, 9], params)
    pk = F.read_pk(blob)
    assert pk.ast == ast and pk.params == params and list(pk.ignore_signals) == [4, 9] and pk.num_constraints == 2
    assert list(pk.ptr[0]) == [0, 1, 3] and list(pk.ptr[1]) == [0, 1, 1] and list(pk.ptr[2]) == [0, 1, 3]
    assert list(pk.sig[0]) == [2, 0, 2] and list(pk.sig[2]) == [1, 3, 1]
    assert O.np_to_frs(pk.coeff[0]) == [1, 5, R - 3] and O.np_to_frs(pk.coeff[2]) == [R - 1, 2 ** 200 + 7, 1]
    # write_pk reproduces the same bytes
    assert F.write_pk(pk.ptr, pk.sig, pk.coeff, pk.ignore_signals, pk.params, ast=pk.ast) == blob
    for cut in (3, 30, len(blob) - len(params) - 1):
        with pytest.raises(za_b200.ZaError):
            F.read_pk(blob[:cut])


def test_synthesize_example_circuit_matches_survey_layout():
    pk = F.read_pk(container(F.EMPTY_AST, EXAMPLE_QEQS, [], b""))
    values = O.frs_to_np([1, 6, 2, 3]).reshape(4, 32)
    s = F.synthesize(4, [0, 1, 0, 0], [], pk.ptr, pk.sig, pk.coeff, values)
    # inputs = [1, 6], aux = [2, 3]; A = [p], B = [q], C = -c = [r]
    assert (s["num_inputs"], s["num_aux"]) == (2, 2)
    assert O.np_to_frs(s["inputs"]) == [1, 6] and O.np_to_frs(s["aux"]) == [2, 3]
    assert list(s["var"][0]) == [0x80000000] and list(s["var"][1]) == [0x80000001] and list(s["var"][2]) == [1]
    assert O.np_to_frs(s["coeff"][2]) == [1]
    # and it is the circuit the golden proof was computed for
    cs = P.example_factor_circuit()
    ocs = O.CS(s["num_inputs"], s["num_aux"], s["ptr"], s["var"], s["coeff"])
    ref = O.CS.from_rows(cs.num_inputs, cs.num_aux, cs.rows)
    for w in range(3):
        assert np.array_equal(ocs.var[w], ref.var[w]) and np.array_equal(ocs.coeff[w], ref.coeff[w])


def test_synthesize_skips_ignored_signals_and_rejects_their_use():
    # signals: 0 one, 1 main.out (public), 2 main.x (private), 3 t (ignored by the optimiser), 4 u
    qeqs = [([(2, 1)], [(4, 1)], [(1, R - 1)])]
    pk = F.read_pk(container(F.EMPTY_AST, qeqs, [3], b""))
    s = F.synthesize(5, [0, 1, 0, 0, 0], pk.ignore_signals, pk.ptr, pk.sig, pk.coeff)
    assert list(s["var_of_signal"]) == [0, 1, 0x80000000, 0xFFFFFFFF, 0x80000001]
    bad = F.read_pk(container(F.EMPTY_AST, [([(3, 1)], [], [])], [3], b""))
    with pytest.raises(za_b200.ZaError):                      # format.rs:215-217 "signal {} not defined"
        F.synthesize(5, [0, 1, 0, 0, 0], bad.ignore_signals, bad.ptr, bad.sig, bad.coeff)
