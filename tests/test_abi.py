"""CPU: the C-ABI library loads and exports every symbol include/za_b200.h declares; without a GPU every
compute entry point fails loudly (no fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "za_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(za_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from za_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_library_exports_every_declared_symbol():
    from za_b200 import _lib
    L = ctypes.CDLL(_lib.SO_PATH)
    for name in declared_symbols():
        assert hasattr(L, name), f"libza_b200.so does not export {name}"
    assert _lib.lib().za_version() >= 1


def test_no_cpu_fallback_without_device():
    import za_b200
    from za_b200 import _lib
    if _lib.lib().za_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.Context(0)
    assert e.value.code == -1 and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under za_b200/ or include/ may reference it."""
    for base in ("za_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dirpath, f), errors="replace").read()
                    assert "za_oracle" not in txt and "tests.oracle" not in txt and "libza_oracle" not in txt, (dirpath, f)


def test_proof_json_format():
    """format.rs:80-128: compact JSON, a/b/c hex coordinates, decimal public inputs; buffer rule of binding/c lib.rs:23."""
    import json
    import za_b200
    from za_b200 import _lib
    proof = bytes(range(256))
    js = za_b200.proof_to_json(proof, [21, 0])
    d = json.loads(js)
    assert list(d) == ["a", "b", "c", "public_inputs"] and " " not in js
    assert d["a"][0] == "0x" + proof[31::-1].hex() and d["b"][1][1] == "0x" + proof[191:159:-1].hex()
    assert d["public_inputs"] == ["21", "0"]
    buf = ctypes.create_string_buffer(len(js))          # len >= size -> too small
    import numpy as np
    p = np.frombuffer(proof, np.uint8); pi = np.zeros(64, np.uint8); pi[0] = 21
    rc = _lib.lib().za_proof_to_json(p.ctypes.data_as(ctypes.c_void_p), pi.ctypes.data_as(ctypes.c_void_p), 2, buf, len(js))
    assert rc == -9
