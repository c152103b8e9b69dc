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


# ---------------------------------------------------------------- libza2c: the outer C ABI of za's bindings (include/za2c.h)
def _za2c():
    from za_b200 import build as zb
    L = ctypes.CDLL(zb.OUT_ZA2C)
    L.verify.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    L.verify.restype = ctypes.c_int
    L.prove.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    L.prove.restype = ctypes.c_int
    L.setup.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    L.setup.restype = ctypes.c_int
    L.verbose.argtypes = [ctypes.c_int]
    L.verbose.restype = None
    return L


def test_za2c_exports_the_reference_symbols():
    """binding/go/lib.go:6-9 links `verbose`, `setup`, `prove`, `verify` from -lza2c."""
    text = open(os.path.join(ROOT, "include", "za2c.h")).read()
    L = _za2c()
    for name in ("verbose", "setup", "prove", "verify"):
        assert re.search(r"\b%s\s*\(" % name, text) and hasattr(L, name)


def test_za2c_verify_like_the_bindings_use_it():
    """binding/python3/test/test.py:22-28 / binding/go/test/test.go:37-50: verify(vk_json, proof_json) == true; return codes
    of binding/c/native/src/lib.rs:10-13 and the `len >= size` buffer rule of lib.rs:23."""
    import json
    import za_b200
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "groth16_example.json")))
    vk = {k: bytes.fromhex(v) for k, v in g["vk"].items() if k != "ic"}
    vk["ic"] = [bytes.fromhex(x) for x in g["vk"]["ic"]]
    vk_json = za_b200.vk_to_json(vk, ["main.r"]).encode()
    L = _za2c()
    L.verbose(0)
    err = ctypes.create_string_buffer(512)
    assert L.verify(vk_json, g["proof_json"].encode(), err, 512) == 0
    tampered = json.loads(g["proof_json"]); tampered["public_inputs"] = ["7"]
    assert L.verify(vk_json, json.dumps(tampered).encode(), err, 512) == 2          # ERR_VERIFICATION_FAILED
    assert L.verify(vk_json, b'{"a":1}', err, 512) == 100 and len(err.value) > 0    # ERR_CUSTOM, text in err_buf
    assert L.verify(vk_json, b'{"a":1}', err, 4) == 1                               # ERR_BUFFER_TOO_SMALL


def test_za2c_setup_and_prove_argument_rules_and_no_cpu_fallback(tmp_path):
    """lib.rs:51-101: verifier type checked first ("invalid validator type"), errors as Debug text in err_buf, `len >= size`
    is "too small".  Without a CUDA device the GPU stages fail loudly — there is no CPU path behind setup / prove."""
    import torch
    L = _za2c()
    err = ctypes.create_string_buffer(1024)
    out = ctypes.create_string_buffer(64)
    assert L.setup(b"circuit.za", b"proving.key", b"yaml", out, 64, err, 1024) == 100 and err.value == b"invalid validator type"
    assert L.setup(b"/nonexistent/circuit.za", b"proving.key", b"json", out, 64, err, 1024) == 100 and len(err.value) > 0
    assert L.prove(b"/nonexistent/proving.key", b"{}", out, 64, err, 8) == 1          # error text longer than err_buf
    assert L.prove(b"/nonexistent/proving.key", b"{", out, 64, err, 1024) == 100      # malformed inputs JSON
    if not torch.cuda.is_available():
        circuit = tmp_path / "c.za"
        circuit.write_text("template T() { signal private input p; signal output r; r <== p*p; }\ncomponent main = T();\n")
        rc = L.setup(str(circuit).encode(), str(tmp_path / "proving.key").encode(), b"json", out, 64, err, 1024)
        assert rc == 100 and b"no CPU fallback" in err.value
        assert not (tmp_path / "proving.key").exists()


def test_headers_are_plain_c_and_a_c_program_links(tmp_path):
    """The boundary is a C ABI: both headers compile as C99 (no C++, no torch types) and a C program that calls through
    them links against the two libraries and runs the host-only entry points (no GPU needed for these)."""
    import subprocess
    from za_b200 import build as zb
    src = tmp_path / "abi_smoke.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "za_b200.h"
#include "za2c.h"
int main(void) {
    char err[256];
    if (za_version() < 1) return 1;
    if (verify("{", "{", err, sizeof err) != ZA2C_ERR_CUSTOM) return 2;      /* malformed JSON -> 100, text in err */
    if (strlen(err) == 0) return 3;
    if (prove("proving.key", "{}", err, 4, err, 4) != ZA2C_ERR_BUFFER_TOO_SMALL) return 4;
    printf("devices %d\n", za_device_count());
    return 0;
}
''')
    exe = tmp_path / "abi_smoke"
    lib_dir = os.path.dirname(zb.OUT)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L" + lib_dir, "-lza2c", "-lza_b200", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.startswith("devices ")
