"""ctypes mirror of libza2c — the outer ABI of za's bindings (include/za2c.h):
verbose / setup / prove / verify with the calling convention of /root/reference/binding/python3/native/src/lib.rs:9-60
(strings in, string or bool out, TypeError carrying the Debug text of the error), plus the front-end seam
(parse / evaluate / `za test`) that needs no GPU."""
import ctypes
import json
import os

from . import build as _build

ERR_NONE, ERR_BUFFER_TOO_SMALL, ERR_VERIFICATION_FAILED, ERR_CUSTOM = 0, 1, 2, 100
_L = None


def lib():
    global _L
    if _L is None:
        if not os.path.exists(_build.OUT_ZA2C):       # like libza_b200.so: built by __graft_entry__.build(), never silently rebuilt on the GPU box
            _build.build()
        L = ctypes.CDLL(_build.OUT_ZA2C)
        cp, sz, ci = ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int
        L.verbose.argtypes, L.verbose.restype = [ci], None
        L.setup.argtypes, L.setup.restype = [cp, cp, cp, cp, sz, cp, sz], ci
        L.prove.argtypes, L.prove.restype = [cp, cp, cp, sz, cp, sz], ci
        L.verify.argtypes, L.verify.restype = [cp, cp, cp, sz], ci
        L.za2c_release.argtypes, L.za2c_release.restype = [], None
        L.za2c_parse.argtypes, L.za2c_parse.restype = [ci, cp, cp, sz, cp, sz], ci
        L.za2c_eval.argtypes, L.za2c_eval.restype = [ci, cp, cp, cp, ci, cp, sz, cp, sz], ci
        L.za2c_test.argtypes, L.za2c_test.restype = [cp, cp, cp, sz, cp, sz], ci
        _L = L
    return _L


def _b(s):
    return None if s is None else (s if isinstance(s, bytes) else str(s).encode())


def _call(fn, *args, out_size=1 << 16):
    """Caller-allocated output buffer that grows on ERR_BUFFER_TOO_SMALL (what the Go binding does with its fixed buffers)."""
    err = ctypes.create_string_buffer(1 << 16)
    while True:
        out = ctypes.create_string_buffer(out_size)
        rc = fn(*args, out, out_size, err, len(err))
        if rc == ERR_BUFFER_TOO_SMALL and out_size < (1 << 30):
            out_size *= 8
            continue
        if rc != ERR_NONE:
            raise TypeError(err.value.decode(errors="replace") if rc == ERR_CUSTOM else f"libza2c error {rc}")
        return out.value.decode()


def verbose(on):
    lib().verbose(1 if on else 0)
    return bool(on)


def setup(circuit_path, pk_path, verifier_type):
    if verifier_type not in ("json", "solidity"):
        raise TypeError("invalid verifier type")
    return _call(lib().setup, _b(circuit_path), _b(pk_path), _b(verifier_type))


def prove(pk_path, inputs):
    return _call(lib().prove, _b(pk_path), _b(inputs))


def verify(verifying_key, proof_with_inputs):
    err = ctypes.create_string_buffer(1 << 16)
    rc = lib().verify(_b(verifying_key), _b(proof_with_inputs), err, len(err))
    if rc == ERR_NONE:
        return True
    if rc == ERR_VERIFICATION_FAILED:
        return False
    raise TypeError(err.value.decode(errors="replace"))


def release():
    lib().za2c_release()


# ---- front-end seam -----------------------------------------------------------------------------------------------------
EXPRESSION, STATEMENT, BODY_ELEMENT, BODY_BINCODE, PREPROCESS = 0, 1, 2, 3, 4


def parse(what, text):
    return _call(lib().za2c_parse, what, _b(text))


def evaluate(source=None, file_path=None, witness=False, deferred=None, check=False):
    d = None if deferred is None else json.dumps({k: str(v) for k, v in deferred.items()})
    return json.loads(_call(lib().za2c_eval, 2 if witness else 1, _b(source), _b(file_path), _b(d), 1 if check else 0, out_size=1 << 20))


def run_tests(file_path, prefix=""):
    return json.loads(_call(lib().za2c_test, _b(file_path), _b(prefix)))
