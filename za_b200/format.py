"""proving.key container and circuit synthesis — host-side mirror of the reference's format.rs / prover.rs.

  read_pk / write_pk   <- /root/reference/prover/src/groth16/format.rs:223-293
  synthesize           <- CircomCircuit::synthesize, /root/reference/prover/src/groth16/prover.rs:45-103
All parsing is done by libza_b200.so (za_pkfile_*, za_synthesize); this module only moves numpy buffers.
"""
import ctypes

import numpy as np

from ._lib import check, lib

EMPTY_AST = bytes(8)      # bincode(Vec::<BodyElementP>::new()): a u64 length of zero


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _pp(arrs):
    return (ctypes.c_void_p * 3)(*[a.ctypes.data_as(ctypes.c_void_p).value for a in arrs])


class ProvingKeyFile:
    """What read_pk returns (format.rs:26-31) minus the parsed AST: constraints over signal ids, the ignored
    signals, the opaque AST blob and bellman's Parameters bytes (hand those to Parameters.read)."""

    def __init__(self, ast, ptr, sig, coeff, ignore_signals, params):
        self.ast, self.ptr, self.sig, self.coeff, self.ignore_signals, self.params = ast, ptr, sig, coeff, ignore_signals, params

    @property
    def num_constraints(self):
        return len(self.ptr[0]) - 1


def read_pk(data):
    buf = np.frombuffer(bytes(data), np.uint8)
    info = (ctypes.c_uint64 * 6)()
    off, ao, al = ctypes.c_size_t(0), ctypes.c_size_t(0), ctypes.c_size_t(0)
    check(lib().za_pkfile_scan(_p(buf), buf.shape[0], info, ctypes.byref(off), ctypes.byref(ao), ctypes.byref(al)))
    nc, nig = int(info[0]), int(info[1])
    ptr = [np.zeros(nc + 1, np.uint32) for _ in range(3)]
    sig = [np.zeros(max(int(info[2 + w]), 1), np.uint32) for w in range(3)]
    coeff = [np.zeros((max(int(info[2 + w]), 1), 32), np.uint8) for w in range(3)]
    ignore = np.zeros(max(nig, 1), np.uint32)
    check(lib().za_pkfile_read(_p(buf), buf.shape[0], _pp(ptr), _pp(sig), _pp(coeff), _p(ignore)))
    sig = [s[:int(info[2 + w])] for w, s in enumerate(sig)]
    coeff = [c[:int(info[2 + w])] for w, c in enumerate(coeff)]
    b = bytes(data)
    return ProvingKeyFile(b[ao.value:ao.value + al.value], ptr, sig, coeff, ignore[:nig].copy(), b[off.value:])


def write_pk(ptr, sig, coeff, ignore_signals, params, ast=EMPTY_AST):
    ptr = [np.ascontiguousarray(p, np.uint32) for p in ptr]
    sig = [np.ascontiguousarray(s, np.uint32) if len(s) else np.zeros(1, np.uint32) for s in sig]
    coeff = [np.ascontiguousarray(c, np.uint8).reshape(-1, 32) if len(c) else np.zeros((1, 32), np.uint8) for c in coeff]
    ign = np.ascontiguousarray(ignore_signals, np.uint32) if len(ignore_signals) else np.zeros(1, np.uint32)
    astb = np.frombuffer(bytes(ast), np.uint8)
    prm = np.frombuffer(bytes(params), np.uint8)
    need = ctypes.c_size_t(0)
    nc = len(ptr[0]) - 1
    args = (_p(astb), astb.shape[0], nc, _pp(ptr), _pp(sig), _pp(coeff), _p(ign), len(ignore_signals), _p(prm), prm.shape[0])
    lib().za_pkfile_write(*args, None, 0, ctypes.byref(need))
    out = np.zeros(need.value, np.uint8)
    check(lib().za_pkfile_write(*args, _p(out), out.shape[0], ctypes.byref(need)))
    return out.tobytes()


def synthesize(n_signals, is_public, ignore_signals, ptr, sig, coeff, values=None):
    """Signals -> bellman variables and enforce rows (C negated).  Returns a dict with num_inputs, num_aux,
    var (per matrix, per term), coeff (A, B unchanged; C negated), var_of_signal and, when `values`
    ((n_signals, 32) canonical) is given, the inputs / aux witness arrays."""
    ptr = [np.ascontiguousarray(p, np.uint32) for p in ptr]
    sigs = [np.ascontiguousarray(s, np.uint32) if len(s) else np.zeros(1, np.uint32) for s in sig]
    nc = len(ptr[0]) - 1
    pub = np.ascontiguousarray(is_public, np.uint8)
    ign = np.ascontiguousarray(ignore_signals, np.uint32) if len(ignore_signals) else np.zeros(1, np.uint32)
    c_in = np.ascontiguousarray(coeff[2], np.uint8).reshape(-1, 32) if len(coeff[2]) else np.zeros((1, 32), np.uint8)
    var_of = np.zeros(n_signals, np.uint32)
    out_var = [np.zeros(max(len(s), 1), np.uint32) for s in sig]
    c_out = np.zeros_like(c_in)
    ni, na = ctypes.c_uint32(0), ctypes.c_uint32(0)
    check(lib().za_synthesize(n_signals, _p(pub), _p(ign), len(ignore_signals), nc, _pp(ptr), _pp(sigs), _p(c_in), _p(var_of), _pp(out_var),
                              _p(c_out), ctypes.byref(ni), ctypes.byref(na)))
    res = dict(num_inputs=ni.value, num_aux=na.value, ptr=ptr, var=[v[:len(s)] for v, s in zip(out_var, sig)],
               coeff=[np.ascontiguousarray(coeff[0], np.uint8).reshape(-1, 32), np.ascontiguousarray(coeff[1], np.uint8).reshape(-1, 32),
                      c_out[:len(sig[2])]], var_of_signal=var_of)
    if values is not None:
        values = np.ascontiguousarray(values, np.uint8).reshape(n_signals, 32)
        inputs = np.zeros((ni.value, 32), np.uint8)
        aux = np.zeros((na.value, 32), np.uint8)
        for s in range(n_signals):
            v = int(var_of[s])
            if v == 0xFFFFFFFF:
                continue
            if v & 0x80000000:
                aux[v & 0x7FFFFFFF] = values[s]
            else:
                inputs[v] = values[s]
        inputs[0] = 0
        inputs[0, 0] = 1
        res["inputs"], res["aux"] = inputs, aux
    return res
