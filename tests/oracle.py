"""ctypes binding of the CPU oracle (oracle/libza_oracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package za_b200 never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(_ORACLE_DIR, "libza_oracle.so")

AUX = 0x80000000
u8p = ctypes.POINTER(ctypes.c_uint8)
u32p = ctypes.POINTER(ctypes.c_uint32)


class OraR1CS(ctypes.Structure):
    _fields_ = [("num_inputs", ctypes.c_uint32), ("num_aux", ctypes.c_uint32), ("num_constraints", ctypes.c_uint32),
                ("ptr", u32p * 3), ("var", u32p * 3), ("coeff", u8p * 3)]


class OraTrace(ctypes.Structure):
    _fields_ = [(n, u8p) for n in ("a_eval", "b_eval", "c_eval", "h_coeffs", "msm_g1", "msm_g2",
                                   "a_aux_density", "b_input_density", "b_aux_density")]


def build(force=False):
    src = [os.path.join(_ORACLE_DIR, f) for f in ("za_oracle.c", "za_oracle.h", "field.h", "curve_tmpl.h")]
    if force or not os.path.exists(_SO) or (all(os.path.exists(s) for s in src) and
                                            os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in src)):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "libza_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.ora_generate_parameters.restype = ctypes.c_void_p
        L.ora_params_read.restype = ctypes.c_void_p
        L.ora_params_synthetic.restype = ctypes.c_void_p
        L.ora_params_synthetic.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.ora_params_size.restype = ctypes.c_size_t
        L.ora_params_size.argtypes = [ctypes.c_void_p]
        L.ora_params_write.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ora_params_free.argtypes = [ctypes.c_void_p]
        L.ora_params_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ora_params_read.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        L.ora_generate_parameters.argtypes = [ctypes.c_void_p] + [ctypes.c_char_p] * 7 + [ctypes.c_int, ctypes.c_void_p]
        L.ora_create_proof.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ora_verify_proof.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.ora_fft.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.ora_h_poly.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        for g in ("g1", "g2"):
            getattr(L, f"ora_multiexp_{g}").argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
            getattr(L, f"ora_{g}_multiples").argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        _lib = L
    return _lib


# ---------------------------------------------------------------- encodings
def fr_bytes(x): return int(x).to_bytes(32, "little")
def fr_int(b): return int.from_bytes(bytes(b), "little")
def frs_to_np(xs): return np.frombuffer(b"".join(fr_bytes(x) for x in xs), dtype=np.uint8).copy() if len(xs) else np.zeros(0, np.uint8)
def np_to_frs(a): b = bytes(a); return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]
def g1_bytes(p): return b"\0" * 64 if p is None else fr_bytes(p[0]) + fr_bytes(p[1])
def g1_tuple(b): b = bytes(b); return None if b == b"\0" * 64 else (fr_int(b[:32]), fr_int(b[32:64]))
def g2_bytes(p): return b"\0" * 128 if p is None else fr_bytes(p[0][0]) + fr_bytes(p[0][1]) + fr_bytes(p[1][0]) + fr_bytes(p[1][1])
def g2_tuple(b):
    b = bytes(b)
    return None if b == b"\0" * 128 else ((fr_int(b[:32]), fr_int(b[32:64])), (fr_int(b[64:96]), fr_int(b[96:128])))


def _ptr(a): return a.ctypes.data_as(ctypes.c_void_p)


def random_frs(n, seed):
    """n uniform canonical Fr elements as an (n, 32) uint8 array (rejection sampling on 254 bits)."""
    from tests.pyref import R_MOD  # noqa
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 4), dtype=np.uint64)
    todo = np.arange(n)
    mod = [(R_MOD >> (64 * i)) & (2 ** 64 - 1) for i in range(4)]
    while len(todo):
        v = rng.integers(0, 2 ** 64, size=(len(todo), 4), dtype=np.uint64)
        v[:, 3] &= np.uint64((1 << 62) - 1)
        lt = np.zeros(len(todo), dtype=bool)
        decided = np.zeros(len(todo), dtype=bool)
        for i in (3, 2, 1, 0):
            m = np.uint64(mod[i])
            lt |= (~decided) & (v[:, i] < m)
            decided |= v[:, i] != m
        out[todo[lt]] = v[lt]
        todo = todo[~lt]
    return out.view(np.uint8).reshape(n, 32)


# ---------------------------------------------------------------- field / curve
def field_op(field, op, a, b=0):
    out = ctypes.create_string_buffer(32)
    lib().ora_field_op(field, op, fr_bytes(a), fr_bytes(b), out)
    return fr_int(out.raw)


def fq2_op(op, a, b=(0, 0)):
    out = ctypes.create_string_buffer(64)
    lib().ora_fq2_op(op, fr_bytes(a[0]) + fr_bytes(a[1]), fr_bytes(b[0]) + fr_bytes(b[1]), out)
    return (fr_int(out.raw[:32]), fr_int(out.raw[32:]))


def g1_mul(p, k):
    out = ctypes.create_string_buffer(64); lib().ora_g1_mul(g1_bytes(p), fr_bytes(k), out); return g1_tuple(out.raw)
def g1_add(p, q):
    out = ctypes.create_string_buffer(64); lib().ora_g1_add(g1_bytes(p), g1_bytes(q), out); return g1_tuple(out.raw)
def g2_mul(p, k):
    out = ctypes.create_string_buffer(128); lib().ora_g2_mul(g2_bytes(p), fr_bytes(k), out); return g2_tuple(out.raw)
def g2_add(p, q):
    out = ctypes.create_string_buffer(128); lib().ora_g2_add(g2_bytes(p), g2_bytes(q), out); return g2_tuple(out.raw)


def g1_generator():
    out = ctypes.create_string_buffer(64); lib().ora_g1_generator(out); return out.raw
def g2_generator():
    out = ctypes.create_string_buffer(128); lib().ora_g2_generator(out); return out.raw


def g1_multiples(n, base=None):
    """(n, 64) uint8: (i+1)*base"""
    out = np.zeros((n, 64), np.uint8); lib().ora_g1_multiples(base or g1_generator(), n, _ptr(out)); return out
def g2_multiples(n, base=None):
    out = np.zeros((n, 128), np.uint8); lib().ora_g2_multiples(base or g2_generator(), n, _ptr(out)); return out


def pairing(p, q):
    out = ctypes.create_string_buffer(384); lib().ora_pairing(g1_bytes(p), g2_bytes(q), out); return out.raw


# ---------------------------------------------------------------- domain / multiexp
def fft(data, log_n, mode, threads=1):
    """data: (n,32) uint8 canonical; returns new array.  mode 0 fft 1 ifft 2 coset_fft 3 icoset_fft"""
    d = np.ascontiguousarray(data, dtype=np.uint8).copy()
    rc = lib().ora_fft(_ptr(d), log_n, mode, threads)
    assert rc == 0, rc
    return d


def h_poly(a, b, c, threads=1, checkpoints=False):
    n = len(a); m = 1
    while m < n: m *= 2
    out = np.zeros((m - 1, 32), np.uint8)
    ck = np.zeros((8, m, 32), np.uint8) if checkpoints else None
    rc = lib().ora_h_poly(_ptr(np.ascontiguousarray(a)), _ptr(np.ascontiguousarray(b)), _ptr(np.ascontiguousarray(c)), n,
                          _ptr(out), _ptr(ck) if checkpoints else None, threads)
    assert rc == 0, rc
    return (out, ck) if checkpoints else out


def multiexp(group, bases, scalars, density=None, threads=1):
    """bases (nb, 64|128) uint8, scalars (ne, 32) uint8, density (ne,) uint8 or None -> (rc, point bytes)"""
    sz = 64 if group == "g1" else 128
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, sz)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8).reshape(-1, 32)
    out = ctypes.create_string_buffer(sz)
    d = None if density is None else np.ascontiguousarray(density, dtype=np.uint8)
    rc = getattr(lib(), f"ora_multiexp_{group}")(_ptr(bases), len(bases), _ptr(scalars), len(scalars),
                                                 None if d is None else _ptr(d), out, threads)
    return rc, out.raw


# ---------------------------------------------------------------- Groth16
class CS:
    """CSR form of the rows bellman's enforce() receives.  Build from python rows
    [(A_terms, B_terms, C_terms)] with terms [(coeff_int, var)], or from arrays."""

    def __init__(self, num_inputs, num_aux, ptr, var, coeff):
        self.num_inputs, self.num_aux = num_inputs, num_aux
        self.ptr = [np.ascontiguousarray(p, dtype=np.uint32) for p in ptr]
        self.var = [np.ascontiguousarray(v, dtype=np.uint32) for v in var]
        self.coeff = [np.ascontiguousarray(c, dtype=np.uint8).reshape(-1, 32) for c in coeff]
        self.num_constraints = len(self.ptr[0]) - 1

    @classmethod
    def from_rows(cls, num_inputs, num_aux, rows):
        ptr, var, coeff = [], [], []
        for w in range(3):
            p, v, c = [0], [], []
            for row in rows:
                for co, va in row[w]:
                    v.append(va); c.append(fr_bytes(co))
                p.append(len(v))
            ptr.append(np.array(p, np.uint32)); var.append(np.array(v, np.uint32))
            coeff.append(np.frombuffer(b"".join(c), np.uint8).reshape(-1, 32) if c else np.zeros((0, 32), np.uint8))
        return cls(num_inputs, num_aux, ptr, var, coeff)

    def c_struct(self):
        s = OraR1CS()
        s.num_inputs, s.num_aux, s.num_constraints = self.num_inputs, self.num_aux, self.num_constraints
        for w in range(3):
            s.ptr[w] = self.ptr[w].ctypes.data_as(u32p)
            s.var[w] = self.var[w].ctypes.data_as(u32p)
            s.coeff[w] = self.coeff[w].ctypes.data_as(u8p)
        return s


class Params:
    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_params_free(self.h); self.h = None

    @classmethod
    def generate(cls, cs, toxic, g1=None, g2=None, threads=1):
        err = ctypes.c_int(0)
        st = cs.c_struct()
        h = lib().ora_generate_parameters(ctypes.byref(st), *[fr_bytes(t) for t in toxic], g1 or g1_generator(),
                                          g2 or g2_generator(), threads, ctypes.byref(err))
        if not h:
            raise RuntimeError(f"ora_generate_parameters failed: {err.value}")
        return cls(h)

    @classmethod
    def read(cls, data, checked=True):
        err = ctypes.c_int(0)
        buf = np.frombuffer(data, np.uint8)
        h = lib().ora_params_read(_ptr(buf), len(buf), 1 if checked else 0, ctypes.byref(err))
        if not h:
            raise RuntimeError(f"ora_params_read failed: {err.value}")
        return cls(h)

    @classmethod
    def synthetic(cls, ic, h, l, a, b_g1, b_g2, threads=1):
        c = (ctypes.c_uint32 * 6)(ic, h, l, a, b_g1, b_g2)
        return cls(lib().ora_params_synthetic(c, threads))

    def write(self):
        n = lib().ora_params_size(self.h)
        buf = np.zeros(n, np.uint8); lib().ora_params_write(self.h, _ptr(buf)); return buf.tobytes()

    def counts(self):
        c = (ctypes.c_uint32 * 6)(); lib().ora_params_counts(self.h, c)
        return dict(zip(("ic", "h", "l", "a", "b_g1", "b_g2"), list(c)))

    def create_proof(self, cs, inputs, aux, r, s, threads=1, trace=False):
        """inputs/aux: (n,32) uint8 canonical (inputs[0] = 1).  Returns (rc, proof 256 bytes[, trace dict])."""
        inputs = np.ascontiguousarray(inputs, np.uint8).reshape(-1, 32); aux = np.ascontiguousarray(aux, np.uint8).reshape(-1, 32)
        proof = ctypes.create_string_buffer(256)
        st = cs.c_struct()
        tr = None; bufs = {}
        if trace:
            n = cs.num_constraints + cs.num_inputs; m = 1
            while m < n: m *= 2
            bufs = dict(a_eval=np.zeros((n, 32), np.uint8), b_eval=np.zeros((n, 32), np.uint8), c_eval=np.zeros((n, 32), np.uint8),
                        h_coeffs=np.zeros((m - 1, 32), np.uint8), msm_g1=np.zeros((7, 64), np.uint8), msm_g2=np.zeros((2, 128), np.uint8),
                        a_aux_density=np.zeros(cs.num_aux, np.uint8), b_input_density=np.zeros(cs.num_inputs, np.uint8),
                        b_aux_density=np.zeros(cs.num_aux, np.uint8))
            tr = OraTrace(**{k: v.ctypes.data_as(u8p) for k, v in bufs.items()})
        rc = lib().ora_create_proof(self.h, ctypes.byref(st), _ptr(inputs), _ptr(aux), fr_bytes(r), fr_bytes(s), proof,
                                    ctypes.byref(tr) if tr is not None else None, threads)
        return (rc, proof.raw, bufs) if trace else (rc, proof.raw)

    def verify(self, proof, public_inputs):
        pi = np.frombuffer(b"".join(fr_bytes(x) for x in public_inputs), np.uint8) if public_inputs else np.zeros(0, np.uint8)
        return lib().ora_verify_proof(self.h, proof, _ptr(pi) if len(pi) else None, len(public_inputs))


def proof_tuple(proof):
    return (g1_tuple(proof[:64]), g2_tuple(proof[64:192]), g1_tuple(proof[192:256]))
