"""CPU: the device limb schedule of za_b200/csrc/ff.cuh run on the host through the emulated PTX carry
flag (ZA_FF_EMULATE_PTX), and the fast 64-bit host path, both against python integers."""
import ctypes
import os
import random
import subprocess

import pytest

from tests import pyref as P

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "ff_host_shim.cpp")
A = ctypes.c_uint32 * 8
A2 = ctypes.c_uint32 * 16
R = 1 << 256


def _build(tag, defs):
    out = os.path.join(HERE, "csrc", f"libffhost_{tag}.so")
    deps = [SRC, os.path.join(HERE, "..", "za_b200", "csrc", "ff.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++"] + defs + [SRC, "-o", out])
    L = ctypes.CDLL(out)
    L.shim_lost_carries.restype = ctypes.c_uint64
    return L


@pytest.fixture(scope="module", params=["emu", "fast"])
def shim(request):
    return _build(request.param, ["-DZA_FF_EMULATE_PTX=1"] if request.param == "emu" else [])


def tol(x): return A(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
def frl(a): return sum(int(a[i]) << (32 * i) for i in range(len(a)))


def binop(L, fn, x, y):
    o = A(); getattr(L, fn)(tol(x), tol(y), o); return frl(o)


def unop(L, fn, x):
    o = A(); getattr(L, fn)(tol(x), o); return frl(o)


def test_derived_constants():
    # R, R^2, -p^-1 of ff.cuh re-derived from the moduli
    for p, one0, inv in ((P.R_MOD, 0x4ffffffb, 0xefffffff), (P.Q_MOD, 0xc58f0d9d, 0xe4866389)):
        assert (R % p) & 0xFFFFFFFF == one0
        assert (-pow(p, -1, 1 << 32)) % (1 << 32) == inv


@pytest.mark.parametrize("name,p", [("fr", P.R_MOD), ("fq", P.Q_MOD)])
def test_field_ops(shim, name, p):
    rng = random.Random(1)
    Ri = pow(R, -1, p)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, R % p, (R * R) % p, 1 << 253, p - (1 << 200)]
    vals = edge + [rng.randrange(p) for _ in range(150)]
    for x in vals:
        for y in rng.sample(vals, 8) + edge:
            assert binop(shim, f"shim_{name}_mul", x, y) == x * y * Ri % p
            assert binop(shim, f"shim_{name}_add", x, y) == (x + y) % p
            assert binop(shim, f"shim_{name}_sub", x, y) == (x - y) % p
        assert unop(shim, f"shim_{name}_neg", x) == (-x) % p
        assert unop(shim, f"shim_{name}_to_mont", x) == x * R % p
        assert unop(shim, f"shim_{name}_from_mont", x) == x * Ri % p
    for x in vals[:12]:
        assert unop(shim, f"shim_{name}_inv", x * R % p) == (pow(x, -1, p) * R % p if x else 0)
    assert shim.shim_lost_carries() == 0, "a carry was dropped by an instruction that cannot report it"


@pytest.mark.parametrize("name,p", [("fr", P.R_MOD), ("fq", P.Q_MOD)])
def test_kaliski_inverse(shim, name, p):
    """The binary-Euclid inverse the batched-affine bucket rounds use equals the a^(p-2) ladder / python pow."""
    rng = random.Random(5)
    edge = [0, 1, 2, 3, 4, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, R % p, 1 << 253, 1 << 200, (1 << 253) + 1, p - (1 << 200), 1 << 32, 1 << 64, 3 << 96]
    for x in edge + [rng.randrange(p) for _ in range(400)] + [rng.randrange(1 << 40) for _ in range(50)]:
        got = unop(shim, f"shim_{name}_inv_kaliski", x)
        assert got == (pow(x, -1, p) * R * R % p if x else 0), hex(x)
    assert shim.shim_lost_carries() == 0


def test_fq2_ops(shim):
    rng = random.Random(2)
    q = P.Q_MOD
    def mont2(a): return (a[0] * R % q, a[1] * R % q)
    def pack(a): return A2(*[(v >> (32 * i)) & 0xFFFFFFFF for v in a for i in range(8)])
    def unpack(o): return (frl(o[:8]), frl(o[8:]))
    for _ in range(50):
        a = (rng.randrange(q), rng.randrange(q)); b = (rng.randrange(q), rng.randrange(q))
        o = A2(); shim.shim_fq2_mul(pack(mont2(a)), pack(mont2(b)), o)
        assert unpack(o) == mont2(P.f2_mul(a, b))
        o = A2(); shim.shim_fq2_sqr(pack(mont2(a)), o)
        assert unpack(o) == mont2(P.f2_mul(a, a))
        o = A2(); shim.shim_fq2_inv(pack(mont2(a)), o)
        assert unpack(o) == mont2(P.f2_inv(a))


def test_curve_ops_host(shim):
    """XYZZ formulas of ec.cuh on the host, including the exceptional cases."""
    rng = random.Random(3)
    q = P.Q_MOD
    PT = ctypes.c_uint32 * 16
    def pack(p): return PT(*([0] * 16)) if p is None else PT(*[((v * R % q) >> (32 * i)) & 0xFFFFFFFF for v in p for i in range(8)])
    def unpack(o):
        x, y = frl(o[:8]), frl(o[8:])
        Ri = pow(R, -1, q)
        return None if x == 0 and y == 0 else (x * Ri % q, y * Ri % q)
    g = P.G1_GEN
    p5, p7 = P.g1_mul(g, 5), P.g1_mul(g, 7)
    cases = [(p5, p7), (p5, p5), (p5, P.ec_neg(P.Fq1Ops, p5)), (None, p7), (p5, None), (None, None)]
    for a, b in cases:
        o = PT(); shim.shim_g1_add_mixed(pack(a), pack(b), o); assert unpack(o) == P.g1_add(a, b), ("madd", a, b)
        o = PT(); shim.shim_g1_add(pack(a), pack(b), o); assert unpack(o) == P.g1_add(a, b), ("add", a, b)
    k = rng.randrange(P.R_MOD)
    o = PT(); shim.shim_g1_mul(pack(p5), tol(k), o); assert unpack(o) == P.g1_mul(p5, k)
    o = PT(); shim.shim_g1_mul(pack(p5), tol(P.R_MOD), o); assert unpack(o) is None
    # external known answer (EIP-196 test vectors): 2 * (1, 2), through the product's own ec.cuh
    two_g = (0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3, 0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4)
    o = PT(); shim.shim_g1_mul(pack(P.G1_GEN), tol(2), o); assert unpack(o) == two_g
    o = PT(); shim.shim_g1_add(pack(P.G1_GEN), pack(P.G1_GEN), o); assert unpack(o) == two_g


def test_wide_products_and_reduction(shim):
    """u256_mul_wide / u256_sqr_wide on arbitrary 256-bit operands (the lazily reduced Fq2 product feeds unreduced sums
    below 2^255 and full limbs), fp_redc_wide on every T < p 2^256, and the dedicated squaring."""
    rng = random.Random(7)
    full = (1 << 256) - 1
    ops = [0, 1, full, full - 1, 1 << 255, (1 << 255) - 1, 0xFFFFFFFF, full ^ 0xFFFFFFFF, P.Q_MOD, 2 * P.Q_MOD - 2, P.R_MOD - 1] + \
          [rng.getrandbits(256) for _ in range(200)] + [rng.getrandbits(256) | (full << 224 & full) for _ in range(20)]
    T16 = ctypes.c_uint32 * 16
    for x in ops:
        o = T16(); shim.shim_sqr_wide(tol(x), o); assert frl(o) == x * x, hex(x)
        for y in rng.sample(ops, 6) + [full, 0, 1]:
            o = T16(); shim.shim_mul_wide(tol(x), tol(y), o); assert frl(o) == x * y, (hex(x), hex(y))
    for name, p in (("fr", P.R_MOD), ("fq", P.Q_MOD)):
        Ri = pow(R, -1, p)
        wide = [0, 1, p, p * R - 1, p * R - p, (p - 1) * (p - 1), 2 * (p - 1) * (p - 1), R - 1, R, R + 1, (R - 1) * p] + \
               [rng.randrange(p * R) for _ in range(300)]
        for t in wide:
            o = A(); getattr(shim, f"shim_{name}_redc_wide")(T16(*[(t >> (32 * i)) & 0xFFFFFFFF for i in range(16)]), o)
            assert frl(o) == t * Ri % p, hex(t)
        edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, R % p, (1 << 253), p - (1 << 200), (1 << 254) - 1 if (1 << 254) - 1 < p else p - 3]
        for x in edge + [rng.randrange(p) for _ in range(300)]:
            assert unop(shim, f"shim_{name}_sqr", x) == x * x * Ri % p, hex(x)
    assert shim.shim_lost_carries() == 0


def test_fq2_lazy_product(shim):
    """Three wide products + two reductions == Karatsuba with three full products == the python reference,
    including the operands that make a0 b0 - a1 b1 negative, zero, and the largest sums."""
    rng = random.Random(8)
    q = P.Q_MOD
    def pack(a): return A2(*[(v >> (32 * i)) & 0xFFFFFFFF for v in a for i in range(8)])
    def unpack(o): return (frl(o[:8]), frl(o[8:]))
    Ri = pow(R, -1, q)
    edge = [0, 1, q - 1, q - 2, (q - 1) // 2, R % q]
    cases = [((a0, a1), (b0, b1)) for a0 in edge for a1 in edge for b0 in (0, 1, q - 1) for b1 in (0, q - 1, 5)]
    cases += [((rng.randrange(q), rng.randrange(q)), (rng.randrange(q), rng.randrange(q))) for _ in range(300)]
    for a, b in cases:
        want = ((a[0] * b[0] - a[1] * b[1]) * Ri % q, (a[0] * b[1] + a[1] * b[0]) * Ri % q)
        o = A2(); shim.shim_fq2_mul_lazy(pack(a), pack(b), o); assert unpack(o) == want, (a, b)
        o = A2(); shim.shim_fq2_mul_karatsuba(pack(a), pack(b), o); assert unpack(o) == want, (a, b)
    assert shim.shim_lost_carries() == 0
