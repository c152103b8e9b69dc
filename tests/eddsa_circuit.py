"""Config 3 (BASELINE.json: "interop circomlib BabyJubJub/EdDSA circuit"): the STATEMENT of circomlib's EdDSAMiMCVerifier
(/root/reference/interop/circuits/circomlib/circuits/eddsamimc.circom:27-128) as an R1CS built by hand.

za's front-end (parser, evaluator, optimiser) is out of the hot path's scope, so the circuit cannot be compiled from its
source here; what the prover sees is a constraint system and a witness, and this module builds both for the same
statement with its own decomposition:
    S < 2^253, S * B8 == R8 + H(R8x, R8y, Ax, Ay, M) * (8 A),   H = MultiMiMC7(5, 91) with key 0  (mimc.circom:20-158)
on the inputs of the reference's own test (za_test/eddsamimc.za:4-13; all seven signals public).
The gadgets: Num2Bits (booleanity + recomposition), MiMC7 rounds (x^7 in four multiplications), BabyAdd
(babyjub.circom:23-50, complete twisted-Edwards addition, six constraints), a double-and-add ladder for h * (8 A), a
fixed-base sum of selected 2^i B8 for S * B8, and `enabled`-gated equality (ForceEqualIfEnabled).  It is NOT the
constraint system za would emit (window methods, Montgomery form and the optimiser give a different, smaller system); it
has the same witness character — thousands of boolean signals — which is what the multiexps see.

Every `enforce` is checked against the witness as it is added, so a circuit that builds is satisfied.
"""
import json
import os

import numpy as np

from tests.pyref import R_MOD as R

AUX = 0x80000000
A_ED, D_ED = 168700, 168696
BASE8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
         16950150798460657717958625567821834550301663161624707787222815936182638968203)
# za_test/eddsamimc.za:6-12
KAT = dict(enabled=1,
           Ax=13277427435165878497778222415993513565335242147425444199013288855685581939618,
           Ay=13622229784656158136036771217484571176836296686641868549125388198837476602820,
           R8x=11220723668893468001994760120794694848178115379170651044669708829805665054484,
           R8y=2367470421002446880004241260470975644531657398480773647535134774673409612366,
           S=1701898193987160140374512573986329501685719384866194117894109500242212188181,
           M=1234)


def mimc7_constants():
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mimc7_constants.json")
    return [int(x) for x in json.load(open(p))["c"]]


# ------------------------------------------------------------------ plain-integer semantics (the witness generator)
def baby_add(p, q):
    x1, y1 = p
    x2, y2 = q
    beta, gamma = x1 * y2 % R, y1 * x2 % R
    delta = (-A_ED * x1 + y1) * (x2 + y2) % R
    tau = beta * gamma % R
    return ((beta + gamma) * pow(1 + D_ED * tau, -1, R) % R, (delta + A_ED * beta - gamma) * pow(1 - D_ED * tau, -1, R) % R)


def baby_mul(p, k):
    acc = (0, 1)
    for b in bin(k)[2:]:
        acc = baby_add(acc, acc)
        if b == "1":
            acc = baby_add(acc, p)
    return acc


def mimc7(x, k, c):
    t = (k + x) % R
    for i in range(len(c)):
        if i:
            t = (k + t7 + c[i]) % R
        t2 = t * t % R
        t4 = t2 * t2 % R
        t6 = t4 * t2 % R
        if i < len(c) - 1:
            t7 = t6 * t % R
        else:
            return (t6 * t + k) % R


def multi_mimc7(ins, k, c):
    r = k
    for x in ins:
        r = (r + x + mimc7(x, r, c)) % R
    return r


# ------------------------------------------------------------------ R1CS builder
class Builder:
    """Linear combinations are dicts {variable: coefficient}; variable 0 is `one`, inputs 1.., aux AUX | k."""

    def __init__(self):
        self.inputs = [1]
        self.aux = []
        self.rows = []

    def input(self, v):
        self.inputs.append(v % R)
        return {len(self.inputs) - 1: 1}

    def new_aux(self, v):
        self.aux.append(v % R)
        return {AUX | (len(self.aux) - 1): 1}

    def value(self, lc):
        s = 0
        for v, c in lc.items():
            s += c * (self.aux[v & 0x7fffffff] if v & AUX else self.inputs[v])
        return s % R

    def enforce(self, a, b, c):
        assert self.value(a) * self.value(b) % R == self.value(c), "constraint not satisfied by the witness"
        self.rows.append((a, b, c))

    def product(self, a, b):
        out = self.new_aux(self.value(a) * self.value(b))
        self.enforce(a, b, out)
        return out

    def tuple(self):
        ptr = [[0], [0], [0]]
        var = [[], [], []]
        coeff = [[], [], []]
        for row in self.rows:
            for w in range(3):
                for v, c in sorted(row[w].items()):
                    if c % R:
                        var[w].append(v)
                        coeff[w].append(np.frombuffer(int(c % R).to_bytes(32, "little"), np.uint8))
                ptr[w].append(len(var[w]))
        fr = lambda x: np.frombuffer(int(x).to_bytes(32, "little"), np.uint8)
        return (len(self.inputs), len(self.aux), [np.array(p, np.uint32) for p in ptr], [np.array(v, np.uint32) for v in var],
                [np.stack(c) if c else np.zeros((0, 32), np.uint8) for c in coeff],
                np.stack([fr(v) for v in self.inputs]), np.stack([fr(v) for v in self.aux]))


def lc_add(*lcs):
    out = {}
    for lc in lcs:
        for v, c in lc.items():
            out[v] = (out.get(v, 0) + c) % R
    return out


def lc_scale(lc, k):
    return {v: c * k % R for v, c in lc.items()}


def const(k):
    return {0: k % R}


ONE = {0: 1}


def g_bits(b, lc, n):
    """Num2Bits(n): n boolean aux signals whose weighted sum is `lc`."""
    val = b.value(lc)
    assert val < (1 << n)
    bits = []
    acc = {}
    for i in range(n):
        bit = b.new_aux((val >> i) & 1)
        b.enforce(bit, lc_add(bit, const(-1)), {})                 # bit * (bit - 1) = 0
        bits.append(bit)
        acc = lc_add(acc, lc_scale(bit, 1 << i))
    b.enforce(acc, ONE, lc)
    return bits


def g_baby_add(b, p, q):
    """babyjub.circom:23-50 on linear combinations."""
    (x1, y1), (x2, y2) = p, q
    beta = b.product(x1, y2)
    gamma = b.product(y1, x2)
    delta = b.product(lc_add(lc_scale(x1, -A_ED), y1), lc_add(x2, y2))
    tau = b.product(beta, gamma)
    xo, yo = baby_add((b.value(x1), b.value(y1)), (b.value(x2), b.value(y2)))
    xout, yout = b.new_aux(xo), b.new_aux(yo)
    b.enforce(lc_add(ONE, lc_scale(tau, D_ED)), xout, lc_add(beta, gamma))
    b.enforce(lc_add(ONE, lc_scale(tau, -D_ED)), yout, lc_add(delta, lc_scale(beta, A_ED), lc_scale(gamma, -1)))
    return xout, yout


def g_mimc7(b, x, k, c):
    t = lc_add(k, x)
    for i in range(len(c)):
        if i:
            t = lc_add(k, t7, const(c[i]))
        t2 = b.product(t, t)
        t4 = b.product(t2, t2)
        t6 = b.product(t4, t2)
        t7 = b.product(t6, t)
    return lc_add(t7, k)


def eddsa_mimc_verifier(enabled, Ax, Ay, R8x, R8y, S, M, rounds=91):
    """Returns (cs_tuple, info).  Public inputs in za's order of declaration: enabled, Ax, Ay, S, R8x, R8y, M."""
    c = mimc7_constants()[:rounds]
    b = Builder()
    v_enabled, v_ax, v_ay, v_s, v_r8x, v_r8y, v_m = (b.input(x) for x in (enabled, Ax, Ay, S, R8x, R8y, M))
    s_bits = g_bits(b, v_s, 253)                                                   # S < 2^253 (the circuit also compares with the subgroup order)
    r = const(0)                                                                   # MultiMiMC7(5, rounds), key 0
    for x in (v_r8x, v_r8y, v_ax, v_ay, v_m):
        r = lc_add(r, x, g_mimc7(b, x, r, c))
    h_val = b.value(r)
    h_bits = g_bits(b, r, 254)
    p = (v_ax, v_ay)                                                               # 8 A by three doublings
    for _ in range(3):
        p = g_baby_add(b, p, p)
    acc = (const(0), const(1))                                                     # h * (8 A): double-and-add, most significant bit first
    for i in reversed(range(254)):
        acc = g_baby_add(b, acc, acc)
        sx = b.product(h_bits[i], p[0])                                            # bit ? P : (0, 1)
        sy = lc_add(b.product(h_bits[i], lc_add(p[1], const(-1))), const(1))
        acc = g_baby_add(b, acc, (sx, sy))
    right = g_baby_add(b, (v_r8x, v_r8y), acc)
    left = (const(0), const(1))                                                    # S * B8: sum of the selected 2^i B8 (constants)
    q = BASE8
    for i in range(253):
        sel = (lc_scale(s_bits[i], q[0]), lc_add(lc_scale(s_bits[i], q[1] - 1), const(1)))
        left = g_baby_add(b, left, sel)
        q = baby_add(q, q)
    b.enforce(lc_add(left[0], lc_scale(right[0], -1)), v_enabled, {})              # ForceEqualIfEnabled
    b.enforce(lc_add(left[1], lc_scale(right[1], -1)), v_enabled, {})
    return b.tuple(), dict(h=h_val, constraints=len(b.rows), left=(b.value(left[0]), b.value(left[1])))
