"""Independent python-integer reference for BN254 / Groth16 (test infrastructure).

Used by tests/golden/make_golden.py to produce the committed golden vectors and
by the CPU tests to cross-check the C oracle.  It shares no code and no
algorithm with either the oracle or the CUDA product: fields are python ints,
curve points are affine tuples, the NTT is the O(n^2) definition, the MSM is the
naive sum and the Groth16 proof is computed in closed form *from the toxic
waste* (scalars first, one scalar multiplication per proof element).

Constants pinned by the reference tree:
  r : /root/reference/compiler/src/algebra/fs.rs:15-16
  q, G1 generator, G2 generator : /root/reference/prover/src/groth16/ethereum.rs:37,22,28-31
Conventions restated from the un-vendored bellman_ce/pairing_ce:
  multiplicative generator 7, 2-adicity 28, omega = (7^((r-1)/2^28))^(2^(28-e)).
"""
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
FR_S = 28
FR_GEN = 7
G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


# ---------------------------------------------------------------- Fq2
def f2_add(a, b): return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)
def f2_sub(a, b): return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)
def f2_neg(a): return ((-a[0]) % Q_MOD, (-a[1]) % Q_MOD)
def f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, Q_MOD)
    return (a[0] * n % Q_MOD, (-a[1] * n) % Q_MOD)


B1 = 3
B2 = f2_mul((3, 0), f2_inv((9, 1)))


class Fq1Ops:
    zero = 0
    @staticmethod
    def add(a, b): return (a + b) % Q_MOD
    @staticmethod
    def sub(a, b): return (a - b) % Q_MOD
    @staticmethod
    def mul(a, b): return a * b % Q_MOD
    @staticmethod
    def inv(a): return pow(a, -1, Q_MOD)
    @staticmethod
    def neg(a): return (-a) % Q_MOD


class Fq2Ops:
    zero = (0, 0)
    add = staticmethod(f2_add)
    sub = staticmethod(f2_sub)
    mul = staticmethod(f2_mul)
    inv = staticmethod(f2_inv)
    neg = staticmethod(f2_neg)


def ec_add(F, p, q):
    """affine addition; None is the point at infinity"""
    if p is None: return q
    if q is None: return p
    if p[0] == q[0]:
        if p[1] != q[1] or p[1] == F.zero:
            return None
        x2 = F.mul(p[0], p[0])
        lam = F.mul(F.add(F.add(x2, x2), x2), F.inv(F.add(p[1], p[1])))
    else:
        lam = F.mul(F.sub(q[1], p[1]), F.inv(F.sub(q[0], p[0])))
    x3 = F.sub(F.sub(F.mul(lam, lam), p[0]), q[0])
    y3 = F.sub(F.mul(lam, F.sub(p[0], x3)), p[1])
    return (x3, y3)


def ec_neg(F, p): return None if p is None else (p[0], F.neg(p[1]))


def ec_mul(F, p, k):
    r = None
    k %= R_MOD
    while k:
        if k & 1: r = ec_add(F, r, p)
        p = ec_add(F, p, p)
        k >>= 1
    return r


def g1_mul(p, k): return ec_mul(Fq1Ops, p, k)
def g1_add(p, q): return ec_add(Fq1Ops, p, q)
def g2_mul(p, k): return ec_mul(Fq2Ops, p, k)
def g2_add(p, q): return ec_add(Fq2Ops, p, q)
def g1_on_curve(p): return p is None or (p[1] * p[1] - p[0] ** 3 - B1) % Q_MOD == 0
def g2_on_curve(p): return p is None or f2_sub(f2_mul(p[1], p[1]), f2_add(f2_mul(f2_mul(p[0], p[0]), p[0]), B2)) == (0, 0)


def msm(F, bases, scalars):
    acc = None
    for b, s in zip(bases, scalars):
        acc = ec_add(F, acc, ec_mul(F, b, s))
    return acc


# ---------------------------------------------------------------- domain
def omega(log_n):
    w = pow(FR_GEN, (R_MOD - 1) >> FR_S, R_MOD)
    for _ in range(log_n, FR_S):
        w = w * w % R_MOD
    return w


def dft(v, w):
    """out[k] = sum_j v[j] w^(jk)  — the definition, O(n^2)"""
    n = len(v)
    return [sum(v[j] * pow(w, j * k, R_MOD) for j in range(n)) % R_MOD for k in range(n)]


def fft(v): return dft(v, omega(len(v).bit_length() - 1))
def ifft(v):
    n = len(v)
    ninv = pow(n, -1, R_MOD)
    return [x * ninv % R_MOD for x in dft(v, pow(omega(n.bit_length() - 1), -1, R_MOD))]
def coset_fft(v): return fft([x * pow(FR_GEN, i, R_MOD) % R_MOD for i, x in enumerate(v)])
def icoset_fft(v):
    gi = pow(FR_GEN, -1, R_MOD)
    return [x * pow(gi, i, R_MOD) % R_MOD for i, x in enumerate(ifft(v))]


def h_poly(a, b, c):
    """bellman create_proof step 4 by the definitions above."""
    m = 1
    while m < len(a): m *= 2
    pad = lambda v: list(v) + [0] * (m - len(v))
    A, B, C = (coset_fft(ifft(pad(v))) for v in (a, b, c))
    zinv = pow(pow(FR_GEN, m, R_MOD) - 1, -1, R_MOD)
    t = [(x * y - z) * zinv % R_MOD for x, y, z in zip(A, B, C)]
    return icoset_fft(t)[: m - 1]


# ---------------------------------------------------------------- R1CS + closed-form Groth16
AUX = 0x80000000


class R1CS:
    """rows of bellman enforce(A, B, C); a term is (coeff, var); var has bit 31 set for aux."""
    def __init__(self, num_inputs, num_aux, rows):
        self.num_inputs, self.num_aux, self.rows = num_inputs, num_aux, rows

    def value(self, var, inputs, aux):
        return aux[var & ~AUX] if var & AUX else inputs[var]

    def evals(self, inputs, aux):
        out = ([], [], [])
        for row in self.rows:
            for w in range(3):
                out[w].append(sum(c * self.value(v, inputs, aux) for c, v in row[w]) % R_MOD)
        for i in range(self.num_inputs):
            out[0].append(inputs[i] % R_MOD); out[1].append(0); out[2].append(0)
        return out


def lagrange_at(tau, m):
    """L_j(tau) over the domain {omega^j}"""
    w = omega(m.bit_length() - 1)
    z = (pow(tau, m, R_MOD) - 1) % R_MOD
    minv = pow(m, -1, R_MOD)
    return [z * minv % R_MOD * pow(w, j, R_MOD) % R_MOD * pow(tau - pow(w, j, R_MOD), -1, R_MOD) % R_MOD for j in range(m)]


def groth16_closed_form(cs, inputs, aux, toxic, r, s, g1=G1_GEN, g2=G2_GEN):
    """Proof (A, B, C) computed as scalars from the toxic waste, then three scalar multiplications.
    Also returns the verifying-key scalars.  Independent of any FFT/MSM algorithm."""
    alpha, beta, gamma, delta, tau = toxic
    nrows = len(cs.rows) + cs.num_inputs
    m = 1
    while m < nrows: m *= 2
    L = lagrange_at(tau, m)
    nv = cs.num_inputs + cs.num_aux
    at, bt, ct = [0] * nv, [0] * nv, [0] * nv
    slot = lambda v: cs.num_inputs + (v & ~AUX) if v & AUX else v
    for k, row in enumerate(cs.rows):
        for acc, terms in zip((at, bt, ct), row):
            for c, v in terms:
                acc[slot(v)] = (acc[slot(v)] + c * L[k]) % R_MOD
    for i in range(cs.num_inputs):
        at[i] = (at[i] + L[len(cs.rows) + i]) % R_MOD
    w = list(inputs) + list(aux)
    A_t = sum(x * y for x, y in zip(w, at)) % R_MOD
    B_t = sum(x * y for x, y in zip(w, bt)) % R_MOD
    C_t = sum(x * y for x, y in zip(w, ct)) % R_MOD
    z = (pow(tau, m, R_MOD) - 1) % R_MOD
    assert z != 0
    dinv = pow(delta, -1, R_MOD)
    a_s = (alpha + A_t + r * delta) % R_MOD
    b_s = (beta + B_t + s * delta) % R_MOD
    l_part = sum(w[i] * (beta * at[i] + alpha * bt[i] + ct[i]) for i in range(cs.num_inputs, nv)) % R_MOD
    hz = (A_t * B_t - C_t) % R_MOD          # = h(tau) Z(tau) for a satisfying witness
    c_s = ((l_part + hz) * dinv + a_s * s + b_s * r - r * s * delta) % R_MOD
    proof = (g1_mul(g1, a_s), g2_mul(g2, b_s), g1_mul(g1, c_s))
    ginv = pow(gamma, -1, R_MOD)
    ic = [g1_mul(g1, (beta * at[i] + alpha * bt[i] + ct[i]) * ginv % R_MOD) for i in range(cs.num_inputs)]
    vk = dict(alpha_g1=g1_mul(g1, alpha), beta_g1=g1_mul(g1, beta), beta_g2=g2_mul(g2, beta),
              gamma_g2=g2_mul(g2, gamma), delta_g1=g1_mul(g1, delta), delta_g2=g2_mul(g2, delta), ic=ic)
    return proof, vk


def example_factor_circuit():
    """/root/reference/example/circuit.za (`p * q === r`, r public) after CircomCircuit::synthesize
    (prover.rs:45-103): signals one, main.r (public input), main.p, main.q (aux);
    za constraint a*b + c = 0 with a=[p], b=[q], c=[-r]  ->  bellman A=[p], B=[q], C=-c=[r]."""
    return R1CS(2, 2, [([(1, AUX | 0)], [(1, AUX | 1)], [(1, 1)])])


def mul_circuit_test():
    """prover.rs:226-236 `c <== a*b` : signals one, main.c (output -> input 1), main.a, main.b (aux)."""
    return R1CS(2, 2, [([(1, AUX | 0)], [(1, AUX | 1)], [(1, 1)])])
