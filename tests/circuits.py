"""Synthetic constraint systems used by the parity tests and the bench.

Everything is expressed as what bellman's ConstraintSystem sees after CircomCircuit::synthesize
(/root/reference/prover/src/groth16/prover.rs:45-103): inputs (input 0 = one), aux, and
enforce(A, B, C) rows in CSR form with 32-byte LE canonical coefficients.
"""
import numpy as np

from tests.pyref import R_MOD

AUX = 0x80000000


def _fr(x):
    return np.frombuffer(int(x % R_MOD).to_bytes(32, "little"), np.uint8)


def mul_chain(num_constraints, x0=3, extra_linear=True):
    """Config 2/4: signals `out` (public), x0 (private), x1..; constraints x_{i+1} = x_i * x_i,
    last one constrains out.  Full A and B density.  With extra_linear every third row also carries a
    linear term with a non-unit coefficient so the coefficient multiply path is exercised.

    Returns (num_inputs, num_aux, ptr[3], var[3], coeff[3], inputs(ni,32), aux(na,32))."""
    nc = num_constraints
    na = nc                     # x0 .. x_{nc-1}; x_nc is the public output
    vals = [x0 % R_MOD]
    for _ in range(nc):
        vals.append(vals[-1] * vals[-1] % R_MOD)
    out = vals[nc]
    one = _fr(1)
    ptr = [[0], [0], [0]]
    var = [[], [], []]
    coeff = [[], [], []]
    for k in range(nc):
        # A = x_k (+ 5*one - 5*one folded as two terms on some rows), B = x_k, C = x_{k+1}
        var[0].append(AUX | k); coeff[0].append(one)
        if extra_linear and k % 3 == 0:
            var[0].append(0); coeff[0].append(_fr(5))
            var[0].append(0); coeff[0].append(_fr(R_MOD - 5))
        var[1].append(AUX | k); coeff[1].append(one)
        if k + 1 < nc:
            var[2].append(AUX | (k + 1)); coeff[2].append(one)
        else:
            var[2].append(1); coeff[2].append(one)
        for w in range(3):
            ptr[w].append(len(var[w]))
    inputs = np.stack([_fr(1), _fr(out)])
    aux = np.stack([_fr(v) for v in vals[:nc]])
    return (2, na, [np.array(p, np.uint32) for p in ptr], [np.array(v, np.uint32) for v in var],
            [np.stack(c) if c else np.zeros((0, 32), np.uint8) for c in coeff], inputs, aux)


def mul_chain_fast(num_constraints, x0=3):
    """The config-2/4 circuit (za_b200.synthetic.mul_chain)."""
    from za_b200.synthetic import mul_chain as mc
    return mc(num_constraints, x0)


def witness_like(n, seed):
    """Config 5a second distribution: 40 % zeros, 30 % ones, 30 % uniform."""
    from za_b200.synthetic import witness_like_scalars
    return witness_like_scalars(n, seed)


# ---- circomlib BabyAdd (config 3's building block), extracted by hand from
# /root/reference/interop/circuits/circomlib/circuits/babyjub.circom:23-50; the known answers are the reference's own
# tests, /root/reference/interop/circuits/circomlib/za_test/babyjub.za:4-34.
BABYJUB_A, BABYJUB_D = 168700, 168696
BABYADD_KATS = [
    ((0, 1, 0, 1), (0, 1)),                                                                            # babyjub.za:4-13
    ((5299619240641551281634865583518297030282874472190772894086521144482721001553,
      16950150798460657717958625567821834550301663161624707787222815936182638968203,
      5299619240641551281634865583518297030282874472190772894086521144482721001553,
      16950150798460657717958625567821834550301663161624707787222815936182638968203),
     (10031262171927540148667355526369034398030886437092045105752248699557385197826,
      633281375905621697187330766174974863687049529291089048651929454608812697683)),                    # babyjub.za:15-23
    ((5299619240641551281634865583518297030282874472190772894086521144482721001553,
      16950150798460657717958625567821834550301663161624707787222815936182638968203,
      16540640123574156134436876038791482806971768689494387082833631921987005038935,
      20819045374670962167435360035096875258406992893633759881276124905556507972311),
     (21523367759672787045219488891085974541594844619335119872356290525923952717026,
      9057049290837315782679166527865692184803931233813301783664357505338259629573)),                   # babyjub.za:25-34
]


def babyjub_add(x1, y1, x2, y2):
    """BabyAdd as bellman sees it: inputs [one, xout, yout] (the outputs of main are public), aux [x1, y1, x2, y2, beta,
    gamma, delta, tau]; six enforce(A, B, C) rows, A * B = C:
        x1 * y2 = beta;  y1 * x2 = gamma;  (-a x1 + y1) * (x2 + y2) = delta;  beta * gamma = tau;
        (1 + d tau) * xout = beta + gamma;  (1 - d tau) * yout = delta + a beta - gamma.
    Returns the same tuple as mul_chain plus (xout, yout)."""
    a, d, r = BABYJUB_A, BABYJUB_D, R_MOD
    beta, gamma = x1 * y2 % r, y1 * x2 % r
    delta = (-a * x1 + y1) * (x2 + y2) % r
    tau = beta * gamma % r
    xout = (beta + gamma) * pow(1 + d * tau, -1, r) % r
    yout = (delta + a * beta - gamma) * pow(1 - d * tau, -1, r) % r
    ONE, XOUT, YOUT = 0, 1, 2
    X1, Y1, X2, Y2, BETA, GAMMA, DELTA, TAU = (AUX | i for i in range(8))
    rows = [
        ([(X1, 1)], [(Y2, 1)], [(BETA, 1)]),
        ([(Y1, 1)], [(X2, 1)], [(GAMMA, 1)]),
        ([(X1, -a), (Y1, 1)], [(X2, 1), (Y2, 1)], [(DELTA, 1)]),
        ([(BETA, 1)], [(GAMMA, 1)], [(TAU, 1)]),
        ([(ONE, 1), (TAU, d)], [(XOUT, 1)], [(BETA, 1), (GAMMA, 1)]),
        ([(ONE, 1), (TAU, -d)], [(YOUT, 1)], [(DELTA, 1), (BETA, a), (GAMMA, -1)]),
    ]
    ptr = [[0], [0], [0]]
    var = [[], [], []]
    coeff = [[], [], []]
    for row in rows:
        for w in range(3):
            for v, c in row[w]:
                var[w].append(v); coeff[w].append(_fr(c))
            ptr[w].append(len(var[w]))
    inputs = np.stack([_fr(1), _fr(xout), _fr(yout)])
    aux = np.stack([_fr(v) for v in (x1, y1, x2, y2, beta, gamma, delta, tau)])
    cs = (3, 8, [np.array(p, np.uint32) for p in ptr], [np.array(v, np.uint32) for v in var], [np.stack(c) for c in coeff], inputs, aux)
    return cs, (xout, yout)
