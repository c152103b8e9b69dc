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
