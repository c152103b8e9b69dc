"""Synthetic workloads of the named sizes — inputs only, no arithmetic on the proving path.

mul_chain: the configuration-2/4 circuit as bellman's ConstraintSystem sees it after CircomCircuit::synthesize
(/root/reference/prover/src/groth16/prover.rs:45-103): signals `out` (public), x0 (private), x1..;
constraints x_{i+1} = x_i * x_i; full A and B density; with nc = 2^k - 2 constraints and the inputs
{one, out} the evaluation domain is exactly 2^k.
"""
import numpy as np

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
AUX = 0x80000000


def mul_chain(num_constraints, x0=3):
    """Returns (num_inputs, num_aux, ptr[3], var[3], coeff[3], inputs (2,32), aux (nc,32)); canonical LE bytes.
    The squaring chain is sequential python-integer work (~1 us per step)."""
    nc = num_constraints
    vals = np.zeros((nc + 1, 32), np.uint8)
    v = x0 % R_MOD
    for k in range(nc + 1):
        vals[k] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
        v = v * v % R_MOD
    one = np.zeros((1, 32), np.uint8)
    one[0, 0] = 1
    ptr = np.arange(nc + 1, dtype=np.uint32)
    va = np.arange(nc, dtype=np.uint32) | np.uint32(AUX)
    vc = np.arange(1, nc + 1, dtype=np.uint32) | np.uint32(AUX)
    vc[nc - 1] = 1                                   # the last product is the public output (input 1)
    ones = np.repeat(one, nc, axis=0)
    inputs = np.stack([one[0], vals[nc]])
    return (2, nc, [ptr, ptr.copy(), ptr.copy()], [va, va.copy(), vc], [ones, ones.copy(), ones.copy()], inputs, vals[:nc].copy())


def pk_counts_for_mul_chain(num_constraints):
    """Query sizes of a proving key for mul_chain(nc): ic, h, l, a, b_g1, b_g2."""
    nc = num_constraints
    m = 1
    while m < nc + 2:
        m *= 2
    # every aux variable occurs in A and in B rows; no input occurs in a B row
    return dict(ic=2, h=m - 1, l=nc, a=2 + nc, b_g1=nc, b_g2=nc)


def random_scalars(n, seed):
    """(n, 32) uint8: uniform 253-bit values (all canonical: 2^253 < r)."""
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    s[:, 31] &= 0x1F
    return s


def witness_like_scalars(n, seed):
    """Config 5a second distribution: 40 % zeros, 30 % ones, 30 % uniform."""
    rng = np.random.default_rng(seed)
    s = random_scalars(n, seed + 1)
    u = rng.random(n)
    s[u < 0.4] = 0
    one = np.zeros(32, np.uint8)
    one[0] = 1
    s[(u >= 0.4) & (u < 0.7)] = one
    return s
