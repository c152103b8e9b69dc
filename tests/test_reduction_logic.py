"""Host-side logic of the row / column bucket reduction (za_b200/csrc/msm.cu, K6') restated with python integers.

The kernels compute, per bucket space of B = H * L buckets S_i (weight i + 1; bellman's multiexp.rs sums the buckets of a
window with a running sum, here restated as the weighted sum it equals):
    sum_i (i + 1) S_i = L * sum_hi hi R_hi + sum_lo (lo + 1) C_lo,      R_hi = sum_lo S_{hi L + lo},  C_lo = sum_hi S_{hi L + lo}
and each weighted sum of n <= 1024 points, ipt points per thread, as
    sum_i i X_i = sum_t (L_t + ipt [t >= 1] Suffix_t),   L_t = sum_j j X_{t ipt + j},   Suffix_t = sum of the points of the threads >= t.
Group elements are modelled by their discrete logarithms (integers mod r): the identities are linear, so they hold for
the group iff they hold for the exponents, and buckets at infinity are exponent 0.
"""
import random

import pytest

from tests import pyref as P

R = P.R_MOD


def weighted_by_threads(x, ipt, plus_one):
    """msm_small_weighted_kernel: out[0] (+ out[1] when the weights are i + 1)."""
    n = len(x)
    threads = max(1, n // ipt)
    T = [sum(x[t * ipt + j] for j in range(ipt) if t * ipt + j < n) % R for t in range(threads)]
    Lt = [sum(j * x[t * ipt + j] for j in range(ipt) if t * ipt + j < n) % R for t in range(threads)]
    suffix = [sum(T[t:]) % R for t in range(threads)]
    out0 = sum(Lt[t] + (ipt * suffix[t] if t >= 1 else 0) for t in range(threads)) % R
    out1 = suffix[0]
    return (out0 + (out1 if plus_one else 0)) % R


@pytest.mark.parametrize("n", [1, 2, 8, 256, 512, 1024])
def test_suffix_scan_weighted_sum(n):
    rnd = random.Random(n)
    x = [rnd.randrange(R) if rnd.random() > 0.2 else 0 for _ in range(n)]
    ipt = n // 256 if n > 256 else 1
    assert weighted_by_threads(x, ipt, False) == sum(i * v for i, v in enumerate(x)) % R
    assert weighted_by_threads(x, ipt, True) == sum((i + 1) * v for i, v in enumerate(x)) % R


@pytest.mark.parametrize("nb", [1, 4, 10, 11, 13, 15])
def test_row_column_decomposition(nb):
    """The split used by msm_enqueue: k = min(nb, 10) low bits; host combination of msm_finish."""
    rnd = random.Random(100 + nb)
    B = 1 << nb
    S = [rnd.randrange(R) if rnd.random() > 0.3 else 0 for _ in range(B)]
    k = min(nb, 10)
    Lw, H = 1 << k, B >> k
    rows = [sum(S[hi * Lw:(hi + 1) * Lw]) % R for hi in range(H)]
    cols = [sum(S[hi * Lw + lo] for hi in range(H)) % R for lo in range(Lw)]
    ipt_c = Lw // 256 if Lw > 256 else 1
    total = weighted_by_threads(cols, ipt_c, True)
    if H > 1:
        ipt_r = H // 256 if H > 256 else 1
        total = (total + (weighted_by_threads(rows, ipt_r, False) << k)) % R
    assert total == sum((i + 1) * v for i, v in enumerate(S)) % R


def test_eight_way_stage_tree_is_a_plain_sum():
    """msm_colsum_kernel stages: rows_in -> rows_in / 8 (or / rows_in when fewer than 8 are left)."""
    rnd = random.Random(7)
    for rows in (2, 8, 32, 512, 1024):
        v = [rnd.randrange(R) for _ in range(rows)]
        cur = list(v)
        while len(cur) > 1:
            f = 8 if len(cur) >= 8 else len(cur)
            cur = [sum(cur[g * f:(g + 1) * f]) % R for g in range(len(cur) // f)]
        assert cur[0] == sum(v) % R
