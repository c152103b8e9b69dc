"""Host-only properties of the multi-GPU plan (za_prover_plan_counts): for any query lengths and any number of
devices the ranges of the devices tile every query exactly once (a proof is the sum of the devices' partial results), the H
ranges ascend with the device index (the last NTT pass scatters h by ascending bounds, za_ctx_set_h_scatter), one device gets
everything, and at the benchmark size a device touches few queries with long ranges."""
import random

import pytest

import za_b200


def _tiles(plans, cnt):
    for q in range(5):
        ivs = sorted((lo[q], hi[q]) for lo, hi in plans if hi[q] > lo[q])
        if cnt[q] == 0:
            assert not ivs
            continue
        assert ivs[0][0] == 0 and ivs[-1][1] == cnt[q] and all(a[1] == b[0] for a, b in zip(ivs, ivs[1:])), (q, ivs)
    prev = 0
    for lo, hi in plans:                                     # H: contiguous in device order
        assert lo[0] == prev and hi[0] >= lo[0]
        prev = hi[0]


def test_plan_tiles_every_query_for_random_sizes():
    rnd = random.Random(20261017)
    for _ in range(1500):
        m = 1 << rnd.randint(1, 24)
        b = rnd.randint(0, m)
        cnt = [m - 1, rnd.randint(0, m), rnd.randint(0, m + 5), b, b]
        if rnd.random() < 0.1:
            cnt[1] = 0
        if rnd.random() < 0.1:
            cnt[3] = cnt[4] = 0
        _tiles(za_b200.prover_plan_counts(cnt, m, rnd.choice([1, 2, 3, 4, 5, 7, 8, 16, 33, 64])), cnt)


def test_one_device_gets_everything():
    m = 1 << 12
    cnt = [m - 1, 4000, 4096, 3000, 3000]
    (lo, hi), = za_b200.prover_plan_counts(cnt, m, 1)
    assert lo == [0] * 5 and hi == cnt


@pytest.mark.parametrize("n", [2, 4, 8])
def test_plan_at_the_benchmark_size(n):
    """2^20: every device adds up at most three witness pieces (one contiguous piece of the time line), device 0 (H pipeline
    first) gets the smallest share of H, and nobody gets a sliver of a query."""
    m = 1 << 20
    cnt = [m - 1, m - 2, m, m - 1, m - 1]
    plans = za_b200.prover_plan_counts(cnt, m, n)
    _tiles(plans, cnt)
    for lo, hi in plans:
        pieces = [(hi[q] - lo[q]) / cnt[q] for q in range(1, 5) if hi[q] > lo[q]]
        assert 1 <= len(pieces) <= 3 and min(pieces) > 0.05
        if hi[0] == lo[0]:                                   # no H share: a device that runs nothing but its piece of B (G2)
            assert [q for q in range(1, 5) if hi[q] > lo[q]] == [4]
    shares = [(hi[0] - lo[0]) for lo, hi in plans if hi[0] > lo[0]]
    assert plans[0][1][0] - plans[0][0][0] == min(shares)   # device 0 (H pipeline first) takes the smallest H share


def test_bad_arguments():
    with pytest.raises(za_b200.ZaError):
        za_b200.prover_plan_counts([1, 1, 1, 1, 1], 4, 0)
    with pytest.raises(za_b200.ZaError):
        za_b200.prover_plan_counts([1, 1, 1, 1, 1], 4, 65)
