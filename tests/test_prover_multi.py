"""GPU: za_prover (several GPUs behind one call) against the single-context proof and the CPU oracle; the regression
tests of the round-1 review (a failed proof leaves the context usable, sort sharing with unequal tables, two contexts
in one process).  Device counts above what the box has are skipped."""
import numpy as np
import pytest

from tests import oracle as O
from tests import pyref as P
from tests import circuits

pytestmark = pytest.mark.gpu


def _ndev():
    from za_b200 import _lib
    return _lib.lib().za_device_count()


def _case(nc=9000):
    ni, na, ptr, var, coeff, inputs, aux = circuits.mul_chain_fast(nc, x0=13)
    ocs = O.CS(ni, na, ptr, var, coeff)
    prm = O.Params.generate(ocs, [31, 32, 33, 34, 35], threads=8)
    rc, exp = prm.create_proof(ocs, inputs, aux, 5, 6, threads=8)
    assert rc == 0
    return (ni, na, ptr, var, coeff, inputs, aux), prm, exp


@pytest.mark.parametrize("n_dev", [1, 2, 3, 4, 8])
def test_prover_matches_oracle(n_dev):
    """za_prover_create_proof on n devices: real key through Parameters::read on every device, host-buffer witness and
    resident witness, bit-identical to the oracle's proof."""
    import za_b200
    if _ndev() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    (ni, na, ptr, var, coeff, inputs, aux), prm, exp = _case()
    pr = za_b200.Prover(list(range(n_dev)))
    try:
        pr.load_pk(prm.write(), checked=True)
        pr.set_circuit(ni, na, ptr, var, coeff)
        for _ in range(3):                                     # generations: the h-slice events are re-armed per proof
            assert pr.create_proof(inputs, aux, 5, 6) == exp
        pr.upload_witness(inputs, aux)
        assert pr.create_proof(None, None, 5, 6) == exp
        assert za_b200.verify_proof(pr.vk(), exp, [int.from_bytes(inputs[1].tobytes(), "little")])
        # a witness element >= r is reported with its position, and the next proof on the same prover succeeds
        bad = aux.copy(); bad[na // 2] = 0xFF
        with pytest.raises(za_b200.ZaError) as e:
            pr.create_proof(inputs, bad, 5, 6)
        assert f"aux[{na // 2}]" in str(e.value)
        with pytest.raises(za_b200.ZaError):                   # r >= the modulus: refused before any GPU work
            pr.create_proof(inputs, aux, (1 << 256) - 1, 6)
        assert pr.create_proof(inputs, aux, 5, 6) == exp
        assert pr.launch_count() > 0
    finally:
        pr.close()


def test_prover_synthetic_key_same_proof_on_every_device_count():
    import za_b200
    from za_b200 import synthetic
    nc = (1 << 14) - 2
    ni, na, ptr, var, coeff, inputs, aux = circuits.mul_chain_fast(nc, x0=7)
    counts = synthetic.pk_counts_for_mul_chain(nc)
    proofs = []
    for n_dev in (1, 2, 4, 8):
        if _ndev() < n_dev:
            continue
        pr = za_b200.Prover(list(range(n_dev)))
        try:
            pr.synthetic_pk(counts["ic"], counts["h"], counts["l"], counts["a"], counts["b_g1"], counts["b_g2"])
            pr.set_circuit(ni, na, ptr, var, coeff)
            proofs.append(pr.create_proof(inputs, aux, 11, 13))
        finally:
            pr.close()
    assert len(set(proofs)) == 1


def test_failed_proof_leaves_the_context_usable(ctx):
    """Round-1 review: r >= modulus (or any failure after the first enqueue) must not leave multiexp slots busy."""
    import za_b200
    (ni, na, ptr, var, coeff, inputs, aux), prm, exp = _case(3000)
    pk = za_b200.Parameters.read(ctx, prm.write())
    circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
    for bad_r, bad_s in (((1 << 256) - 1, 6), (5, P.R_MOD)):
        with pytest.raises(za_b200.ZaError):
            za_b200.create_proof(ctx, pk, circ, inputs, aux, bad_r, bad_s)
        assert za_b200.create_proof(ctx, pk, circ, inputs, aux, 5, 6) == exp
    bad = aux.copy(); bad[7] = 0xFF
    with pytest.raises(za_b200.ZaError):
        za_b200.create_proof(ctx, pk, circ, inputs, bad, 5, 6)
    assert za_b200.create_proof(ctx, pk, circ, inputs, aux, 5, 6) == exp


def test_device_resident_scalars_are_range_checked(ctx):
    """Round-1 review: the device-pointer entry points refuse scalars >= r instead of returning a wrong sum."""
    import torch
    import za_b200
    n = 5000
    pts = O.g1_multiples(n)
    bases = za_b200.Bases(ctx, 1, pts)
    sc = O.random_frs(n, 3)
    d = torch.from_numpy(sc.copy()).cuda()
    ok = za_b200.multiexp_device(ctx, bases, d.data_ptr(), n)
    assert ok == za_b200.multiexp(ctx, bases, sc)
    sc2 = sc.copy(); sc2[1234] = 0xFF
    d2 = torch.from_numpy(sc2).cuda()
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.multiexp_device(ctx, bases, d2.data_ptr(), n)
    assert "scalars[1234]" in str(e.value)
    assert za_b200.multiexp_device(ctx, bases, d.data_ptr(), n) == ok


def test_b_queries_with_unequal_tables_do_not_share_a_sort(ctx, monkeypatch):
    """Round-1 review: when only one of b_g1 / b_g2 has a fixed-base table (or the windows differ) the G2 multiexp must
    sort for itself.  Forced here by building the key with tables and dropping the G2 table through a partition of a
    different window size is not expressible from outside, so the two layouts are forced with the table switch: key
    loaded without tables (ZA_MSM_TABLE=0) proves the same proof as the key with tables."""
    import za_b200
    (ni, na, ptr, var, coeff, inputs, aux), prm, exp = _case(6000)
    circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
    monkeypatch.setenv("ZA_MSM_TABLE", "0")
    pk0 = za_b200.Parameters.read(ctx, prm.write())
    assert za_b200.create_proof(ctx, pk0, circ, inputs, aux, 5, 6) == exp
    monkeypatch.delenv("ZA_MSM_TABLE")
    pk1 = za_b200.Parameters.read(ctx, prm.write())
    assert za_b200.create_proof(ctx, pk1, circ, inputs, aux, 5, 6) == exp


def test_two_contexts_in_one_process():
    """Round-1 review: the > 48 KiB shared-memory attribute of the NTT kernels is set per device, not per process."""
    import za_b200
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    d = O.random_frs(1 << 14, 5)
    exp = O.fft(d, 14, 0, threads=8)
    c0, c1 = za_b200.Context(0), za_b200.Context(1)
    try:
        assert np.array_equal(c0.fft(d), exp)
        assert np.array_equal(c1.fft(d), exp)
    finally:
        c0.close(); c1.close()


def test_circuit_satisfied_refuses_a_null_aux(ctx):
    """Round-1 review: za_circuit_satisfied with num_aux > 0 and aux == NULL returns ZA_ERR_INVALID instead of dereferencing it."""
    import ctypes
    import za_b200
    from za_b200 import _lib
    ni, na, ptr, var, coeff, inputs, aux = circuits.mul_chain_fast(300, x0=3)
    circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
    assert circ.first_unsatisfied(inputs, aux) is None
    bad = ctypes.c_int64(0)
    inp = np.ascontiguousarray(inputs, np.uint8)
    rc = _lib.lib().za_circuit_satisfied(ctx.h, circ.h, inp.ctypes.data_as(ctypes.c_void_p), None, ctypes.byref(bad))
    assert rc == -2 and b"aux" in _lib.lib().za_last_error()
    assert circ.first_unsatisfied(inputs, aux) is None       # the context is still usable
