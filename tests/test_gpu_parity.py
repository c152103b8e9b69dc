"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact or fail — everything here is integer arithmetic."""
import numpy as np
import pytest

from tests import oracle as O
from tests import pyref as P
from tests import circuits

pytestmark = pytest.mark.gpu


def test_library_reports_device(ctx):
    from za_b200 import _lib
    assert _lib.lib().za_device_count() >= 1


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_ntt_matches_oracle(ctx, log_n, mode):
    n = 1 << log_n
    d = O.random_frs(n, 100 + log_n)
    got = ctx.ntt(d, mode)
    exp = O.fft(d, log_n, mode, threads=8)
    assert np.array_equal(got, exp)


def test_ntt_edge_values(ctx):
    # all zeros, all r-1, a single one
    n = 1 << 12
    z = np.zeros((n, 32), np.uint8)
    assert np.array_equal(ctx.fft(z), z)
    top = np.tile(np.frombuffer((P.R_MOD - 1).to_bytes(32, "little"), np.uint8), (n, 1))
    assert np.array_equal(ctx.coset_fft(top), O.fft(top, 12, 2))
    e = z.copy(); e[5, 0] = 1
    assert np.array_equal(ctx.ifft(e), O.fft(e, 12, 1))


def test_ntt_rejects_non_canonical(ctx):
    import za_b200
    bad = np.full((8, 32), 0xFF, np.uint8)
    with pytest.raises(za_b200.ZaError):
        ctx.fft(bad)


def test_ntt_roundtrip_large(ctx):
    # size-independent property at a BASELINE size: icoset_fft(coset_fft(x)) == x, ifft(fft(x)) == x
    log_n = 22
    d = O.random_frs(1 << log_n, 7)
    assert np.array_equal(ctx.ifft(ctx.fft(d)), d)
    assert np.array_equal(ctx.icoset_fft(ctx.coset_fft(d)), d)


def test_ntt_spot_evaluation_large(ctx):
    # fft output k is the polynomial evaluated at omega^k: check a few k by Horner in python ints
    log_n = 14
    n = 1 << log_n
    d = O.random_frs(n, 9)
    got = O.np_to_frs(ctx.fft(d))
    coeffs = O.np_to_frs(d)
    w = P.omega(log_n)
    for k in (0, 1, 77, n - 1):
        x = pow(w, k, P.R_MOD)
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + c) % P.R_MOD
        assert got[k] == acc


@pytest.mark.parametrize("length", [1, 3, 4, 5, 100, 2048, 5000, (1 << 16) - 3])
def test_h_poly_matches_oracle(ctx, length):
    a, b, c = (O.random_frs(length, s) for s in (1, 2, 3))
    got, ck = ctx.h_poly(a, b, c, checkpoints=True)
    exp, eck = O.h_poly(a, b, c, threads=8, checkpoints=True)
    assert np.array_equal(ck, eck), "an intermediate EvaluationDomain vector differs"
    assert np.array_equal(got, exp)
    fused = ctx.h_poly(a, b, c)
    assert np.array_equal(fused, exp), "fused pipeline differs from bellman's sequence"


def _msm_case(ctx, group, n, scalars, density=None, offset=0, nbases=None):
    import za_b200
    nbases = nbases or n
    pts = O.g1_multiples(nbases) if group == 1 else O.g2_multiples(nbases)
    bases = za_b200.Bases(ctx, group, pts)
    got = za_b200.multiexp(ctx, bases, scalars, density=density, offset=offset)
    rc, exp = O.multiexp("g1" if group == 1 else "g2", pts[offset:], scalars, density=density, threads=8)
    assert rc == 0
    assert got == exp
    return got


@pytest.mark.parametrize("group", [1, 2])
@pytest.mark.parametrize("n", [1, 2, 31, 64, 65, 100, 1000, 5000])
def test_multiexp_matches_oracle_uniform(ctx, group, n):
    _msm_case(ctx, group, n, O.random_frs(n, 40 + n))


@pytest.mark.parametrize("group", [1, 2])
def test_multiexp_witness_like(ctx, group):
    n = 20000 if group == 1 else 6000
    _msm_case(ctx, group, n, circuits.witness_like(n, 5))


@pytest.mark.parametrize("group", [1, 2])
def test_multiexp_edge_scalars(ctx, group):
    n = 300
    s = O.random_frs(n, 3)
    s[0] = 0
    s[1] = np.frombuffer((1).to_bytes(32, "little"), np.uint8)
    s[2] = np.frombuffer((P.R_MOD - 1).to_bytes(32, "little"), np.uint8)      # maximum scalar: top signed window
    s[3] = np.frombuffer((2 ** 253).to_bytes(32, "little"), np.uint8)
    s[4] = np.frombuffer(((1 << 254) - 1 - ((1 << 254) - 1 - P.R_MOD + 1)).to_bytes(32, "little"), np.uint8)
    s[5:40] = s[2]                                                            # many equal scalars -> one bucket, many adds
    _msm_case(ctx, group, n, s)
    zeros = np.zeros((n, 32), np.uint8)
    assert _msm_case(ctx, group, n, zeros) == b"\0" * (64 if group == 1 else 128)


def test_multiexp_duplicate_and_opposite_points(ctx):
    # P = bucket (doubling branch) and P = -bucket (infinity branch) inside one bucket
    import za_b200
    n = 200
    pts = O.g1_multiples(n)
    pts[1::2] = pts[0::2]                      # every point twice
    neg = pts[10].copy()
    y = int.from_bytes(neg[32:].tobytes(), "little")
    neg[32:] = np.frombuffer((P.Q_MOD - y).to_bytes(32, "little"), np.uint8)
    pts[11] = neg                              # an opposite pair
    s = O.random_frs(n, 77)
    s[1::2] = s[0::2]                          # same scalar -> same buckets
    bases = za_b200.Bases(ctx, 1, pts)
    got = za_b200.multiexp(ctx, bases, s)
    rc, exp = O.multiexp("g1", pts, s, threads=4)
    assert rc == 0 and got == exp


def test_multiexp_density_and_offset(ctx):
    n_exp = 3000
    rng = np.random.default_rng(1)
    dens = (rng.random(n_exp) < 0.6).astype(np.uint8)
    cnt = int(dens.sum())
    for group in (1, 2):
        _msm_case(ctx, group, n_exp, O.random_frs(n_exp, 11), density=dens, offset=7, nbases=cnt + 7)


def test_multiexp_errors_like_bellman(ctx):
    import za_b200
    pts = O.g1_multiples(10)
    bases = za_b200.Bases(ctx, 1, pts)
    with pytest.raises(za_b200.ZaError) as e:            # bases run out -> io error (unexpected EOF)
        za_b200.multiexp(ctx, bases, O.random_frs(11, 1))
    assert e.value.code == -5
    pts[3] = 0                                           # a base at infinity with a non-zero exponent
    bases = za_b200.Bases(ctx, 1, pts)
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.multiexp(ctx, bases, O.random_frs(10, 1))
    assert e.value.code == -3
    s = O.random_frs(10, 1); s[3] = 0                    # ... but a zero exponent skips it
    rc, exp = O.multiexp("g1", pts, s)
    assert rc == 0 and za_b200.multiexp(ctx, bases, s) == exp
    off = O.g1_multiples(4); off[2, 0] ^= 1              # off-curve point
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.Bases(ctx, 1, off)
    assert e.value.code == -6


def test_multiexp_linearity_large(ctx):
    # bases (i+1)G: sum s_i (i+1) G == (sum s_i (i+1) mod r) G — one scalar multiplication checks 2^18 points
    import za_b200
    n = 1 << 18
    pts = O.g1_multiples(n)
    s = O.random_frs(n, 21)
    bases = za_b200.Bases(ctx, 1, pts)
    got = za_b200.multiexp(ctx, bases, s)
    vals = s.view(np.uint64).reshape(n, 4)
    tot = 0
    for limb in range(4):
        col = [int(x) for x in vals[:, limb]]
        tot += sum(c * (i + 1) for i, c in enumerate(col)) << (64 * limb)
    exp = O.g1_mul(P.G1_GEN, tot % P.R_MOD)
    assert O.g1_tuple(got) == exp


def _prove_case(ctx, cs_tuple, toxic, r, s, threads=8):
    import za_b200
    ni, na, ptr, var, coeff, inputs, aux = cs_tuple
    ocs = O.CS(ni, na, ptr, var, coeff)
    prm = O.Params.generate(ocs, toxic, threads=threads)
    blob = prm.write()
    rc, exp_proof, etr = prm.create_proof(ocs, inputs, aux, r, s, threads=threads, trace=True)
    assert rc == 0
    pk = za_b200.Parameters.read(ctx, blob, checked=True)
    assert pk.counts() == prm.counts()
    circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
    proof, tr = za_b200.create_proof(ctx, pk, circ, inputs, aux, r, s, trace=True)
    for k in ("a_eval", "b_eval", "c_eval", "a_aux_density", "b_input_density", "b_aux_density", "h_coeffs"):
        assert np.array_equal(tr[k], etr[k]), k
    assert np.array_equal(tr["msm_g1"][:6], etr["msm_g1"][:6]), "a G1 multiexp result differs"
    assert np.array_equal(tr["msm_g2"], etr["msm_g2"]), "a G2 multiexp result differs"
    assert proof == exp_proof
    pub = [int.from_bytes(inputs[i].tobytes(), "little") for i in range(1, ni)]
    assert prm.verify(proof, pub) == 1
    bad = list(pub); bad[0] = (bad[0] + 1) % P.R_MOD
    assert prm.verify(proof, bad) == 0
    return proof, pub, prm


def test_create_proof_example_circuit(ctx):
    """Config 1: /root/reference/example/circuit.za with example/input.json (p=2, q=3, r=6)."""
    import json, os, za_b200
    cs = P.example_factor_circuit()
    ocs = O.CS.from_rows(cs.num_inputs, cs.num_aux, cs.rows)
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "groth16_example.json")))
    toxic = [int(x) for x in g["toxic"]]
    inputs = O.frs_to_np([1, 6]).reshape(-1, 32); aux = O.frs_to_np([2, 3]).reshape(-1, 32)
    proof, pub, prm = _prove_case(ctx, (2, 2, ocs.ptr, ocs.var, ocs.coeff, inputs, aux), toxic, int(g["r"]), int(g["s"]))
    # against the committed closed-form golden proof and its json text
    assert proof.hex() == g["proof_hex"]
    assert za_b200.proof_to_json(proof, pub) == g["proof_json"]


@pytest.mark.parametrize("nc", [5, 100, 1022, (1 << 12) - 2])
def test_create_proof_mul_chain(ctx, nc):
    toxic = [11 + nc, 12, 13, 14, 15 + nc]
    _prove_case(ctx, circuits.mul_chain(nc, x0=7), toxic, r=123456789 + nc, s=987654321)


def test_create_proof_config2_2pow16(ctx):
    """Config 2: multiplication chain with 2^16 - 2 constraints (m = 2^16), end to end on one GPU."""
    nc = (1 << 16) - 2
    toxic = [0x5A410002, 3, 5, 7, 11]
    _prove_case(ctx, circuits.mul_chain_fast(nc, x0=0x5A410002), toxic, r=2 ** 200 + 17, s=2 ** 100 + 3)


@pytest.mark.parametrize("kat", [0, 1, 2])
def test_create_proof_circomlib_babyadd(ctx, kat):
    """Config 3's building block: circomlib BabyAdd (babyjub.circom:23-50, R1CS extracted by hand in tests/circuits.py)
    on the inputs of the reference's own tests (za_test/babyjub.za:4-34); the public outputs are the values those tests
    assert, the witness is full of zeros in the first case, and the proof is bit-identical to the oracle's."""
    args, expect = circuits.BABYADD_KATS[kat]
    cs, out = circuits.babyjub_add(*args)
    assert out == expect
    proof, pub, _ = _prove_case(ctx, cs, [0x5A410003 + kat, 3, 5, 7, 11], r=2 ** 190 + kat, s=2 ** 90 + 7)
    assert tuple(pub) == expect


def test_create_proof_config3_eddsa_mimc(ctx):
    """Config 3: the EdDSA-MiMC verification statement of circomlib (hand-built R1CS, tests/eddsa_circuit.py: 7 429
    constraints, domain 2^13, thousands of boolean witness signals, partial A/B densities) on the reference's own
    signature vector: every intermediate vector and multiexp and the proof bit-identical to the oracle's."""
    from tests import eddsa_circuit as E
    cs, info = E.eddsa_mimc_verifier(**E.KAT)
    proof, pub, prm = _prove_case(ctx, cs, [0x5A410003, 3, 5, 7, 11], r=2 ** 190 + 1, s=2 ** 90 + 7)
    assert pub[-1] == 1234 and prm.verify(proof, pub[:-1] + [1235]) == 0


def test_create_proof_rejects_non_canonical_witness(ctx):
    """A witness element >= r is refused (ZA_ERR_NOT_CANONICAL, first offending index named) — the range check runs on
    the uploaded copy on the GPU — and the same call with the value reduced still proves."""
    import za_b200
    ni, na, ptr, var, coeff, inputs, aux = circuits.mul_chain(100, x0=7)
    ocs = O.CS(ni, na, ptr, var, coeff)
    prm = O.Params.generate(ocs, [3, 5, 7, 11, 13], threads=4)
    pk = za_b200.Parameters.read(ctx, prm.write(), checked=True)
    circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
    bad_aux = aux.copy()
    bad_aux[17] = np.frombuffer(P.R_MOD.to_bytes(32, "little"), np.uint8)          # == r
    bad_aux[60] = 0xFF
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.create_proof(ctx, pk, circ, inputs, bad_aux, 5, 6)
    assert e.value.code == -10 and "aux[17]" in str(e.value)
    bad_in = inputs.copy()
    bad_in[1] = 0xFF
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.create_proof(ctx, pk, circ, bad_in, aux, 5, 6)
    assert e.value.code == -10 and "inputs[1]" in str(e.value)
    rc, exp = prm.create_proof(ocs, inputs, aux, 5, 6, threads=4)
    assert rc == 0 and za_b200.create_proof(ctx, pk, circ, inputs, aux, 5, 6) == exp


def test_pk_load_rejects_bad_streams(ctx):
    import za_b200
    cs = P.example_factor_circuit()
    ocs = O.CS.from_rows(cs.num_inputs, cs.num_aux, cs.rows)
    blob = bytearray(O.Params.generate(ocs, [2, 3, 4, 5, 6]).write())
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.Parameters.read(ctx, bytes(blob[:-10]))
    assert e.value.code == -5
    bad = bytearray(blob); bad[40] ^= 1                       # alpha_g1.y corrupted -> off curve
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.Parameters.read(ctx, bytes(bad))
    assert e.value.code in (-6, -8)
    bad = bytearray(blob); bad[0] |= 0x80                     # compressed flag
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.Parameters.read(ctx, bytes(bad))
    assert e.value.code == -8


def test_product_verifier_accepts_gpu_proof(ctx):
    """generate_verified_proof's self-check (prover.rs:191-200) with the product's own verifier and vk export."""
    import za_b200
    ni, na, ptr, var, coeff, inputs, aux = circuits.mul_chain(50, x0=11)
    ocs = O.CS(ni, na, ptr, var, coeff)
    prm = O.Params.generate(ocs, [21, 22, 23, 24, 25], threads=8)
    pk = za_b200.Parameters.read(ctx, prm.write())
    circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
    proof = za_b200.create_proof(ctx, pk, circ, inputs, aux, 1001, 1002)
    out = O.fr_int(inputs[1])
    vk = pk.vk()
    assert za_b200.verify_proof(vk, proof, [out]) is True
    assert za_b200.verify_proof(vk, proof, [out + 1]) is False
    assert za_b200.verify(za_b200.vk_to_json(vk, ["main.out"]), za_b200.proof_to_json(proof, [out])) is True


def test_synthetic_pk_proof_has_the_closed_form(ctx):
    """Size-independent property used at full size by bench.py: with a synthetic proving key (every base a known
    multiple of the generator) each proof element is (closed-form scalar) * G.  Checked here at 2^16 with the
    oracle's H coefficients; bench.py checks the 2^20 proof against the oracle's proof."""
    import za_b200
    from za_b200 import synthetic
    log_m = 16
    nc = (1 << log_m) - 2
    ni, na, ptr, var, coeff, inputs, aux = synthetic.mul_chain(nc, x0=77)
    cnt = synthetic.pk_counts_for_mul_chain(nc)
    pk = za_b200.Parameters.synthetic(ctx, cnt["ic"], cnt["h"], cnt["l"], cnt["a"], cnt["b_g1"], cnt["b_g2"])
    circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
    r, s = 0xABCDEF0123456789, 0x1122334455667788
    proof, tr = za_b200.create_proof(ctx, pk, circ, inputs, aux, r, s, trace=True)
    # independent H: the oracle's EvaluationDomain pipeline on the oracle's a, b, c evaluation
    ocs = O.CS(ni, na, ptr, var, coeff)
    R = P.R_MOD
    w_in = O.np_to_frs(inputs); w_aux = O.np_to_frs(aux)
    a_ev = np.zeros((nc + ni, 32), np.uint8); b_ev = np.zeros_like(a_ev); c_ev = np.zeros_like(a_ev)
    a_ev[:nc] = aux; b_ev[:nc] = aux                       # A = B = x_k
    c_ev[:nc - 1] = aux[1:]; c_ev[nc - 1] = inputs[1]      # C = x_{k+1} (last: the public output)
    a_ev[nc:] = inputs
    h = O.np_to_frs(O.h_poly(a_ev, b_ev, c_ev, threads=8))
    assert np.array_equal(tr["h_coeffs"], O.h_poly(a_ev, b_ev, c_ev, threads=8))
    mult = lambda q, i: ((q + 1) << 32) + i + 1
    H = sum(mult(0, i) * x for i, x in enumerate(h)) % R
    L = sum(mult(1, i) * x for i, x in enumerate(w_aux)) % R
    Aq = sum(mult(2, i) * x for i, x in enumerate(w_in + w_aux)) % R
    Bq = sum(mult(3, i) * x for i, x in enumerate(w_aux)) % R          # no input occurs in a B row
    a_s = (11 * r + 3 + Aq) % R
    b_s = (11 * s + 5 + Bq) % R
    c_s = (11 * r * s + 3 * s + 5 * r + s * Aq + r * Bq + H + L) % R
    assert O.g1_tuple(proof[:64]) == O.g1_mul(P.G1_GEN, a_s)
    assert O.g2_tuple(proof[64:192]) == O.g2_mul(P.G2_GEN, b_s)
    assert O.g1_tuple(proof[192:]) == O.g1_mul(P.G1_GEN, c_s)


def test_generated_bases_are_the_stated_multiples(ctx):
    import za_b200
    for group, n, first in ((1, 1000, 1), (1, 70, (1 << 32) + 5), (2, 300, 7)):
        b = za_b200.Bases.generate(ctx, group, n, first)
        got = b.download()
        for i in (0, 1, 31, 32, 33, n - 1):
            if group == 1:
                assert O.g1_tuple(got[i]) == O.g1_mul(P.G1_GEN, first + i)
            else:
                assert O.g2_tuple(got[i]) == O.g2_mul(P.G2_GEN, first + i)


def test_staged_prove_equals_single_call(ctx):
    """The one-process-per-GPU stages run on one device with world = 1, 2, 3: same proof."""
    import torch
    import za_b200
    ni, na, ptr, var, coeff, inputs, aux = circuits.mul_chain_fast(9000, x0=13)
    ocs = O.CS(ni, na, ptr, var, coeff)
    prm = O.Params.generate(ocs, [31, 32, 33, 34, 35], threads=8)
    pk = za_b200.Parameters.read(ctx, prm.write())
    circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
    ref = za_b200.create_proof(ctx, pk, circ, inputs, aux, 5, 6)
    rc, exp = prm.create_proof(ocs, inputs, aux, 5, 6, threads=8)
    assert rc == 0 and ref == exp
    wit = torch.from_numpy(np.concatenate([inputs, aux])).cuda()
    m = 1 << circ.info()["log_m"]
    h = torch.zeros((m, 32), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    assert za_b200.create_proof_device(ctx, pk, circ, wit.data_ptr(), 5, 6) == ref
    za_b200.prove_h_device(ctx, circ, wit.data_ptr(), h.data_ptr())
    # the h slices as stores of the last NTT pass (za_ctx_set_h_scatter, what za_prover does towards its peer devices):
    # three destination buffers by index range; each holds exactly its slice of h, the rest is untouched
    bufs = [torch.full((m, 32), 0xEE, dtype=torch.uint8, device="cuda") for _ in range(3)]
    bounds = [(m - 1) // 3, 2 * (m - 1) // 3, m - 1]
    za_b200.set_h_scatter(ctx, [b.data_ptr() for b in bufs], bounds)
    h2 = torch.full((m, 32), 0xEE, dtype=torch.uint8, device="cuda")
    za_b200.prove_h_device(ctx, circ, wit.data_ptr(), h2.data_ptr())
    torch.cuda.synchronize()
    lo = 0
    for b, hi in zip(bufs, bounds):
        assert torch.equal(b[lo:hi], h[lo:hi]) and bool((b[:lo] == 0xEE).all()) and bool((b[hi:m - 1] == 0xEE).all())
        lo = hi
    assert bool((h2 == 0xEE).all())                         # nothing went to d_h
    za_b200.prove_h_device(ctx, circ, wit.data_ptr(), h2.data_ptr())          # the setting was consumed: plain output again
    torch.cuda.synchronize()
    assert torch.equal(h2[:m - 1], h[:m - 1])
    for world in (1, 2, 3):
        parts = [za_b200.prove_msm_partials(ctx, pk, circ, wit.data_ptr(), h.data_ptr(), k, world) for k in range(world)]
        assert za_b200.prove_assemble(pk, np.stack(parts), 5, 6) == ref
    # per-rank fixed-base tables (za_pk_partition): rank 1 of 2 has tables for its range only; every rank still works
    pk.partition(circ, 1, 2)
    parts = [za_b200.prove_msm_partials(ctx, pk, circ, wit.data_ptr(), h.data_ptr(), k, 2) for k in range(2)]
    assert za_b200.prove_assemble(pk, np.stack(parts), 5, 6) == ref
    # weighted ranges (rank 0 also runs the H pipeline) and the two-phase enqueue: the witness multiexps first,
    # the H multiexp once the h scalars are there, one collect
    for world, w0 in ((2, 0.776), (4, 0.55), (8, 0.106), (3, 0.001)):
        parts = []
        for k in range(world):
            pk.partition(circ, k, world, w0)
            za_b200.prove_msm_enqueue(ctx, pk, circ, wit.data_ptr(), h.data_ptr(), k, world, za_b200.MSM_WITNESS)
            za_b200.prove_msm_enqueue(ctx, pk, circ, wit.data_ptr(), h.data_ptr(), k, world, za_b200.MSM_H)
            parts.append(za_b200.prove_msm_collect(ctx))
        assert za_b200.prove_assemble(pk, np.stack(parts), 5, 6) == ref, (world, w0)
    # explicit query ranges (za_pk_partition_ranges) with the plans za_prover makes (za_prover_plan): the witness queries
    # cut as one work line, one contiguous piece per device — emulated device by device on this GPU
    counts = {"h": m - 1, "l": na}
    for world in (2, 3, 5, 8):
        plans = za_b200.prover_plan(circ, world)
        for q in range(5):                                   # the ranges tile every query
            ivs = sorted((lo[q], hi[q]) for lo, hi in plans if hi[q] > lo[q])
            assert ivs and ivs[0][0] == 0 and all(a[1] == b[0] for a, b in zip(ivs, ivs[1:]))
        assert sorted(hi[0] for lo, hi in plans)[-1] == m - 1 and sorted(hi[1] for lo, hi in plans)[-1] == na
        parts = []
        for lo, hi in plans:
            pk.partition_ranges(circ, lo, hi)
            za_b200.prove_msm_enqueue(ctx, pk, circ, wit.data_ptr(), h.data_ptr(), 0, 1, za_b200.MSM_WITNESS)
            za_b200.prove_msm_enqueue(ctx, pk, circ, wit.data_ptr(), h.data_ptr(), 0, 1, za_b200.MSM_H)
            parts.append(za_b200.prove_msm_collect(ctx))
        assert za_b200.prove_assemble(pk, np.stack(parts), 5, 6) == ref, world
    with pytest.raises(za_b200.ZaError):                     # a range beyond its query
        pk.partition_ranges(circ, [0, 0, 0, 0, 0], [m, 0, 0, 0, 0])
    pk.partition_ranges(circ)                                # the whole queries again
    assert za_b200.create_proof_device(ctx, pk, circ, wit.data_ptr(), 5, 6) == ref
    with pytest.raises(za_b200.ZaError):                     # collect without the H multiexp enqueued
        za_b200.prove_msm_enqueue(ctx, pk, circ, wit.data_ptr(), h.data_ptr(), 0, 2, za_b200.MSM_WITNESS)
        za_b200.prove_msm_collect(ctx)
    za_b200.prove_msm_enqueue(ctx, pk, circ, wit.data_ptr(), h.data_ptr(), 0, 2, za_b200.MSM_H)
    za_b200.prove_msm_collect(ctx)                           # drains the slots again


@pytest.mark.parametrize("group,n", [(1, 5000), (1, 1 << 16), (2, 4096), (2, 20000)])
def test_multiexp_with_fixed_base_table(ctx, group, n):
    """The fixed-base table path (one bucket space, 2^(c w) P_i precomputed) against the oracle and against the
    table-free path, including a sub-range of the query (offset) and witness-like scalars."""
    import za_b200
    pts = O.g1_multiples(n) if group == 1 else O.g2_multiples(n)
    s = O.random_frs(n, 500 + n)
    plain = za_b200.Bases(ctx, group, pts)
    tab = za_b200.Bases(ctx, group, pts)
    c = tab.precompute()
    assert c >= 12
    rc, exp = O.multiexp("g1" if group == 1 else "g2", pts, s, threads=8)
    assert rc == 0
    assert za_b200.multiexp(ctx, tab, s) == exp == za_b200.multiexp(ctx, plain, s)
    off, cnt = 1234, n - 2000
    rc, exp2 = O.multiexp("g1" if group == 1 else "g2", pts[off:off + cnt], s[:cnt], threads=8)
    assert za_b200.multiexp(ctx, tab, s[:cnt], offset=off) == exp2
    w = circuits.witness_like(n, 9)
    rc, exp3 = O.multiexp("g1" if group == 1 else "g2", pts, w, threads=8)
    assert za_b200.multiexp(ctx, tab, w) == exp3


def test_helper_prove_from_a_proving_key_file(ctx):
    """helper::prove / generate_verified_proof flow (helper.rs:91-147, prover.rs:139-208) from a proving.key
    container: read_pk -> synthesize -> constraint check -> proof -> self-verify -> proof.json."""
    import json, struct
    import za_b200
    from za_b200 import format as F, helper
    from tests.test_format_host import container
    # za-side circuit: signals one, main.out (public), x0 (private), x1 .. ; constraints x_k * x_k + (-x_{k+1}) = 0
    nc = 300
    R = P.R_MOD
    qeqs = [([(2 + k, 1)], [(2 + k, 1)], [((3 + k) if k + 1 < nc else 1, R - 1)]) for k in range(nc)]
    n_signals = 2 + nc
    is_public = [0, 1] + [0] * nc
    vals = [1, 0] + [0] * nc
    x = 5
    for k in range(nc):
        vals[2 + k] = x; x = x * x % R
    vals[1] = x
    values = O.frs_to_np(vals).reshape(n_signals, 32)
    # proving key made by the oracle's generate_parameters over the synthesized circuit
    tmp = F.read_pk(container(F.EMPTY_AST, qeqs, [], b""))
    syn = F.synthesize(n_signals, is_public, [], tmp.ptr, tmp.sig, tmp.coeff, values)
    ocs = O.CS(syn["num_inputs"], syn["num_aux"], syn["ptr"], syn["var"], syn["coeff"])
    prm = O.Params.generate(ocs, [41, 42, 43, 44, 45], threads=8)
    pk_bytes = container(F.EMPTY_AST, qeqs, [], prm.write())
    key = helper.LoadedKey(ctx, pk_bytes, is_public)
    js, public = helper.prove(key, values, r=77, s=88)
    rc, exp = prm.create_proof(ocs, syn["inputs"], syn["aux"], 77, 88, threads=8)
    assert rc == 0 and js == za_b200.proof_to_json(exp, [x]) and public == [x]
    assert za_b200.verify(za_b200.vk_to_json(key.params.vk(), ["main.out"]), js) is True
    js2, _ = helper.prove(key, values)                       # random r, s: different proof, still valid
    assert js2 != js and za_b200.verify(za_b200.vk_to_json(key.params.vk()), js2) is True
    bad = values.copy(); bad[5, 0] ^= 1
    with pytest.raises(ValueError):                          # "check_constrains_eval_zero failed"
        helper.prove(key, bad)


# ---------------------------------------------------------------- trusted setup (N2): generate_parameters on the GPU
def _setup_case(ctx, cs_tuple, toxic, g1=None, g2=None):
    import za_b200
    ni, na, ptr, var, coeff = cs_tuple[:5]
    ocs = O.CS(ni, na, ptr, var, coeff)
    exp = O.Params.generate(ocs, toxic, g1=g1, g2=g2, threads=8).write()
    circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
    kw = {}
    if g1 is not None: kw["g1"] = g1
    if g2 is not None: kw["g2"] = g2
    got = za_b200.generate_parameters(ctx, circ, *toxic, **kw)
    assert len(got) == len(exp)
    assert got == exp
    return got


def test_generate_parameters_example_circuit(ctx):
    """Config 1's setup: the proving key of example/circuit.za, byte for byte bellman's Parameters::write stream."""
    cs = P.example_factor_circuit()
    ocs = O.CS.from_rows(cs.num_inputs, cs.num_aux, cs.rows)
    _setup_case(ctx, (2, 2, ocs.ptr, ocs.var, ocs.coeff), [2, 3, 4, 5, 6])


@pytest.mark.parametrize("nc", [1, 6, 100, 1022, (1 << 13) - 2])
def test_generate_parameters_mul_chain(ctx, nc):
    _setup_case(ctx, circuits.mul_chain(nc, x0=3), [0x1234567 + nc, 2 ** 200 + 9, 77, P.R_MOD - 2, 2 ** 253 + nc])


def test_generate_parameters_random_generators(ctx):
    """generate_random_parameters (prover.rs:122) draws g1 and g2 as random group elements, not the generators."""
    g1 = O.g1_bytes(O.g1_mul(P.G1_GEN, 0xDEADBEEF12345))
    g2 = O.g2_bytes(O.g2_mul(P.G2_GEN, 0xFEEDFACE98765))
    _setup_case(ctx, circuits.mul_chain(50, x0=11), [9, 8, 7, 6, 5], g1=g1, g2=g2)


def test_generate_parameters_filters_infinity_and_flags_unconstrained(ctx):
    """Variables absent from A (or B) give points at infinity that bellman filters out of a / b_g1 / b_g2; an aux
    variable absent from all three matrices is SynthesisError::UnconstrainedVariable."""
    import za_b200
    A = circuits.AUX
    # inputs: one, out; aux: x, y, z.   rows: x * y = z ; z * one = out     (y never in A, x never in B, out only in C)
    rows = [([(1, A | 0)], [(1, A | 1)], [(1, A | 2)]), ([(1, A | 2)], [(1, 0)], [(1, 1)])]      # (coeff, var)
    ocs = O.CS.from_rows(2, 3, rows)
    blob = _setup_case(ctx, (2, 3, ocs.ptr, ocs.var, ocs.coeff), [3, 5, 7, 11, 13])
    pk = za_b200.Parameters.read(ctx, blob, checked=True)
    cnt = pk.counts()
    assert cnt["a"] < 5 and cnt["b_g1"] < 5 and cnt["b_g1"] == cnt["b_g2"]
    # and a proof made with the GPU-generated key verifies
    inputs = O.frs_to_np([1, 12]).reshape(-1, 32); aux = O.frs_to_np([3, 4, 12]).reshape(-1, 32)
    circ = za_b200.Circuit(ctx, 2, 3, ocs.ptr, ocs.var, ocs.coeff)
    proof = za_b200.create_proof(ctx, pk, circ, inputs, aux, 5, 6)
    assert za_b200.verify_proof(pk.vk(), proof, [12]) is True
    assert za_b200.verify_proof(pk.vk(), proof, [13]) is False
    # aux 3 is never used
    rows2 = rows
    ocs2 = O.CS.from_rows(2, 4, rows2)
    circ2 = za_b200.Circuit(ctx, 2, 4, ocs2.ptr, ocs2.var, ocs2.coeff)
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.generate_parameters(ctx, circ2, 3, 5, 7, 11, 13)
    assert e.value.code == -11
    with pytest.raises(RuntimeError):
        O.Params.generate(ocs2, [3, 5, 7, 11, 13])


def test_generate_parameters_rejects_bad_toxic_values(ctx):
    import za_b200
    ni, na, ptr, var, coeff = circuits.mul_chain(6)[:5]
    circ = za_b200.Circuit(ctx, ni, na, ptr, var, coeff)
    with pytest.raises(za_b200.ZaError):
        za_b200.generate_parameters(ctx, circ, 1, 2, 0, 4, 5)          # gamma = 0 has no inverse
    with pytest.raises(za_b200.ZaError) as e:
        za_b200.generate_parameters(ctx, circ, P.R_MOD, 2, 3, 4, 5)    # not canonical
    assert e.value.code == -10


# ---------------------------------------------------------------- batched-affine pair rounds (msm.cu K5a)
@pytest.fixture
def force_rounds():
    import os
    old = {k: os.environ.get(k) for k in ("ZA_MSM_ROUNDS", "ZA_MSM_C")}
    def set_(rounds, c=None):
        os.environ["ZA_MSM_ROUNDS"] = str(rounds)
        if c is None: os.environ.pop("ZA_MSM_C", None)
        else: os.environ["ZA_MSM_C"] = str(c)
    yield set_
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v


@pytest.mark.parametrize("group", [1, 2])
@pytest.mark.parametrize("rounds,c", [(0, None), (1, None), (2, 4), (3, 5), (5, 3), (6, 2)])
def test_pair_rounds_match_oracle(ctx, force_rounds, group, rounds, c):
    """Every round count gives the same group element as bellman's multiexp (small windows -> deep buckets)."""
    force_rounds(rounds, c)
    for n in (65, 300, 4099):
        if group == 2 and n > 2000: n = 1500
        _msm_case(ctx, group, n, O.random_frs(n, 900 + n + rounds))
    n = 3000 if group == 1 else 1200
    _msm_case(ctx, group, n, circuits.witness_like(n, 50 + rounds))


@pytest.mark.parametrize("group", [1, 2])
@pytest.mark.parametrize("rounds", [1, 2, 4])
def test_pair_rounds_exceptional_pairs(ctx, force_rounds, group, rounds):
    """P + P (doubling), P + (-P) (infinity) and infinity + P inside the pair rounds: every point four times with
    the same scalar, and blocks of opposite points, so whole buckets cancel and results at infinity feed later rounds."""
    import za_b200
    force_rounds(rounds, 4)
    n = 512
    pts = (O.g1_multiples(n) if group == 1 else O.g2_multiples(n)).copy()
    half = pts.shape[1] // 2
    for k in range(1, 4):
        pts[k::4] = pts[0::4]                              # every point four times
    def negated(p):
        q = p.copy()
        if group == 1:
            y = int.from_bytes(q[32:].tobytes(), "little")
            q[32:] = np.frombuffer(((P.Q_MOD - y) % P.Q_MOD).to_bytes(32, "little"), np.uint8)
        else:
            for o in (64, 96):
                y = int.from_bytes(q[o:o + 32].tobytes(), "little")
                q[o:o + 32] = np.frombuffer(((P.Q_MOD - y) % P.Q_MOD).to_bytes(32, "little"), np.uint8)
        return q
    for i in range(0, 128, 4):                             # groups of P, P, -P, -P: the whole group cancels
        pts[i + 2] = negated(pts[i]); pts[i + 3] = negated(pts[i])
    for i in range(128, 192, 4):                           # P, -P, P, P: infinity + P in the next round
        pts[i + 1] = negated(pts[i])
    s = O.random_frs(n, 1234 + rounds)
    for k in range(1, 4):
        s[k::4] = s[0::4]
    bases = za_b200.Bases(ctx, group, pts)
    got = za_b200.multiexp(ctx, bases, s)
    rc, exp = O.multiexp("g1" if group == 1 else "g2", pts, s, threads=8)
    assert rc == 0 and got == exp
    ones = np.zeros((n, 32), np.uint8); ones[:, 0] = 1     # all scalars 1: a single bucket holds everything
    got = za_b200.multiexp(ctx, bases, ones)
    rc, exp = O.multiexp("g1" if group == 1 else "g2", pts, ones, threads=8)
    assert rc == 0 and got == exp


def test_pair_rounds_fixed_base_table(ctx, force_rounds):
    """Table mode (one bucket space, 2^(cw) P_i entries) through the pair rounds, witness-like and uniform scalars."""
    import za_b200
    n = 1 << 15
    pts = O.g1_multiples(n)
    for rounds in (0, 3):
        force_rounds(rounds)
        tab = za_b200.Bases(ctx, 1, pts)
        tab.precompute()
        for s in (O.random_frs(n, 71), circuits.witness_like(n, 72)):
            rc, exp = O.multiexp("g1", pts, s, threads=8)
            assert rc == 0 and za_b200.multiexp(ctx, tab, s) == exp
