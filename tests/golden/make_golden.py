"""Generates the committed golden vectors from the independent python-integer reference (tests/pyref.py).

Run from the repo root:  python tests/golden/make_golden.py
Nothing here uses the C oracle or the CUDA product; the vectors pin both.
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import pyref as P  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rng = random.Random(0x5A410001)


def hex_le(x):
    return int(x).to_bytes(32, "little").hex()


def g1_hex(p):
    return "00" * 64 if p is None else hex_le(p[0]) + hex_le(p[1])


def g2_hex(p):
    return "00" * 128 if p is None else hex_le(p[0][0]) + hex_le(p[0][1]) + hex_le(p[1][0]) + hex_le(p[1][1])


def field_vectors():
    out = {}
    for name, p in (("fr", P.R_MOD), ("fq", P.Q_MOD)):
        rows = []
        edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2]
        vals = edge + [rng.randrange(p) for _ in range(20)]
        for a in vals:
            b = rng.choice(vals)
            rows.append(dict(a=str(a), b=str(b), add=str((a + b) % p), sub=str((a - b) % p), mul=str(a * b % p),
                             inv=str(pow(a, -1, p) if a else 0), neg=str((-a) % p)))
        out[name] = rows
    # KATs the reference tree holds (compiler/src/algebra/fs.rs:429-436: 1/2 * 6 == 3)
    out["fs_rs_kat"] = dict(half_times_six=str(pow(2, -1, P.R_MOD) * 6 % P.R_MOD))
    return out


def curve_vectors():
    ks = [1, 2, 3, rng.randrange(P.R_MOD), P.R_MOD - 1]
    return dict(g1=[dict(k=str(k), p=g1_hex(P.g1_mul(P.G1_GEN, k))) for k in ks],
                g2=[dict(k=str(k), p=g2_hex(P.g2_mul(P.G2_GEN, k))) for k in ks])


def ntt_vectors():
    out = []
    for log_n in (0, 1, 2, 3, 5):
        n = 1 << log_n
        v = [rng.randrange(P.R_MOD) for _ in range(n)]
        out.append(dict(log_n=log_n, input=[str(x) for x in v], fft=[str(x) for x in P.fft(v)], ifft=[str(x) for x in P.ifft(v)],
                        coset_fft=[str(x) for x in P.coset_fft(v)], icoset_fft=[str(x) for x in P.icoset_fft(v)]))
    a, b, c = ([rng.randrange(P.R_MOD) for _ in range(6)] for _ in range(3))
    return dict(transforms=out, h_poly=dict(a=[str(x) for x in a], b=[str(x) for x in b], c=[str(x) for x in c],
                                            h=[str(x) for x in P.h_poly(a, b, c)]))


def msm_vectors():
    out = []
    for n in (3, 40):
        bases = [P.g1_mul(P.G1_GEN, i + 1) for i in range(n)]
        b2 = [P.g2_mul(P.G2_GEN, i + 1) for i in range(n)]
        sc = [rng.randrange(P.R_MOD) for _ in range(n)]
        sc[0] = 0
        sc[1] = 1
        out.append(dict(n=n, scalars=[str(x) for x in sc], g1=g1_hex(P.msm(P.Fq1Ops, bases, sc)), g2=g2_hex(P.msm(P.Fq2Ops, b2, sc))))
    return out


def groth16_example():
    """Config 1: example/circuit.za, example/input.json; toxic waste, r and s fixed."""
    cs = P.example_factor_circuit()
    toxic = [rng.randrange(1, P.R_MOD) for _ in range(5)]
    r, s = rng.randrange(P.R_MOD), rng.randrange(P.R_MOD)
    proof, vk = P.groth16_closed_form(cs, [1, 6], [2, 3], toxic, r, s)
    proof_hex = g1_hex(proof[0]) + g2_hex(proof[1]) + g1_hex(proof[2])
    hx = lambda v: "0x%064x" % v
    js = {"a": [hx(proof[0][0]), hx(proof[0][1])],
          "b": [[hx(proof[1][0][0]), hx(proof[1][0][1])], [hx(proof[1][1][0]), hx(proof[1][1][1])]],
          "c": [hx(proof[2][0]), hx(proof[2][1])], "public_inputs": ["6"]}
    return dict(toxic=[str(t) for t in toxic], r=str(r), s=str(s), inputs=["1", "6"], aux=["2", "3"], proof_hex=proof_hex,
                proof_json=json.dumps(js, separators=(",", ":")),
                vk=dict(alpha_g1=g1_hex(vk["alpha_g1"]), beta_g1=g1_hex(vk["beta_g1"]), beta_g2=g2_hex(vk["beta_g2"]),
                        gamma_g2=g2_hex(vk["gamma_g2"]), delta_g1=g1_hex(vk["delta_g1"]), delta_g2=g2_hex(vk["delta_g2"]),
                        ic=[g1_hex(p) for p in vk["ic"]]))


if __name__ == "__main__":
    for name, fn in (("field", field_vectors), ("curve", curve_vectors), ("ntt", ntt_vectors), ("msm", msm_vectors),
                     ("groth16_example", groth16_example)):
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
