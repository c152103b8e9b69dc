"""Mirror of the reference's prover helper (/root/reference/prover/src/groth16/helper.rs:91-158) over the GPU
backend, for callers that already hold the witness: za's front-end (parser / evaluator) is out of scope,
so `prove` takes the signal values instead of re-evaluating the circuit's AST.
"""
import secrets

import numpy as np

from . import format as fmt
from .groth16 import Circuit, Parameters, create_proof, proof_to_json, verify, verify_proof  # noqa: F401

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617

_VERBOSE = False


def verbose(on=True):
    """binding verbose(1) / the `info!` spans of prover.rs:121-132,150-205: the stage times of setup and prove go to
    stderr ("Setup time", "Proving key write time", "Proving key read time", "Constraint check time ... for N constraint",
    "Proof generation time", "Proof verification time"), and ZA_DEBUG_TIMELINE switches on the library's per-multiexp
    timeline (the BELLMAN_VERBOSE analogue) for contexts created afterwards."""
    global _VERBOSE
    _VERBOSE = bool(on)
    import os
    if on:
        os.environ["ZA_DEBUG_TIMELINE"] = "1"
    else:
        os.environ.pop("ZA_DEBUG_TIMELINE", None)


class _Span:
    def __init__(self, text):
        self.text = text

    def __enter__(self):
        import time
        self.t = time.perf_counter()
        return self

    def __exit__(self, *exc):
        import sys
        import time
        if _VERBOSE and exc[0] is None:
            print(f"[za] {self.text}: {(time.perf_counter() - self.t) * 1e3:.3f} ms", file=sys.stderr)
        return False


class LoadedKey:
    """read_pk + synthesize + upload, done once per proving.key (the reference re-reads it on every prove)."""

    def __init__(self, ctx, pk_bytes, is_public, checked=True):
        self.ctx = ctx
        with _Span("Proving key read time"):
            self.file = fmt.read_pk(pk_bytes)
            self.is_public = np.ascontiguousarray(is_public, np.uint8)
            self.n_signals = len(self.is_public)
            s = fmt.synthesize(self.n_signals, self.is_public, self.file.ignore_signals, self.file.ptr, self.file.sig, self.file.coeff)
            self.var_of_signal = s["var_of_signal"]
            self.circuit = Circuit(ctx, s["num_inputs"], s["num_aux"], s["ptr"], s["var"], s["coeff"])
            self.params = Parameters.read(ctx, self.file.params, checked=checked)          # format.rs:285


def prove(key, values, r=None, s=None, self_verify=True):
    """generate_verified_proof (prover.rs:139-208): constraint check, create_(random_)proof, public inputs in signal
    order, self-verification, proof.json.  values: (n_signals, 32) canonical signal values.  Returns
    (proof_json, [public input ints])."""
    values = np.ascontiguousarray(values, np.uint8).reshape(key.n_signals, 32)
    inputs = np.zeros((key.circuit.num_inputs, 32), np.uint8)
    aux = np.zeros((key.circuit.num_aux, 32), np.uint8)
    v = key.var_of_signal
    live = v != 0xFFFFFFFF
    is_aux = (v & 0x80000000) != 0
    aux[(v[live & is_aux] & 0x7FFFFFFF)] = values[live & is_aux]
    inputs[v[live & ~is_aux]] = values[live & ~is_aux]
    inputs[0] = 0
    inputs[0, 0] = 1
    with _Span(f"Constraint check time for {key.circuit.num_constraints} constraint"):
        bad = key.circuit.first_unsatisfied(inputs, aux)
    if bad is not None:
        raise ValueError(f"check_constrains_eval_zero failed: constraint {bad}")          # prover.rs:155-157
    r = secrets.randbelow(R_MOD) if r is None else r                                     # create_random_proof
    s = secrets.randbelow(R_MOD) if s is None else s
    with _Span("Proof generation time"):
        proof = create_proof(key.ctx, key.params, key.circuit, inputs, aux, r, s)
    with _Span("Proof verification time"):
        public = [int.from_bytes(inputs[i].tobytes(), "little") for i in range(1, key.circuit.num_inputs)]   # prover.rs:181-189
        if self_verify and not verify_proof(key.params.vk(), proof, public):                  # prover.rs:191-200
            raise RuntimeError("proof does not verify")
        out = proof_to_json(proof, public)
    return out, public
