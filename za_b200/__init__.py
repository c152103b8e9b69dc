"""za_b200 — B200-native Groth16 proving backend for za (the bellman-bn128 hot path only).

The product is the CUDA library `libza_b200.so` (C ABI in include/za_b200.h); this package is the thin
host-side mirror of the reference's prover interface over it (za_b200.groth16).  There is no CPU,
PyTorch or oracle fallback anywhere in this package.
"""
from . import _lib  # noqa: F401
from ._lib import ZaError  # noqa: F401
from .groth16 import (Context, Bases, Parameters, Circuit, Prover, create_proof, create_proof_device, multiexp,  # noqa: F401
                      multiexp_device, point_sum, prove_h_device, set_h_scatter, prove_msm_partials, prove_msm_enqueue, prove_msm_collect, MSM_WITNESS, MSM_H, prove_assemble, share, share_weighted, prover_plan, prover_plan_counts,
                      imad_peak, proof_to_json, verify_proof, vk_to_json, vk_to_solidity, verify, generate_parameters, PARTIALS_BYTES, FFT, IFFT, COSET_FFT, ICOSET_FFT)
