"""ctypes loader for libza_b200.so (the C ABI in include/za_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and if no CUDA
device is present every compute call fails with ZA_ERR_CUDA.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZA_B200_SO: development override (a variant build from `python -m za_b200.build --variant`); it must still be a build
# of this library — every symbol of include/za_b200.h is resolved below
SO_PATH = os.environ.get("ZA_B200_SO") or os.path.join(_HERE, "libza_b200.so")

ZA_OK = 0
ERR_NAMES = {
    0: "ZA_OK", -1: "ZA_ERR_CUDA", -2: "ZA_ERR_INVALID", -3: "ZA_ERR_UNEXPECTED_IDENTITY",
    -4: "ZA_ERR_POLY_DEGREE_TOO_LARGE", -5: "ZA_ERR_IO", -6: "ZA_ERR_NOT_ON_CURVE", -7: "ZA_ERR_NOT_IN_SUBGROUP",
    -8: "ZA_ERR_BAD_ENCODING", -9: "ZA_ERR_BUFFER_TOO_SMALL", -10: "ZA_ERR_NOT_CANONICAL", -11: "ZA_ERR_UNCONSTRAINED_VARIABLE",
}


class ZaError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {text}")
        self.code = code


u8p = ctypes.POINTER(ctypes.c_uint8)
u32p = ctypes.POINTER(ctypes.c_uint32)
vp = ctypes.c_void_p
sz = ctypes.c_size_t
ci = ctypes.c_int


class ZaR1CS(ctypes.Structure):
    _fields_ = [("num_inputs", ctypes.c_uint32), ("num_aux", ctypes.c_uint32), ("num_constraints", ctypes.c_uint32),
                ("ptr", u32p * 3), ("var", u32p * 3), ("coeff", u8p * 3)]


class ZaTrace(ctypes.Structure):
    _fields_ = [(n, u8p) for n in ("a_eval", "b_eval", "c_eval", "h_coeffs", "msm_g1", "msm_g2",
                                   "a_aux_density", "b_input_density", "b_aux_density")]


# every symbol include/za_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "za_last_error": (ctypes.c_char_p, []),
    "za_version": (ci, []),
    "za_device_count": (ci, []),
    "za_ctx_create": (ci, [ci, ctypes.POINTER(vp)]),
    "za_ctx_destroy": (None, [vp]),
    "za_ctx_set_stream": (ci, [vp, vp]),
    "za_ctx_use_own_stream": (ci, [vp]),
    "za_ctx_synchronize": (ci, [vp]),
    "za_ctx_launch_count": (ctypes.c_uint64, [vp]),
    "za_ctx_profile": (ci, [vp, ci]),
    "za_ctx_profile_read": (ci, [vp, vp]),
    "za_ntt": (ci, [vp, vp, ci, ci]),
    "za_ntt_device": (ci, [vp, vp, ci, ci, ci]),
    "za_fr_convert_device": (ci, [vp, vp, sz, ci]),
    "za_h_poly": (ci, [vp, vp, vp, vp, sz, vp, vp]),
    "za_h_poly_device": (ci, [vp, vp, vp, vp, ci]),
    "za_bases_upload": (ci, [vp, ci, vp, sz, ctypes.POINTER(vp)]),
    "za_bases_free": (None, [vp]),
    "za_bases_len": (sz, [vp]),
    "za_multiexp": (ci, [vp, vp, sz, vp, sz, vp, vp]),
    "za_multiexp_device": (ci, [vp, vp, sz, vp, sz, vp]),
    "za_multiexp_partial_device": (ci, [vp, vp, sz, vp, sz, vp]),
    "za_point_sum": (ci, [ci, vp, sz, vp]),
    "za_pk_load": (ci, [vp, vp, sz, ci, ctypes.POINTER(vp)]),
    "za_pk_free": (None, [vp]),
    "za_pk_counts": (ci, [vp, vp]),
    "za_pk_vk": (ci, [vp, vp, sz]),
    "za_circuit_upload": (ci, [vp, vp, ctypes.POINTER(vp)]),
    "za_circuit_free": (None, [vp]),
    "za_create_proof": (ci, [vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "za_create_proof_device": (ci, [vp, vp, vp, vp, vp, vp, vp]),
    "za_circuit_satisfied": (ci, [vp, vp, vp, vp, ctypes.POINTER(ctypes.c_int64)]),
    "za_circuit_info": (ci, [vp, vp]),
    "za_pk_partition": (ci, [vp, vp, vp, ci, ci]),
    "za_pk_partition_weighted": (ci, [vp, vp, vp, ci, ci, ctypes.c_uint32]),
    "za_pk_partition_ranges": (ci, [vp, vp, vp, vp, vp]),
    "za_prover_plan": (ci, [vp, ci, vp, vp]),
    "za_prover_plan_counts": (ci, [vp, ctypes.c_uint64, ci, vp, vp]),
    "za_share_weighted": (ci, [ctypes.c_uint64, ci, ci, ctypes.c_uint32, vp, vp]),
    "za_prove_msm_enqueue": (ci, [vp, vp, vp, vp, vp, ci, ci, ci]),
    "za_prove_msm_collect": (ci, [vp, vp]),
    "za_prove_h_device": (ci, [vp, vp, vp, vp]),
    "za_ctx_set_h_scatter": (ci, [vp, ci, vp, vp]),
    "za_prove_msm_partials": (ci, [vp, vp, vp, vp, vp, ci, ci, vp]),
    "za_prove_assemble": (ci, [vp, vp, ci, vp, vp, vp]),
    "za_bases_generate": (ci, [vp, ci, sz, ctypes.c_uint64, ctypes.POINTER(vp)]),
    "za_bases_download": (ci, [vp, vp, sz, sz, vp]),
    "za_bases_precompute": (ci, [vp, vp]),
    "za_pk_synthetic": (ci, [vp, vp, ctypes.POINTER(vp)]),
    "za_imad_peak": (ci, [vp, ctypes.POINTER(ctypes.c_double)]),
    "za_verify_proof": (ci, [vp, sz, vp, vp, sz, ctypes.POINTER(ci)]),
    "za_vk_to_json": (ci, [vp, sz, ctypes.POINTER(ctypes.c_char_p), sz, ctypes.c_char_p, sz]),
    "za_verify_json": (ci, [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ci)]),
    "za_vk_to_solidity": (ci, [vp, sz, ctypes.POINTER(ctypes.c_char_p), sz, ctypes.c_char_p, ctypes.c_char_p, sz, ctypes.POINTER(sz)]),
    "za_pkfile_scan": (ci, [vp, sz, vp, vp, vp, vp]),
    "za_pkfile_read": (ci, [vp, sz, vp, vp, vp, vp]),
    "za_pkfile_write": (ci, [vp, sz, ctypes.c_uint32, vp, vp, vp, vp, ctypes.c_uint32, vp, sz, vp, sz, vp]),
    "za_synthesize": (ci, [ctypes.c_uint32, vp, vp, ctypes.c_uint32, ctypes.c_uint32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "za_parameters_max_size": (sz, [vp]),
    "za_generate_parameters": (ci, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, ctypes.POINTER(sz)]),
    "za_proof_to_json": (ci, [vp, vp, sz, ctypes.c_char_p, sz]),
    "za_prover_create": (ci, [ctypes.POINTER(ci), ci, ctypes.POINTER(vp)]),
    "za_prover_destroy": (None, [vp]),
    "za_prover_device_count": (ci, [vp]),
    "za_prover_ctx": (vp, [vp, ci]),
    "za_prover_load_pk": (ci, [vp, vp, sz, ci]),
    "za_prover_synthetic_pk": (ci, [vp, vp]),
    "za_prover_set_circuit": (ci, [vp, vp]),
    "za_prover_vk": (ci, [vp, vp, sz]),
    "za_prover_pk_counts": (ci, [vp, vp]),
    "za_prover_upload_witness": (ci, [vp, vp, vp]),
    "za_prover_create_proof": (ci, [vp, vp, vp, vp, vp, vp]),
    "za_prover_launch_count": (ctypes.c_uint64, [vp]),
    "za_prover_info": (ci, [vp, vp]),
}

_lib = None


def lib():
    """Load libza_b200.so; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: build it with `python -m za_b200.build` "
                              "(za_b200 has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)      # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != ZA_OK:
        raise ZaError(rc, lib().za_last_error().decode("utf-8", "replace"))
