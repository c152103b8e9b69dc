"""Host-side mirror of the interface the reference's prover crate drives, over the za_b200 C ABI.

Names follow the reference / bellman:
  Parameters.read            <- bellman Parameters::read  (/root/reference/prover/src/groth16/format.rs:285)
  create_proof               <- bellman create_proof / create_random_proof (prover.rs:173)
  multiexp                   <- bellman multiexp (multiexp.rs)
  Context.fft/ifft/coset_fft/icoset_fft <- EvaluationDomain (domain.rs)
  proof_to_json              <- JsonProofAndInput (format.rs:80-128)
Scalars are (n, 32) uint8 arrays of little-endian canonical values, points (n, 64|128) uint8 arrays
(include/za_b200.h).  Everything runs on the GPU; failures raise ZaError.
"""
import ctypes

import numpy as np

from ._lib import check, lib, ZaR1CS, ZaTrace, u8p, u32p

FFT, IFFT, COSET_FFT, ICOSET_FFT = 0, 1, 2, 3
AUX = 0x80000000


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _u8(a, width=None):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if width is not None:
        a = a.reshape(-1, width)
    return a


class Context:
    """One per GPU (za_ctx): owns the stream, the cached EvaluationDomain tables and scratch memory."""

    def __init__(self, device=0):
        h = ctypes.c_void_p()
        check(lib().za_ctx_create(device, ctypes.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib().za_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(lib().za_ctx_set_stream(self.h, ctypes.c_void_p(cuda_stream)))

    def synchronize(self):
        check(lib().za_ctx_synchronize(self.h))

    def launch_count(self):
        return int(lib().za_ctx_launch_count(self.h))

    PROFILE_CLASSES = ("msm_accumulate_g1", "msm_accumulate_g2", "ntt", "msm_sort", "msm_reduce", "r1cs_eval", "h_pointwise", "other")

    def profile(self, on=True):
        check(lib().za_ctx_profile(self.h, 1 if on else 0))

    def profile_read(self):
        """{class: dict(ms=, work=, spans=)} accumulated since the last read (device time, CUDA events)."""
        out = (ctypes.c_double * 24)()
        check(lib().za_ctx_profile_read(self.h, out))
        return {n: dict(ms=out[i], work=out[8 + i], spans=int(out[16 + i])) for i, n in enumerate(self.PROFILE_CLASSES)}

    # ---- EvaluationDomain -------------------------------------------------------------------
    def ntt(self, data, mode):
        """data: (2^k, 32) canonical scalars on the host -> transformed copy."""
        d = _u8(data, 32).copy()
        n = d.shape[0]
        log_n = n.bit_length() - 1
        if n != 1 << log_n:
            raise ValueError("length must be a power of two")
        check(lib().za_ntt(self.h, _p(d), log_n, mode))
        return d

    def fft(self, data):
        return self.ntt(data, FFT)

    def ifft(self, data):
        return self.ntt(data, IFFT)

    def coset_fft(self, data):
        return self.ntt(data, COSET_FFT)

    def icoset_fft(self, data):
        return self.ntt(data, ICOSET_FFT)

    def ntt_device(self, dptr, log_n, mode, batch=1):
        check(lib().za_ntt_device(self.h, ctypes.c_void_p(dptr), log_n, mode, batch))

    def fr_convert_device(self, dptr, n, to_canonical):
        check(lib().za_fr_convert_device(self.h, ctypes.c_void_p(dptr), n, 1 if to_canonical else 0))

    def h_poly(self, a, b, c, checkpoints=False):
        """create_proof's H block.  Returns h ((m-1, 32)) and, if asked, the 8 checkpoint vectors."""
        a, b, c = _u8(a, 32), _u8(b, 32), _u8(c, 32)
        n = a.shape[0]
        m = 1
        while m < n:
            m *= 2
        out = np.zeros((max(m - 1, 1), 32), np.uint8)
        ck = np.zeros((8, m, 32), np.uint8) if checkpoints else None
        check(lib().za_h_poly(self.h, _p(a), _p(b), _p(c), n, _p(out), _p(ck)))
        out = out[:m - 1]
        return (out, ck) if checkpoints else out

    def h_poly_device(self, da, db, dc, log_m):
        check(lib().za_h_poly_device(self.h, ctypes.c_void_p(da), ctypes.c_void_p(db), ctypes.c_void_p(dc), log_m))


class Bases:
    """Device-resident G1 (group=1) or G2 (group=2) affine bases."""

    def __init__(self, ctx, group, points):
        self.ctx, self.group = ctx, group
        pts = _u8(points, 64 if group == 1 else 128)
        h = ctypes.c_void_p()
        check(lib().za_bases_upload(ctx.h, group, _p(pts), pts.shape[0], ctypes.byref(h)))
        self.h = h

    @classmethod
    def generate(cls, ctx, group, n, first_multiple=1):
        """bases[i] = (first_multiple + i) * G, generated on the GPU (synthetic inputs)."""
        self = cls.__new__(cls)
        self.ctx, self.group = ctx, group
        h = ctypes.c_void_p()
        check(lib().za_bases_generate(ctx.h, group, n, first_multiple, ctypes.byref(h)))
        self.h = h
        return self

    def precompute(self):
        """Build the fixed-base table (2^(c w) * P_i); returns the window size c, 0 if none was built."""
        rc = lib().za_bases_precompute(self.ctx.h, self.h)
        if rc < 0:
            check(rc)
        return rc

    def download(self, offset=0, n=None):
        n = len(self) - offset if n is None else n
        out = np.zeros((n, 64 if self.group == 1 else 128), np.uint8)
        check(lib().za_bases_download(self.ctx.h, self.h, offset, n, _p(out)))
        return out

    def __len__(self):
        return int(lib().za_bases_len(self.h))

    def close(self):
        if getattr(self, "h", None):
            lib().za_bases_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def multiexp(ctx, bases, scalars, density=None, offset=0):
    """bellman multiexp: sum scalars[i] * bases[offset + k(i)].  Returns the affine point bytes."""
    s = _u8(scalars, 32)
    d = None if density is None else _u8(density)
    out = np.zeros(64 if bases.group == 1 else 128, np.uint8)
    check(lib().za_multiexp(ctx.h, bases.h, offset, _p(s), s.shape[0], _p(d), _p(out)))
    return out.tobytes()


def multiexp_device(ctx, bases, d_scalars, n, offset=0, partial=False):
    """Scalars already resident on the GPU (raw device pointer).  partial=True returns the XYZZ partial sum."""
    size = (64 if bases.group == 1 else 128) * (2 if partial else 1)
    out = np.zeros(size, np.uint8)
    fn = lib().za_multiexp_partial_device if partial else lib().za_multiexp_device
    check(fn(ctx.h, bases.h, offset, ctypes.c_void_p(d_scalars), n, _p(out)))
    return out.tobytes()


def point_sum(group, partials):
    """Add XYZZ partial sums (multi-GPU combine) and normalise to affine."""
    buf = np.frombuffer(b"".join(partials), np.uint8)
    out = np.zeros(64 if group == 1 else 128, np.uint8)
    check(lib().za_point_sum(group, _p(buf), len(partials), _p(out)))
    return out.tobytes()


class Parameters:
    """bellman Parameters<Bn256> resident on the GPU."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    @classmethod
    def read(cls, ctx, data, checked=True):
        buf = np.frombuffer(bytes(data), np.uint8)
        h = ctypes.c_void_p()
        check(lib().za_pk_load(ctx.h, _p(buf), buf.shape[0], 1 if checked else 0, ctypes.byref(h)))
        return cls(ctx, h)

    @classmethod
    def synthetic(cls, ctx, ic, h, l, a, b_g1, b_g2):
        """Proving key of the given query sizes whose bases are known multiples of the generators
        (include/za_b200.h: za_pk_synthetic).  Not a valid CRS; same work per proof."""
        c = (ctypes.c_uint32 * 6)(ic, h, l, a, b_g1, b_g2)
        hd = ctypes.c_void_p()
        check(lib().za_pk_synthetic(ctx.h, c, ctypes.byref(hd)))
        return cls(ctx, hd)

    def partition(self, circuit, rank, world, rank0_weight=1.0):
        """One process per GPU: keep fixed-base tables only for the point range this rank owns (za_pk_partition).
        rank0_weight < 1: rank 0 (which also runs the H-polynomial pipeline) takes that fraction of an ordinary
        rank's share of the witness multiexps (za_pk_partition_weighted)."""
        w = max(1, min(1000, int(round(rank0_weight * 1000))))
        check(lib().za_pk_partition_weighted(self.ctx.h, self.h, circuit.h, rank, world, w))

    def partition_ranges(self, circuit, lo=None, hi=None):
        """Explicit point ranges [lo[q], hi[q]) of the five queries (H, L, A, B in G1, B in G2) for this device
        (za_pk_partition_ranges); None = the whole queries again."""
        if lo is None:
            check(lib().za_pk_partition_ranges(self.ctx.h, self.h, circuit.h, None, None))
            return
        a = (ctypes.c_uint64 * 5)(*[int(x) for x in lo]); b = (ctypes.c_uint64 * 5)(*[int(x) for x in hi])
        check(lib().za_pk_partition_ranges(self.ctx.h, self.h, circuit.h, a, b))

    def counts(self):
        c = (ctypes.c_uint32 * 6)()
        check(lib().za_pk_counts(self.h, c))
        return dict(zip(("ic", "h", "l", "a", "b_g1", "b_g2"), list(c)))

    def vk(self):
        n_ic = self.counts()["ic"]
        buf = np.zeros(64 * 3 + 128 * 3 + 64 * n_ic, np.uint8)
        check(lib().za_pk_vk(self.h, _p(buf), buf.shape[0]))
        b = buf.tobytes()
        return dict(alpha_g1=b[0:64], beta_g1=b[64:128], beta_g2=b[128:256], gamma_g2=b[256:384], delta_g1=b[384:448],
                    delta_g2=b[448:576], ic=[b[576 + 64 * i:640 + 64 * i] for i in range(n_ic)])

    def close(self):
        if getattr(self, "h", None):
            lib().za_pk_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _r1cs_struct(num_inputs, num_aux, ptr, var, coeff):
    """za_r1cs over numpy arrays; returns (struct, the arrays it points into)."""
    ptr = [np.ascontiguousarray(p, dtype=np.uint32) for p in ptr]
    var = [np.ascontiguousarray(v, dtype=np.uint32) for v in var]
    coeff = [_u8(c, 32) for c in coeff]
    s = ZaR1CS()
    s.num_inputs, s.num_aux, s.num_constraints = num_inputs, num_aux, len(ptr[0]) - 1
    for w in range(3):
        s.ptr[w] = ptr[w].ctypes.data_as(u32p)
        s.var[w] = var[w].ctypes.data_as(u32p)
        s.coeff[w] = coeff[w].ctypes.data_as(u8p)
    return s, (ptr, var, coeff)


class Circuit:
    """The enforce(A, B, C) rows CircomCircuit::synthesize (prover.rs:45-103) hands to bellman, on the GPU.

    ptr/var/coeff: per matrix (A, B, C) CSR arrays; var has bit 31 set for aux variables.
    """

    def __init__(self, ctx, num_inputs, num_aux, ptr, var, coeff):
        self.ctx = ctx
        self.num_inputs, self.num_aux = num_inputs, num_aux
        s, self._keep = _r1cs_struct(num_inputs, num_aux, ptr, var, coeff)
        self.num_constraints = s.num_constraints
        h = ctypes.c_void_p()
        check(lib().za_circuit_upload(ctx.h, ctypes.byref(s), ctypes.byref(h)))
        self.h = h

    def first_unsatisfied(self, inputs, aux):
        """constraints.satisfies_with_signals on the GPU: None if all rows hold, else the first failing row."""
        inputs, aux = _u8(inputs, 32), _u8(aux, 32)
        bad = ctypes.c_int64(-1)
        check(lib().za_circuit_satisfied(self.ctx.h, self.h, _p(inputs), _p(aux), ctypes.byref(bad)))
        return None if bad.value < 0 else int(bad.value)

    def info(self):
        c = (ctypes.c_uint32 * 7)()
        check(lib().za_circuit_info(self.h, c))
        return dict(zip(("num_inputs", "num_aux", "num_constraints", "a_aux_total", "b_input_total", "b_aux_total", "log_m"), list(c)))

    @classmethod
    def from_rows(cls, ctx, num_inputs, num_aux, rows):
        """rows: [(A_terms, B_terms, C_terms)], terms [(coeff_int, var)]"""
        ptr, var, coeff = [], [], []
        for w in range(3):
            p, v, c = [0], [], []
            for row in rows:
                for co, va in row[w]:
                    v.append(va)
                    c.append(int(co).to_bytes(32, "little"))
                p.append(len(v))
            ptr.append(np.array(p, np.uint32))
            var.append(np.array(v, np.uint32))
            coeff.append(np.frombuffer(b"".join(c), np.uint8).reshape(-1, 32) if c else np.zeros((0, 32), np.uint8))
        return cls(ctx, num_inputs, num_aux, ptr, var, coeff)

    def close(self):
        if getattr(self, "h", None):
            lib().za_circuit_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _BorrowedContext(Context):
    """The context of one device of a Prover (owned by the Prover)."""

    def __init__(self, handle, device):
        self.h, self.device = handle, device

    def close(self):
        self.h = None


class Prover:
    """za_prover: create_proof over several GPUs of one box behind one call (include/za_b200.h).
    One host thread per device inside the library; nothing but CUDA runtime calls on the path."""

    def __init__(self, devices):
        self.devices = list(devices)
        arr = (ctypes.c_int * len(self.devices))(*self.devices)
        h = ctypes.c_void_p()
        check(lib().za_prover_create(arr, len(self.devices), ctypes.byref(h)))
        self.h = h
        self.num_inputs = self.num_aux = None

    def close(self):
        if getattr(self, "h", None):
            lib().za_prover_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ctx(self, k=0):
        return _BorrowedContext(ctypes.c_void_p(lib().za_prover_ctx(self.h, k)), self.devices[k])

    def load_pk(self, data, checked=True):
        buf = np.frombuffer(bytes(data), np.uint8)
        check(lib().za_prover_load_pk(self.h, _p(buf), buf.shape[0], 1 if checked else 0))

    def synthetic_pk(self, ic, h, l, a, b_g1, b_g2):
        c = (ctypes.c_uint32 * 6)(ic, h, l, a, b_g1, b_g2)
        check(lib().za_prover_synthetic_pk(self.h, c))

    def set_circuit(self, num_inputs, num_aux, ptr, var, coeff):
        s, keep = _r1cs_struct(num_inputs, num_aux, ptr, var, coeff)
        check(lib().za_prover_set_circuit(self.h, ctypes.byref(s)))
        self.num_inputs, self.num_aux = num_inputs, num_aux

    def vk(self):
        c = (ctypes.c_uint32 * 6)()
        check(lib().za_prover_pk_counts(self.h, c))
        n_ic = c[0]
        buf = np.zeros(64 * 3 + 128 * 3 + 64 * n_ic, np.uint8)
        check(lib().za_prover_vk(self.h, _p(buf), buf.shape[0]))
        b = buf.tobytes()
        return dict(alpha_g1=b[0:64], beta_g1=b[64:128], beta_g2=b[128:256], gamma_g2=b[256:384], delta_g1=b[384:448],
                    delta_g2=b[448:576], ic=[b[576 + 64 * i:640 + 64 * i] for i in range(n_ic)])

    def upload_witness(self, inputs, aux):
        inputs, aux = _u8(inputs, 32), _u8(aux, 32)
        if inputs.shape[0] != self.num_inputs or aux.shape[0] != self.num_aux:
            raise ValueError("witness size does not match the circuit")
        check(lib().za_prover_upload_witness(self.h, _p(inputs), _p(aux)))

    def create_proof(self, inputs, aux, r, s):
        """inputs / aux: host arrays (uploaded inside the call), or None, None: the resident witness again."""
        if inputs is not None:
            inputs, aux = _u8(inputs, 32), _u8(aux, 32)
            if inputs.shape[0] != self.num_inputs or aux.shape[0] != self.num_aux:
                raise ValueError("witness size does not match the circuit")
        proof = np.zeros(256, np.uint8)
        check(lib().za_prover_create_proof(self.h, _p(inputs), _p(aux), _p(_scalar(r)), _p(_scalar(s)), _p(proof)))
        return proof.tobytes()

    def launch_count(self):
        return int(lib().za_prover_launch_count(self.h))

    def info(self):
        v = (ctypes.c_uint64 * 3)()
        check(lib().za_prover_info(self.h, v))
        return dict(h2d_bytes_per_proof=int(v[0]), device0_weight=v[1] / 1000.0, d2h_bytes_per_proof=int(v[2]))


def _domain_size(n):
    m = 1
    while m < n:
        m *= 2
    return m


def create_proof(ctx, params, circuit, inputs, aux, r, s, trace=False):
    """bellman create_proof(circuit, params, r, s).  inputs[0] must be 1.  Returns 256 proof bytes
    (a || b || c) and, with trace=True, the intermediates named in za_trace."""
    inputs, aux = _u8(inputs, 32), _u8(aux, 32)
    if inputs.shape[0] != circuit.num_inputs or aux.shape[0] != circuit.num_aux:
        raise ValueError("witness size does not match the circuit")
    rb = np.frombuffer(int(r).to_bytes(32, "little"), np.uint8)
    sb = np.frombuffer(int(s).to_bytes(32, "little"), np.uint8)
    proof = np.zeros(256, np.uint8)
    tr, bufs = None, {}
    n = circuit.num_constraints + circuit.num_inputs
    m = _domain_size(n)
    if trace:
        bufs = dict(a_eval=np.zeros((n, 32), np.uint8), b_eval=np.zeros((n, 32), np.uint8), c_eval=np.zeros((n, 32), np.uint8),
                    h_coeffs=np.zeros((max(m - 1, 1), 32), np.uint8), msm_g1=np.zeros((7, 64), np.uint8),
                    msm_g2=np.zeros((2, 128), np.uint8), a_aux_density=np.zeros(max(circuit.num_aux, 1), np.uint8),
                    b_input_density=np.zeros(circuit.num_inputs, np.uint8), b_aux_density=np.zeros(max(circuit.num_aux, 1), np.uint8))
        tr = ZaTrace(**{k: v.ctypes.data_as(u8p) for k, v in bufs.items()})
    check(lib().za_create_proof(ctx.h, params.h, circuit.h, _p(inputs), _p(aux), _p(rb), _p(sb), _p(proof),
                                ctypes.byref(tr) if tr is not None else None))
    if trace:
        bufs["h_coeffs"] = bufs["h_coeffs"][:m - 1]
        bufs["a_aux_density"] = bufs["a_aux_density"][:circuit.num_aux]
        bufs["b_aux_density"] = bufs["b_aux_density"][:circuit.num_aux]
        return proof.tobytes(), bufs
    return proof.tobytes()


PARTIALS_BYTES = 1280


def _scalar(x):
    return np.frombuffer(int(x).to_bytes(32, "little"), np.uint8)


def create_proof_device(ctx, params, circuit, d_witness, r, s):
    """create_proof with the witness [inputs | aux] (canonical) already resident on the GPU."""
    proof = np.zeros(256, np.uint8)
    check(lib().za_create_proof_device(ctx.h, params.h, circuit.h, ctypes.c_void_p(d_witness), _p(_scalar(r)), _p(_scalar(s)), _p(proof)))
    return proof.tobytes()


def prove_h_device(ctx, circuit, d_witness, d_h):
    """Stage 1 (one GPU): witness -> H coefficients into d_h (m * 32 bytes of device memory)."""
    check(lib().za_prove_h_device(ctx.h, circuit.h, ctypes.c_void_p(d_witness), ctypes.c_void_p(d_h)))


def set_h_scatter(ctx, outs, his):
    """Destinations of the next prove_h_device on ctx: h[k] goes to outs[j] + 32 k for the first j with k < his[j]
    (device pointers, possibly of peer devices; za_ctx_set_h_scatter)."""
    n = len(outs)
    po = (ctypes.c_void_p * max(n, 1))(*[ctypes.c_void_p(int(o)) for o in outs])
    ph = (ctypes.c_uint64 * max(n, 1))(*[int(x) for x in his])
    check(lib().za_ctx_set_h_scatter(ctx.h, n, po, ph))


def prove_msm_partials(ctx, params, circuit, d_witness, d_h, rank, world):
    """Stage 2 (every GPU): the eight multiexps over this rank's point range -> PARTIALS_BYTES record."""
    out = np.zeros(PARTIALS_BYTES, np.uint8)
    check(lib().za_prove_msm_partials(ctx.h, params.h, circuit.h, ctypes.c_void_p(d_witness), ctypes.c_void_p(d_h), rank, world, _p(out)))
    return out


MSM_WITNESS, MSM_H = 1, 2


def prove_msm_enqueue(ctx, params, circuit, d_witness, d_h, rank, world, which=MSM_WITNESS | MSM_H):
    """Stage 2, asynchronous: enqueue the witness multiexps (L, A, B) and / or the H multiexp of this rank's range."""
    check(lib().za_prove_msm_enqueue(ctx.h, params.h, circuit.h, ctypes.c_void_p(d_witness), ctypes.c_void_p(d_h), rank, world, which))


def prove_msm_collect(ctx):
    """Wait for the five enqueued multiexps -> PARTIALS_BYTES record."""
    out = np.zeros(PARTIALS_BYTES, np.uint8)
    check(lib().za_prove_msm_collect(ctx.h, _p(out)))
    return out


def prove_assemble(params, partials, r, s):
    """Stage 3 (host): add the per-rank partial records and assemble the proof."""
    partials = np.ascontiguousarray(partials, dtype=np.uint8).reshape(-1, PARTIALS_BYTES)
    proof = np.zeros(256, np.uint8)
    check(lib().za_prove_assemble(params.h, _p(partials), partials.shape[0], _p(_scalar(r)), _p(_scalar(s)), _p(proof)))
    return proof.tobytes()


def prover_plan(circuit, n_devices):
    """The query ranges za_prover gives each device: list of (lo[5], hi[5]) (za_prover_plan)."""
    lo = (ctypes.c_uint64 * (5 * n_devices))(); hi = (ctypes.c_uint64 * (5 * n_devices))()
    check(lib().za_prover_plan(circuit.h, n_devices, lo, hi))
    return [(list(lo[5 * k:5 * k + 5]), list(hi[5 * k:5 * k + 5])) for k in range(n_devices)]


def prover_plan_counts(counts, domain, n_devices):
    """prover_plan from the five query lengths (H, L, A, B in G1, B in G2) and the domain size; host only."""
    c = (ctypes.c_uint64 * 5)(*[int(x) for x in counts])
    lo = (ctypes.c_uint64 * (5 * n_devices))(); hi = (ctypes.c_uint64 * (5 * n_devices))()
    check(lib().za_prover_plan_counts(c, int(domain), n_devices, lo, hi))
    return [(list(lo[5 * k:5 * k + 5]), list(hi[5 * k:5 * k + 5])) for k in range(n_devices)]


def share(count, rank, world):
    """The contiguous [lo, hi) share of `count` items that rank owns (same rule as the library)."""
    return count * rank // world, count * (rank + 1) // world


def share_weighted(count, rank, world, rank0_weight=1.0):
    """[lo, hi) of `count` items for `rank` when rank 0 takes rank0_weight of an ordinary share (za_share_weighted)."""
    w = max(1, min(1000, int(round(rank0_weight * 1000))))
    lo, hi = ctypes.c_uint64(0), ctypes.c_uint64(0)
    check(lib().za_share_weighted(count, rank, world, w, ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value


def imad_peak(ctx):
    """Measured integer multiply-add throughput of the device, IMAD/s."""
    v = ctypes.c_double(0)
    check(lib().za_imad_peak(ctx.h, ctypes.byref(v)))
    return v.value


def vk_bytes(vk):
    """vk dict (Parameters.vk()) or raw bytes -> (bytes, n_ic) in the za_pk_vk layout."""
    if isinstance(vk, dict):
        b = vk["alpha_g1"] + vk["beta_g1"] + vk["beta_g2"] + vk["gamma_g2"] + vk["delta_g1"] + vk["delta_g2"] + b"".join(vk["ic"])
    else:
        b = bytes(vk)
    return b, (len(b) - 576) // 64


def verify_proof(vk, proof, public_inputs):
    """bellman verify_proof(prepare_verifying_key(vk), proof, public_inputs) -> bool (host-side pairing check)."""
    b, n_ic = vk_bytes(vk)
    vb = np.frombuffer(b, np.uint8)
    p = np.frombuffer(bytes(proof), np.uint8)
    pi = np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in public_inputs), np.uint8) if public_inputs else None
    ok = ctypes.c_int(0)
    check(lib().za_verify_proof(_p(vb), n_ic, _p(p), _p(pi), len(public_inputs), ctypes.byref(ok)))
    return bool(ok.value)


def vk_to_json(vk, input_names=()):
    """JsonVerifyingKey::to_json (format.rs:130-193)."""
    b, n_ic = vk_bytes(vk)
    vb = np.frombuffer(b, np.uint8)
    names = (ctypes.c_char_p * max(len(input_names), 1))(*[n.encode() for n in input_names])
    size = 4096 + 160 * n_ic + sum(len(n) + 8 for n in input_names)
    buf = ctypes.create_string_buffer(size)
    check(lib().za_vk_to_json(_p(vb), n_ic, names, len(input_names), buf, size))
    return buf.value.decode()


def vk_to_solidity(vk, input_names=(), contract_template=None):
    """generate_solidity (ethereum.rs:216-261).  contract_template=None: the library's own contract text."""
    b, n_ic = vk_bytes(vk)
    vb = np.frombuffer(b, np.uint8)
    names = (ctypes.c_char_p * max(len(input_names), 1))(*[n.encode() for n in input_names])
    tmpl = contract_template.encode() if contract_template is not None else None
    need = ctypes.c_size_t(0)
    probe = ctypes.create_string_buffer(1)
    lib().za_vk_to_solidity(_p(vb), n_ic, names, len(input_names), tmpl, probe, 1, ctypes.byref(need))
    if need.value == 0:
        check(-2)
    buf = ctypes.create_string_buffer(need.value)
    check(lib().za_vk_to_solidity(_p(vb), n_ic, names, len(input_names), tmpl, buf, need.value, None))
    return buf.value.decode()


def verify(vk_json, proof_json):
    """helper::verify (helper.rs:149-158): both arguments are JSON text -> bool."""
    ok = ctypes.c_int(0)
    check(lib().za_verify_json(vk_json.encode(), proof_json.encode(), ctypes.byref(ok)))
    return bool(ok.value)


G1_GENERATOR = (1).to_bytes(32, "little") + (2).to_bytes(32, "little")
G2_GENERATOR = b"".join(int(x).to_bytes(32, "little") for x in (
    10857046999023057135944570762232829481370756359578518086990519993285655852781,
    11559732032986387107991004021392285783925812861821192530917403151452391805634,
    8495653923123431417604973247489272438418190587263600148770280649306958101930,
    4082367875863433681332203403145435568316851327593401208105741076214120093531))   # ethereum.rs:28-31


def generate_parameters(ctx, circuit, alpha, beta, gamma, delta, tau, g1=G1_GENERATOR, g2=G2_GENERATOR):
    """bellman generate_parameters with explicit toxic values (generate_random_parameters, prover.rs:122, draws
    them — and g1, g2 — from the RNG).  Returns the bytes bellman's Parameters::write produces."""
    size = int(lib().za_parameters_max_size(circuit.h))
    out = np.zeros(size, np.uint8)
    n = ctypes.c_size_t(0)
    sc = [_scalar(x) for x in (alpha, beta, gamma, delta, tau)]
    b1 = np.frombuffer(bytes(g1), np.uint8)
    b2 = np.frombuffer(bytes(g2), np.uint8)
    check(lib().za_generate_parameters(ctx.h, circuit.h, *[_p(x) for x in sc], _p(b1), _p(b2), _p(out), size, ctypes.byref(n)))
    return out[:n.value].tobytes()


def proof_to_json(proof, public_inputs):
    """JsonProofAndInput (format.rs:80-128).  public_inputs: ints (decimal strings in the JSON)."""
    p = np.frombuffer(bytes(proof), np.uint8)
    pi = np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in public_inputs), np.uint8) if public_inputs else None
    size = 1024 + 80 * len(public_inputs)
    buf = ctypes.create_string_buffer(size)
    check(lib().za_proof_to_json(_p(p), _p(pi), len(public_inputs), buf, size))
    return buf.value.decode()
