"""Context and commit-key handles over the C ABI (include/jf_b200.h)."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from .errors import raise_for_status, InvalidParameters
from .fields import LIMBS, TWO_ADICITY


def _u64p(a: np.ndarray):
    return a.ctypes.data_as(_ffi.c_u64p)


def _as_u64(a, cols: Optional[int] = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if cols is not None and (a.ndim != 2 or a.shape[1] != cols):
        raise InvalidParameters("expected a (n, %d) uint64 limb array, got %r" % (cols, a.shape))
    return a


class Context:
    """One GPU + stream + workspace (`jf_ctx`).  Thread-safe; calls are serialised per context."""

    def __init__(self, device: int = 0):
        self._lib = _ffi.lib()
        h = ctypes.c_void_p()
        rc = self._lib.jf_ctx_create(device, ctypes.byref(h))
        if rc != _ffi.JF_OK:
            raise RuntimeError("jf_ctx_create(device=%d) failed with status %d: no usable CUDA device "
                               "(this library has no CPU fallback)" % (device, rc))
        self._h = h
        self.device = device

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != _ffi.JF_OK:
            raise_for_status(rc, (self._lib.jf_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.jf_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        """Run on an external stream (e.g. `torch.cuda.current_stream().cuda_stream`)."""
        self._check(self._lib.jf_ctx_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def sync(self):
        self._check(self._lib.jf_ctx_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.jf_ctx_launch_count(self._h))

    # -- commit key ---------------------------------------------------------------------------
    def load_srs(self, curve: str, points: np.ndarray, window_bits: int = 0, precompute: bool = True) -> "CommitKey":
        """points: (n, 2L) uint64, affine x || y in Montgomery form, identity = all zero."""
        L = _ffi.CURVE_FQ_LIMBS[curve]
        pts = _as_u64(points, 2 * L)
        h = ctypes.c_void_p()
        self._check(self._lib.jf_srs_load(self._h, _ffi.CURVES[curve], pts.ctypes.data_as(ctypes.c_void_p), pts.shape[0],
                                          16 * L, -1, window_bits, int(precompute), ctypes.byref(h)))
        return CommitKey(self, curve, h)

    def generate_srs_for_testing(self, curve: str, beta: int, n: int, window_bits: int = 0,
                                 precompute: bool = True, first_power: int = 0) -> "CommitKey":
        """`gen_srs_for_testing` with a known beta: key[i] = beta^(first_power + i) * G (srs.rs:118-153)."""
        b = np.array([(beta >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)], dtype=np.uint64)
        h = ctypes.c_void_p()
        self._check(self._lib.jf_srs_generate_for_testing(self._h, _ffi.CURVES[curve], _u64p(b), first_power, n, window_bits,
                                                          int(precompute), ctypes.byref(h)))
        return CommitKey(self, curve, h)

    # -- MSM ------------------------------------------------------------------------------------
    def msm(self, key: "CommitKey", scalars: np.ndarray, base_offset: int = 0, montgomery: bool = False):
        """-> (xy (2L,) uint64 Montgomery affine, is_infinity)."""
        L = _ffi.CURVE_FQ_LIMBS[key.curve]
        s = _as_u64(scalars, 4)
        out = np.zeros(2 * L, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self._check(self._lib.jf_msm(self._h, key._h, base_offset, _u64p(s), s.shape[0], int(montgomery), _u64p(out),
                                     ctypes.byref(inf)))
        return out, bool(inf.value)

    def msm_batch(self, key: "CommitKey", scalar_vectors: Sequence[np.ndarray], base_offsets: Optional[Sequence[int]] = None,
                  montgomery: bool = False):
        L = _ffi.CURVE_FQ_LIMBS[key.curve]
        vecs = [_as_u64(s, 4) for s in scalar_vectors]
        B = len(vecs)
        ptrs = (_ffi.c_u64p * B)(*[_u64p(v) for v in vecs])
        lens = (ctypes.c_size_t * B)(*[v.shape[0] for v in vecs])
        offs = (ctypes.c_size_t * B)(*(base_offsets if base_offsets is not None else [0] * B))
        out = np.zeros((B, 2 * L), dtype=np.uint64)
        inf = (ctypes.c_int * B)()
        self._check(self._lib.jf_msm_batch(self._h, key._h, ptrs, lens, offs, B, int(montgomery), _u64p(out), inf))
        return out, [bool(x) for x in inf]

    def kzg_open(self, key: "CommitKey", polys: Sequence[np.ndarray], points: np.ndarray):
        """`UnivariateKzgPCS::open` for a batch: polys[i] (len_i, 4) Montgomery coefficients, points (batch, 4)
        Montgomery -> (proofs (batch, 2L), infinity flags, evaluations (batch, 4) Montgomery)."""
        L = _ffi.CURVE_FQ_LIMBS[key.curve]
        vecs = [_as_u64(np.asarray(p).reshape(-1, 4), 4) for p in polys]
        B = len(vecs)
        pts = _as_u64(np.asarray(points).reshape(-1, 4), 4)
        if pts.shape[0] != B:
            raise InvalidParameters("poly length %d is different from points length %d" % (B, pts.shape[0]))
        ptrs = (_ffi.c_u64p * B)(*[_u64p(v) for v in vecs])
        lens = (ctypes.c_size_t * B)(*[v.shape[0] for v in vecs])
        out = np.zeros((B, 2 * L), dtype=np.uint64)
        inf = (ctypes.c_int * B)()
        evals = np.zeros((B, 4), dtype=np.uint64)
        self._check(self._lib.jf_kzg_open(self._h, key._h, ptrs, lens, B, _u64p(pts), _u64p(out), inf, _u64p(evals)))
        return out, [bool(x) for x in inf], evals

    def msm_device(self, key: "CommitKey", d_scalars: int, n: int, d_out_xyzz: int, base_offset: int = 0,
                   montgomery: bool = False):
        """Scalars and the XYZZ result stay in HBM (raw device pointers); asynchronous on the stream."""
        self._check(self._lib.jf_msm_device(self._h, key._h, base_offset, ctypes.c_void_p(d_scalars), n, int(montgomery),
                                            ctypes.c_void_p(d_out_xyzz)))

    def msm_combine(self, curve: str, xyzz_parts: np.ndarray):
        """Sum XYZZ partial results (parts, 4L) on the host and normalise -> (xy, is_infinity)."""
        L = _ffi.CURVE_FQ_LIMBS[curve]
        p = _as_u64(xyzz_parts, 4 * L)
        out = np.zeros(2 * L, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self._check(self._lib.jf_msm_combine(self._h, _ffi.CURVES[curve], _u64p(p), p.shape[0], _u64p(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    # -- NTT ------------------------------------------------------------------------------------
    def ntt(self, field: str, data: np.ndarray, log_n: int, inverse: bool = False, coset_offset: Optional[np.ndarray] = None,
            in_len: Optional[int] = None) -> np.ndarray:
        """In place on `data` ((n, 4) or (batch, n, 4) uint64, C-contiguous); returns `data`."""
        if data.dtype != np.uint64 or not data.flags["C_CONTIGUOUS"] or not data.flags["WRITEABLE"]:
            raise InvalidParameters("ntt wants a writable C-contiguous uint64 array")
        n = 1 << log_n if log_n < 63 else 0
        batch = 1 if data.ndim == 2 else data.shape[0]
        too_large = log_n > TWO_ADICITY.get(field, 64)  # the library reports DomainCreationError
        if not too_large and (data.ndim not in (2, 3) or data.shape[-2] != n or data.shape[-1] != 4):
            raise InvalidParameters("ntt: array shape %r does not match log_n=%d" % (data.shape, log_n))
        off = None
        if coset_offset is not None:
            off = np.ascontiguousarray(coset_offset, dtype=np.uint64).reshape(4)
        self._check(self._lib.jf_ntt(self._h, _ffi.FIELDS[field], _u64p(data), n if in_len is None else in_len, log_n,
                                     int(inverse), _u64p(off) if off is not None else None, batch, data.shape[-2]))
        return data

    def ntt_cosets(self, field: str, polys: np.ndarray, log_n: int, offsets: np.ndarray) -> np.ndarray:
        """`get_coset(offsets[r]).fft` of every row of `polys` ((p, in_len, 4), in_len <= 2n) -> (p, rows, n, 4)."""
        polys = np.ascontiguousarray(polys, dtype=np.uint64)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64).reshape(-1, 4)
        if polys.ndim != 3 or polys.shape[-1] != 4:
            raise InvalidParameters("ntt_cosets wants a (polys, in_len, 4) uint64 array")
        n = 1 << log_n if log_n < 63 else 0
        out = np.zeros((polys.shape[0], offsets.shape[0], n, 4), dtype=np.uint64)
        self._check(self._lib.jf_ntt_cosets(self._h, _ffi.FIELDS[field], _u64p(polys), polys.shape[1], polys.shape[1],
                                            polys.shape[0], log_n, 0, _u64p(offsets), offsets.shape[0], _u64p(out)))
        return out

    def intt_cosets(self, field: str, rows: np.ndarray, log_n: int, offsets: np.ndarray) -> np.ndarray:
        """`get_coset(offsets[r]).ifft` in place on `rows` ((p, rows, n, 4)); returns `rows`."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64).reshape(-1, 4)
        n = 1 << log_n if log_n < 63 else 0
        if (rows.dtype != np.uint64 or not rows.flags["C_CONTIGUOUS"] or not rows.flags["WRITEABLE"] or rows.ndim != 4
                or rows.shape[1:] != (offsets.shape[0], n, 4)):
            raise InvalidParameters("intt_cosets wants a writable C-contiguous (polys, rows, n, 4) uint64 array")
        self._check(self._lib.jf_ntt_cosets(self._h, _ffi.FIELDS[field], None, n, n, rows.shape[0], log_n, 1, _u64p(offsets),
                                            offsets.shape[0], _u64p(rows)))
        return rows

    def ntt_device(self, field: str, d_data: int, log_n: int, inverse: bool = False,
                   coset_offset: Optional[np.ndarray] = None, in_len: Optional[int] = None, batch: int = 1,
                   batch_stride: Optional[int] = None):
        n = 1 << log_n
        off = None
        if coset_offset is not None:
            off = np.ascontiguousarray(coset_offset, dtype=np.uint64).reshape(4)
        self._check(self._lib.jf_ntt_device(self._h, _ffi.FIELDS[field], ctypes.c_void_p(d_data),
                                            n if in_len is None else in_len, log_n, int(inverse),
                                            _u64p(off) if off is not None else None, batch,
                                            n if batch_stride is None else batch_stride))

    # -- small helpers ----------------------------------------------------------------------------
    def field_op(self, field: str, op: str, a: np.ndarray, b: Optional[np.ndarray] = None) -> np.ndarray:
        a = _as_u64(a, LIMBS[field])
        bb = _as_u64(b, LIMBS[field]) if b is not None else None
        out = np.empty_like(a)
        self._check(self._lib.jf_field_op(self._h, _ffi.FIELDS[field], _ffi.FIELD_OPS[op], _u64p(a),
                                          _u64p(bb) if bb is not None else None, _u64p(out), a.shape[0]))
        return out

    def fixed_base_mul(self, curve: str, scalars: np.ndarray) -> np.ndarray:
        L = _ffi.CURVE_FQ_LIMBS[curve]
        s = _as_u64(scalars, 4)
        out = np.zeros((s.shape[0], 2 * L), dtype=np.uint64)
        self._check(self._lib.jf_fixed_base_mul(self._h, _ffi.CURVES[curve], _u64p(s), s.shape[0], _u64p(out)))
        return out

    # -- measurement hooks ------------------------------------------------------------------------
    def profile(self, on, dominant_only: bool = False):
        """Per-kernel event timing; dominant_only brackets msm_accumulate / ntt_pass launches only (no cost to the step)."""
        self._check(self._lib.jf_profile_enable(self._h, 2 if (on and dominant_only) else int(bool(on))))

    def profile_collect(self) -> dict:
        """{kernel name: (launches, total_ms)} since the last collect (synchronises)."""
        buf = ctypes.create_string_buffer(1 << 16)
        n = self._lib.jf_profile_collect(self._h, buf, len(buf))
        if n < 0:
            self._check(int(n))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split()
            out[name] = (int(cnt), float(ms))
        return out

    def microbench(self, kind: int) -> float:
        r = ctypes.c_double(0)
        self._check(self._lib.jf_microbench(self._h, kind, ctypes.byref(r)))
        return r.value

    def dev_alloc(self, nbytes: int) -> int:
        p = ctypes.c_void_p()
        self._check(self._lib.jf_dev_alloc(self._h, nbytes, ctypes.byref(p)))
        return p.value

    def dev_free(self, ptr: int):
        self._check(self._lib.jf_dev_free(self._h, ctypes.c_void_p(ptr)))

    def dev_upload(self, dst: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        self._check(self._lib.jf_dev_upload(self._h, ctypes.c_void_p(dst), src.ctypes.data_as(ctypes.c_void_p), src.nbytes))

    def dev_download(self, dst: np.ndarray, src: int):
        self._check(self._lib.jf_dev_download(self._h, dst.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(src), dst.nbytes))


class CommitKey:
    """Device-resident `UnivariateProverParam::powers_of_g` (`jf_srs`)."""

    def __init__(self, ctx: Context, curve: str, handle):
        self.ctx = ctx
        self.curve = curve
        self._h = handle

    def __len__(self) -> int:
        return int(self.ctx._lib.jf_srs_len(self._h))

    @property
    def window_bits(self) -> int:
        return int(self.ctx._lib.jf_srs_window_bits(self._h))

    def read(self, first: int, count: int) -> np.ndarray:
        L = _ffi.CURVE_FQ_LIMBS[self.curve]
        out = np.zeros((count, 2 * L), dtype=np.uint64)
        self.ctx._check(self.ctx._lib.jf_srs_read(self.ctx._h, self._h, first, count, _u64p(out)))
        return out

    def lagrange(self, log_n: int, mask_points: bool = False) -> "CommitKey":
        """The key in the Lagrange basis of the size-2^log_n domain: out[j] = [L_j(beta)] G (`jf_srs_lagrange`), so that a
        commitment is an MSM over a polynomial's values on the domain."""
        h = ctypes.c_void_p()
        self.ctx._check(self.ctx._lib.jf_srs_lagrange(self.ctx._h, self._h, log_n, int(mask_points), ctypes.byref(h)))
        return CommitKey(self.ctx, self.curve, h)

    def free(self):
        if self._h is not None and self.ctx._h:
            self.ctx._lib.jf_srs_free(self.ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
