"""ctypes front-end of oracle/libjf_oracle.so (TEST INFRASTRUCTURE ONLY -- see jf_oracle.c).

Arrays are numpy uint64, little-endian limbs, shape (n, limbs) for field elements and
(n, 2*limbs) for affine points (x || y, identity = all zero).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FIELD_ID = {"bn254_fr": 0, "bn254_fq": 1, "bls12_381_fr": 2, "bls12_381_fq": 3}
CURVE_ID = {"bn254": 0, "bls12_381": 1}
FIELD_LIMBS = {"bn254_fr": 4, "bn254_fq": 4, "bls12_381_fr": 4, "bls12_381_fq": 6}
CURVE_FQ_LIMBS = {"bn254": 4, "bls12_381": 6}
OPS = {"mul": 0, "add": 1, "sub": 2, "sqr": 3, "inv": 4, "to_mont": 5, "from_mont": 6, "neg": 7}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libjf_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("jf_oracle.c", "mont_tmpl.h", "ec_tmpl.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        u64p = ctypes.POINTER(ctypes.c_uint64)
        _LIB.jfo_field_op.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, u64p, ctypes.c_size_t]
        _LIB.jfo_ntt.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_int, u64p,
                                 ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]
        _LIB.jfo_msm.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, u64p, ctypes.c_size_t, u64p,
                                 ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        _LIB.jfo_msm_windows.argtypes = [ctypes.c_size_t, ctypes.c_int]
        _LIB.jfo_fixed_base_mul.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, u64p, ctypes.c_int]
        _LIB.jfo_gen_srs.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, u64p, ctypes.c_int]
        _LIB.jfo_poly_eval.argtypes = [ctypes.c_int, u64p, ctypes.c_size_t, u64p, u64p]
        _LIB.jfo_random_field_elems.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int, u64p]
    return _LIB


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def max_threads() -> int:
    return lib().jfo_max_threads()


def field_op(field: str, op: str, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    bp = _p(np.ascontiguousarray(b, dtype=np.uint64)) if b is not None else None
    rc = lib().jfo_field_op(FIELD_ID[field], OPS[op], _p(a), bp, _p(out), a.shape[0])
    assert rc == 0
    return out


def ntt(field: str, data: np.ndarray, log_n: int, inverse: bool = False, coset_offset: np.ndarray | None = None,
        in_len: int | None = None, threads: int = 0) -> np.ndarray:
    """data: (batch, n, 4) or (n, 4) Montgomery limbs; returns a transformed copy."""
    x = np.array(data, dtype=np.uint64, order="C", copy=True)
    n = 1 << log_n
    batch = 1 if x.ndim == 2 else x.shape[0]
    assert x.shape[-2] == n
    off = _p(np.ascontiguousarray(coset_offset, dtype=np.uint64)) if coset_offset is not None else None
    rc = lib().jfo_ntt(FIELD_ID[field], _p(x), n if in_len is None else in_len, log_n, int(inverse), off, batch, n,
                       threads)
    if rc != 0:
        raise ValueError("jfo_ntt failed: %d" % rc)
    return x


def msm(curve: str, points: np.ndarray, scalars: np.ndarray, threads: int = 0):
    """points (n, 2L) Montgomery affine, scalars (m, 4) canonical -> (xy (2L,), is_infinity)."""
    L = CURVE_FQ_LIMBS[curve]
    points = np.ascontiguousarray(points, dtype=np.uint64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    out = np.zeros(2 * L, dtype=np.uint64)
    inf = ctypes.c_int(0)
    rc = lib().jfo_msm(CURVE_ID[curve], _p(points), points.shape[0], _p(scalars), scalars.shape[0], _p(out),
                       ctypes.byref(inf), threads)
    assert rc == 0
    return out, bool(inf.value)


def msm_windows(n: int, scalar_bits: int) -> int:
    return lib().jfo_msm_windows(n, scalar_bits)


def fixed_base_mul(curve: str, scalars: np.ndarray, threads: int = 0) -> np.ndarray:
    L = CURVE_FQ_LIMBS[curve]
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    out = np.zeros((scalars.shape[0], 2 * L), dtype=np.uint64)
    assert lib().jfo_fixed_base_mul(CURVE_ID[curve], _p(scalars), scalars.shape[0], _p(out), threads) == 0
    return out


def gen_srs(curve: str, beta_limbs: np.ndarray, n: int, threads: int = 0) -> np.ndarray:
    L = CURVE_FQ_LIMBS[curve]
    out = np.zeros((n, 2 * L), dtype=np.uint64)
    b = np.ascontiguousarray(beta_limbs, dtype=np.uint64)
    assert lib().jfo_gen_srs(CURVE_ID[curve], _p(b), n, _p(out), threads) == 0
    return out


def poly_eval(field: str, coeffs: np.ndarray, x: np.ndarray) -> np.ndarray:
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    assert lib().jfo_poly_eval(FIELD_ID[field], _p(coeffs), coeffs.shape[0], _p(np.ascontiguousarray(x)), _p(out)) == 0
    return out


def random_field_elems(field: str, n: int, seed: int, montgomery: bool) -> np.ndarray:
    out = np.zeros((n, 4), dtype=np.uint64)
    assert lib().jfo_random_field_elems(FIELD_ID[field], n, seed, int(montgomery), _p(out)) == 0
    return out


# ---- int <-> limb helpers used all over the tests ---------------------------------------


def ints_to_limbs(vals, limbs: int) -> np.ndarray:
    out = np.zeros((len(vals), limbs), dtype=np.uint64)
    for i, v in enumerate(vals):
        for k in range(limbs):
            out[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(arr: np.ndarray):
    arr = np.asarray(arr, dtype=np.uint64)
    flat = arr.reshape(-1, arr.shape[-1])
    return [sum(int(flat[i, k]) << (64 * k) for k in range(flat.shape[1])) for i in range(flat.shape[0])]
