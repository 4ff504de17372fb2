"""Multi-GPU MSM / batched NTT, host side of csrc/comm.cu.

`commit` is a sum over (coefficient, key point) pairs (primitives/src/pcs/univariate_kzg/mod.rs:106-111), so it is sharded by
point range: GPU g keeps key[start_g, end_g) resident and receives the matching scalar slice, runs the full Pippenger
pipeline on it and contributes one XYZZ partial sum (128 bytes for BN254); ONE exchange of those partials follows.

* `Comm` / `ShardedMsm`: one process per GPU (`torchrun`).  The exchange happens inside the library, on the context's own
  stream, right behind the MSM kernels (`jf_msm_sharded`): peer-memory mailboxes over NVLink or `ncclAllGather`.
* `Group`: one process driving several GPUs (`jf_group_*`), the form a Rust prover process uses.
* Batched NTTs shard by polynomial and need no exchange (`poly_owner`, `Group.ntt`).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from .errors import InvalidParameters, raise_for_status


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of [0, n): the first n % world ranks get one extra point."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def poly_owner(poly_index: int, world: int) -> int:
    """Batched NTTs: polynomial p lives on GPU p mod world."""
    return poly_index % world


def combine_partials(curve: str, xyzz_parts: np.ndarray):
    """Sum XYZZ partial results (parts, 4L uint64) and normalise -> (xy, is_infinity).
    Pure host code inside libjf_b200.so (a few dozen field operations); needs no GPU."""
    L = _ffi.CURVE_FQ_LIMBS[curve]
    p = np.ascontiguousarray(xyzz_parts, dtype=np.uint64).reshape(-1, 4 * L)
    out = np.zeros(2 * L, dtype=np.uint64)
    inf = ctypes.c_int(0)
    rc = _ffi.lib().jf_msm_combine(None, _ffi.CURVES[curve], p.ctypes.data_as(_ffi.c_u64p), p.shape[0],
                                   out.ctypes.data_as(_ffi.c_u64p), ctypes.byref(inf))
    raise_for_status(rc, "jf_msm_combine failed")
    return out, bool(inf.value)


def all_gather_partials(local_xyzz, group=None):
    """all-gather one XYZZ partial per rank through `torch.distributed` (kept for callers that own the exchange
    themselves and for the gloo tests; `Comm` does this step inside the library).  `local_xyzz` is a torch tensor
    (int64, 4L words) on the device the process group's backend wants (cuda for nccl, cpu for gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    flat = torch.empty((world * local_xyzz.numel(),), dtype=local_xyzz.dtype, device=local_xyzz.device)
    dist.all_gather_into_tensor(flat, local_xyzz.reshape(-1).contiguous(), group=group)
    return flat.reshape(world, local_xyzz.numel())


TRANSPORTS = {"auto": 0, "nccl": 1, "p2p": 2}


def _preload_nccl():
    """The library binds NCCL at run time by SONAME (`libnccl.so.2`).  A process holds ONE library per SONAME, so if the
    system copy were loaded first and torch (which needs the newer copy it bundles) imported later, torch would fail to
    resolve its symbols.  In a Python process that has the bundled copy, load that one first; everywhere else (a Rust
    prover process) the library finds the system's."""
    import importlib.util
    import os
    import sys
    if "torch" in sys.modules:
        return
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            path = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(path):
                ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
                return
    except Exception:
        pass


class Comm:
    """This rank's end of a one-process-per-GPU group (`jf_comm`)."""

    def __init__(self, ctx, rank: int, nranks: int, unique_id: bytes, transport: str = "auto"):
        if len(unique_id) != _ffi.JF_COMM_ID_BYTES:
            raise InvalidParameters("unique_id must be %d bytes" % _ffi.JF_COMM_ID_BYTES)
        self.ctx, self.rank, self.nranks = ctx, rank, nranks
        _preload_nccl()
        h = ctypes.c_void_p()
        ctx._check(ctx._lib.jf_comm_init(ctx._h, rank, nranks, unique_id, TRANSPORTS[transport], ctypes.byref(h)))
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        _preload_nccl()
        buf = ctypes.create_string_buffer(_ffi.JF_COMM_ID_BYTES)
        raise_for_status(_ffi.lib().jf_comm_unique_id(buf), "jf_comm_unique_id failed (is libnccl.so.2 loadable?)")
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, ctx, group=None, transport: str = "auto") -> "Comm":
        """Bootstrap over an existing `torch.distributed` group (any backend): rank 0's id is broadcast."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(ctx, rank, world, box[0], transport)

    @property
    def transport(self) -> str:
        return {1: "nccl", 2: "p2p"}.get(int(self.ctx._lib.jf_comm_transport(self._h)), "?")

    def msm(self, key_slice, scalars: np.ndarray, base_offset: int = 0, montgomery: bool = False):
        """Collective.  This rank's scalar slice (host) -> the affine result of the whole MSM, on every rank."""
        L = _ffi.CURVE_FQ_LIMBS[key_slice.curve]
        s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros(2 * L, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self.ctx._check(self.ctx._lib.jf_msm_sharded(self.ctx._h, self._h, key_slice._h, base_offset, s.ctypes.data_as(_ffi.c_u64p),
                                                     s.shape[0], int(montgomery), out.ctypes.data_as(_ffi.c_u64p), ctypes.byref(inf)))
        return out, bool(inf.value)

    def msm_device(self, key_slice, d_scalars: int, n_local: int, d_out_parts: int, base_offset: int = 0,
                   montgomery: bool = False):
        """Collective, asynchronous: scalars in HBM; leaves the `nranks` XYZZ partials in d_out_parts on every rank."""
        self.ctx._check(self.ctx._lib.jf_msm_sharded_device(self.ctx._h, self._h, key_slice._h, base_offset, ctypes.c_void_p(d_scalars),
                                                            n_local, int(montgomery), ctypes.c_void_p(d_out_parts)))

    def close(self):
        if getattr(self, "_h", None) and self.ctx._h:
            self.ctx._lib.jf_comm_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedMsm:
    """Range-sharded `msm_bigint`: build with this rank's key slice, call `msm` with this rank's scalar slice (device
    pointer); every rank returns the same affine result.  The MSM kernels, the exchange and the read-back are all ordered on
    the context's stream inside the library, whatever stream torch is on."""

    def __init__(self, ctx, key_slice, group=None, transport: str = "auto", comm: Optional[Comm] = None):
        self.ctx, self.key = ctx, key_slice
        self.comm = comm if comm is not None else Comm.from_torch_distributed(ctx, group, transport)
        L = _ffi.CURVE_FQ_LIMBS[key_slice.curve]
        self._d_parts = ctx.dev_alloc(32 * L * self.comm.nranks)
        self._h_parts = np.zeros((self.comm.nranks, 4 * L), dtype=np.uint64)

    def msm(self, d_scalars: int, n_local: int, montgomery: bool = False):
        self.comm.msm_device(self.key, d_scalars, n_local, self._d_parts, montgomery=montgomery)
        self.ctx.dev_download(self._h_parts, self._d_parts)  # stream-ordered behind the exchange; synchronises
        self.ctx.sync()  # surfaces a scalar-range / exchange error of this rank
        return combine_partials(self.key.curve, self._h_parts)

    def msm_host(self, scalars: np.ndarray, montgomery: bool = False):
        return self.comm.msm(self.key, scalars, montgomery=montgomery)

    def close(self):
        if self._d_parts:
            self.ctx.dev_free(self._d_parts)
            self._d_parts = 0


class GroupKey:
    """A commit key split by point range over the GPUs of a `Group` (`jf_group_srs`)."""

    def __init__(self, group: "Group", curve: str, handle, n: int):
        self.group, self.curve, self._h, self.n = group, curve, handle, n

    def __len__(self):
        return self.n

    def free(self):
        if self._h is not None and self.group._h:
            self.group._lib.jf_group_srs_free(self.group._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Group:
    """Several GPUs driven by one process (`jf_group`): the drop-in behind `UnivariateKzgPCS::commit` and the batched
    coset transforms of `prover.rs:552-567` for a single prover process."""

    def __init__(self, devices: Sequence[int]):
        self._lib = _ffi.lib()
        arr = (ctypes.c_int * len(devices))(*devices)
        h = ctypes.c_void_p()
        rc = self._lib.jf_group_create(arr, len(devices), ctypes.byref(h))
        if rc != _ffi.JF_OK:
            raise RuntimeError("jf_group_create(%r) failed with status %d: no usable CUDA device "
                               "(this library has no CPU fallback)" % (list(devices), rc))
        self._h = h
        self.devices = list(devices)

    def _check(self, rc: int):
        if rc != _ffi.JF_OK:
            raise_for_status(rc, (self._lib.jf_group_last_error(self._h) or b"").decode())

    def __len__(self):
        return int(self._lib.jf_group_size(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.jf_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_srs(self, curve: str, points: np.ndarray, window_bits: int = 0, precompute: bool = True) -> GroupKey:
        L = _ffi.CURVE_FQ_LIMBS[curve]
        pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 2 * L)
        h = ctypes.c_void_p()
        self._check(self._lib.jf_group_srs_load(self._h, _ffi.CURVES[curve], pts.ctypes.data_as(ctypes.c_void_p), pts.shape[0], 16 * L,
                                                -1, window_bits, int(precompute), ctypes.byref(h)))
        return GroupKey(self, curve, h, pts.shape[0])

    def generate_srs_for_testing(self, curve: str, beta: int, n: int, window_bits: int = 0, precompute: bool = True) -> GroupKey:
        b = np.array([(beta >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)], dtype=np.uint64)
        h = ctypes.c_void_p()
        self._check(self._lib.jf_group_srs_generate_for_testing(self._h, _ffi.CURVES[curve], b.ctypes.data_as(_ffi.c_u64p), n, window_bits,
                                                                int(precompute), ctypes.byref(h)))
        return GroupKey(self, curve, h, n)

    def msm(self, key: GroupKey, scalars: np.ndarray, base_offset: int = 0, montgomery: bool = False):
        L = _ffi.CURVE_FQ_LIMBS[key.curve]
        s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros(2 * L, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self._check(self._lib.jf_group_msm(self._h, key._h, base_offset, s.ctypes.data_as(_ffi.c_u64p), s.shape[0], int(montgomery),
                                           out.ctypes.data_as(_ffi.c_u64p), ctypes.byref(inf)))
        return out, bool(inf.value)

    def ntt(self, field: str, data: np.ndarray, log_n: int, inverse: bool = False, coset_offset: Optional[np.ndarray] = None,
            in_len: Optional[int] = None) -> np.ndarray:
        """In place on `data` ((batch, n, 4) uint64): vector b is transformed on GPU b mod len(self)."""
        if data.dtype != np.uint64 or not data.flags["C_CONTIGUOUS"] or not data.flags["WRITEABLE"] or data.ndim != 3 \
                or data.shape[1:] != (1 << log_n, 4):
            raise InvalidParameters("Group.ntt wants a writable C-contiguous (batch, n, 4) uint64 array")
        off = None
        if coset_offset is not None:
            off = np.ascontiguousarray(coset_offset, dtype=np.uint64).reshape(4)
        n = 1 << log_n
        self._check(self._lib.jf_group_ntt(self._h, _ffi.FIELDS[field], data.ctypes.data_as(_ffi.c_u64p), n if in_len is None else in_len,
                                           log_n, int(inverse), off.ctypes.data_as(_ffi.c_u64p) if off is not None else None,
                                           data.shape[0], n))
        return data
