"""Multi-GPU MSM: one process per GPU, the commit key sharded by point range.

Rank r keeps key[start_r, end_r) resident and receives the matching slice of every scalar
vector; it runs the full Pippenger pipeline on its slice and leaves one XYZZ partial sum in
HBM.  The only exchange step is one all-gather of those 128-byte (BN254) partials
(`torch.distributed`, NCCL over NVLink on GPUs, gloo in the CPU tests), after which every rank
adds the `world` partials and normalises (`jf_msm_combine`).  Batched NTTs shard by polynomial
and need no collective at all (`poly_owner`).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _ffi
from .errors import raise_for_status


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
    """all-gather one XYZZ partial per rank.  `local_xyzz` is a torch tensor (int64, 4L words) on
    the device the process group's backend wants (cuda for nccl, cpu for gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    flat = torch.empty((world * local_xyzz.numel(),), dtype=local_xyzz.dtype, device=local_xyzz.device)
    dist.all_gather_into_tensor(flat, local_xyzz.reshape(-1).contiguous(), group=group)
    return flat.reshape(world, local_xyzz.numel())


class ShardedMsm:
    """Range-sharded `msm_bigint`: build with this rank's key slice, call `msm` with this rank's
    scalar slice (device pointer); every rank returns the same affine result."""

    def __init__(self, ctx, key_slice, group=None):
        self.ctx, self.key, self.group = ctx, key_slice, group

    def msm(self, d_scalars: int, n_local: int, montgomery: bool = False):
        import torch
        L = _ffi.CURVE_FQ_LIMBS[self.key.curve]
        part = torch.zeros((4 * L,), dtype=torch.int64, device="cuda")
        self.ctx.msm_device(self.key, d_scalars, n_local, part.data_ptr(), montgomery=montgomery)
        parts = all_gather_partials(part, self.group)
        host = parts.cpu().numpy().view(np.uint64)
        self.ctx.sync()  # surfaces a scalar-range error of this rank
        return combine_partials(self.key.curve, host)
