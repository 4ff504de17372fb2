"""One rank of the multi-GPU parity run (launched by torchrun, one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/multi/sharded_worker.py

Checks `jf_msm_sharded` / `jf_msm_sharded_device` (both transports) against (a) the known-beta identity
commit(p) == p(beta) * G over ALL ranks' coefficients, (b) the single-GPU `jf_msm` over the whole key on rank 0, for BN254 and
BLS12-381, balanced and ragged splits, an empty slice, Montgomery scalars, and many back-to-back calls (mailbox parities).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

import coracle as co
import mpc_jellyfish_b200 as jf
from mpc_jellyfish_b200.sharded import Comm, ShardedMsm, shard_range

BETA = 0x1D3C7A5B9E8F60412B7A6C5D4E3F20198A7B6C5D4E3F2A1B0C9D8E7F6A5B4C3
MOD = {"bn254": 21888242871839275222246405745257275088548364400416034343698204186575808495617,
       "bls12_381": 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001}
FR = {"bn254": "bn254_fr", "bls12_381": "bls12_381_fr"}


def known_beta_commitment(curve, scalars):
    fr = FR[curve]
    beta_m = co.field_op(fr, "to_mont", co.ints_to_limbs([BETA % MOD[curve]], 4))[0]
    ev = co.poly_eval(fr, co.field_op(fr, "to_mont", scalars), beta_m)
    return co.fixed_base_mul(curve, co.field_op(fr, "from_mont", ev[None, :]))[0]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = jf.Context(local)
    checks = 0
    for transport in ("p2p", "nccl"):
        comm = Comm.from_torch_distributed(ctx, transport=transport)
        assert comm.transport == transport, (comm.transport, transport)
        for curve, n in (("bn254", 1 << 14), ("bn254", (1 << 13) + 37), ("bls12_381", 3001), ("bn254", world - 1)):
            a, b = shard_range(n, world, rank)          # n = world - 1: the last rank's slice is empty
            key = ctx.generate_srs_for_testing(curve, BETA % MOD[curve], b - a, first_power=a)
            full = co.random_field_elems(FR[curve], max(n, 1), 1234 + n, False)[:n]
            want = known_beta_commitment(curve, full) if n else None
            sm = ShardedMsm(ctx, key, comm=comm)
            d = torch.from_numpy(full[a:b].view(np.int64).copy()).cuda()
            for it in range(5):                           # back-to-back: both mailbox parities, several times
                xy, inf = sm.msm(d.data_ptr(), b - a)
                assert not inf and np.array_equal(xy, want), "%s %s n=%d device form, call %d" % (transport, curve, n, it)
            xy, inf = sm.msm_host(full[a:b])
            assert not inf and np.array_equal(xy, want), "%s %s n=%d host form" % (transport, curve, n)
            mont = co.field_op(FR[curve], "to_mont", full[a:b]) if b > a else full[a:b]
            xy, inf = sm.msm_host(mont, montgomery=True)
            assert not inf and np.array_equal(xy, want), "%s %s n=%d Montgomery scalars" % (transport, curve, n)
            if rank == 0 and n:
                whole = ctx.generate_srs_for_testing(curve, BETA % MOD[curve], n)
                one_xy, one_inf = ctx.msm(whole, full)
                assert not one_inf and np.array_equal(one_xy, want), "single-GPU result differs"
                whole.free()
            # all-zero scalars: every partial is the identity
            z = np.zeros((b - a, 4), dtype=np.uint64)
            xy, inf = sm.msm_host(z)
            assert inf and not xy.any()
            # an out-of-range scalar on ONE rank surfaces there as InvalidParameters and nowhere hangs
            bad = full[a:b].copy()
            if rank == world - 1 and b > a:
                bad[0] = np.array([0xFFFFFFFFFFFFFFFF] * 4, dtype=np.uint64)
                try:
                    sm.msm_host(bad)
                    raise AssertionError("expected InvalidParameters")
                except jf.InvalidParameters:
                    pass
            else:
                sm.msm_host(bad)
            sm.close()
            key.free()
            checks += 1
        comm.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.write("rank %d ok (%d configurations x 2 transports)\n" % (rank, checks // 2))
    sys.stdout.flush()


if __name__ == "__main__":
    main()
