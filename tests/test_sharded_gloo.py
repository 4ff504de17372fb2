"""CPU suite, part 4: the N > 1 path on world_size 2 with the gloo backend.  Each rank owns a
point range, produces its partial sum (here by the oracle, standing in for the GPU kernel, which
needs no collective), all-gathers the XYZZ partials through the product's own `all_gather_partials`
and combines them with the product's `jf_msm_combine`; every rank must obtain the oracle's
full-range result."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
import numpy as np, torch, torch.distributed as dist
import coracle as co
import mpc_jellyfish_b200 as jf
from mpc_jellyfish_b200.sharded import all_gather_partials, combine_partials, shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 301
ks = co.random_field_elems("bn254_fr", n, 5, False)
pts = co.fixed_base_mul("bn254", ks)
s = co.random_field_elems("bn254_fr", n, 6, False)
a, b = shard_range(n, world, rank)
xy, inf = co.msm("bn254", pts[a:b], s[a:b])
one = co.field_op("bn254_fq", "to_mont", np.array([[1, 0, 0, 0]], dtype=np.uint64))[0]
part = np.zeros(16, dtype=np.uint64)                       # XYZZ = (x, y, 1, 1) or identity
if not inf:
    part[:8] = xy; part[8:12] = one; part[12:] = one
parts = all_gather_partials(torch.from_numpy(part.view(np.int64)))
got_xy, got_inf = combine_partials("bn254", parts.numpy().view(np.uint64))
want_xy, want_inf = co.msm("bn254", pts, s)
assert got_inf == want_inf and np.array_equal(got_xy, want_xy), "rank %d mismatch" % rank
# identity partials and cancelling partials
z = all_gather_partials(torch.zeros(16, dtype=torch.int64))
assert combine_partials("bn254", z.numpy().view(np.uint64))[1]
neg = part.copy()
if rank == 1:
    neg[4:8] = co.field_op("bn254_fq", "neg", part[None, 4:8])[0]
same = np.zeros(16, dtype=np.uint64); same[:8] = want_xy; same[8:12] = one; same[12:] = one
mine = same.copy()
if rank == 1:
    mine[4:8] = co.field_op("bn254_fq", "neg", same[None, 4:8])[0]
c = all_gather_partials(torch.from_numpy(mine.view(np.int64)))
assert combine_partials("bn254", c.numpy().view(np.uint64))[1], "P + (-P) must be the identity"
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank %d ok\n" % rank); sys.stdout.flush()   # one write: the ranks share a pipe
'''


def test_range_sharded_msm_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    import socket
    with socket.socket() as sk:  # a free port: a fixed one can still be in TIME_WAIT from an earlier run
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
