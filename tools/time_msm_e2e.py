"""Exploration: end-to-end time of `jf_msm` (page-locked host scalars in, affine point out) per size; run once with
JF_MSM_ZEROCOPY=1 (default: the sort kernel reads the scalars straight from host memory) and once with =0 (copy first)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import mpc_jellyfish_b200 as jf
ctx = jf.Context(0)
rng = np.random.default_rng(3)
for log_n in [int(a) for a in sys.argv[1:]] or [16, 18, 20, 22]:
    n = 1 << log_n
    s = rng.integers(0, 1 << 60, size=(4, n, 4), dtype=np.uint64)
    pinned = torch.from_numpy(s.view(np.int64)).pin_memory()
    host = [pinned[k].numpy().view(np.uint64) for k in range(4)]
    key = ctx.generate_srs_for_testing("bn254", 0x1234567 + (7 << 200), n)
    ref = [ctx.msm(key, s[k]) for k in range(4)]          # pageable source: always the copy path
    for i in range(3):
        got = ctx.msm(key, host[i % 4])
        assert np.array_equal(got[0], ref[i % 4][0])
    K = 20
    t0 = time.perf_counter()
    for i in range(K):
        ctx.msm(key, host[i % 4])
    dt = (time.perf_counter() - t0) / K * 1e3
    print("2^%d c=%d: %.3f ms per jf_msm end to end (pinned host scalars, JF_MSM_ZEROCOPY=%s)" % (
        log_n, key.window_bits, dt, os.environ.get("JF_MSM_ZEROCOPY", "1")), flush=True)
    key.free()
