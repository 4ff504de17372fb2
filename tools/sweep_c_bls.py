"""Exploration: BLS12-381 MSM time vs window size c (precomputed tables)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mpc_jellyfish_b200 as jf
ctx = jf.Context(0)
rng = np.random.default_rng(3)
for log_n, cs in ((18, (15, 16, 17, 18)), (20, (16, 17, 18, 19, 20))):
    n = 1 << log_n
    s = rng.integers(0, 1 << 60, size=(n, 4), dtype=np.uint64)
    d_s = ctx.dev_alloc(32 * n); ctx.dev_upload(d_s, s)
    d_out = ctx.dev_alloc(192)
    for c in cs:
        key = ctx.generate_srs_for_testing("bls12_381", 0x1234567 + (7 << 200), n, window_bits=c)
        for _ in range(2):
            ctx.msm_device(key, d_s, n, d_out)
        ctx.sync()
        K = 5
        t0 = time.perf_counter()
        for _ in range(K):
            ctx.msm_device(key, d_s, n, d_out)
        ctx.sync()
        wall = (time.perf_counter() - t0) / K * 1e3
        ctx.profile(True)
        ctx.msm_device(key, d_s, n, d_out)
        prof = ctx.profile_collect(); ctx.profile(False)
        g = lambda k: prof.get(k, (0, 0.0))[1]
        print("bls12_381 2^%d c=%d: %.3f ms  (acc %.3f reduce %.3f count %.3f scatter %.3f bsum %.3f)" % (
            log_n, c, wall, g("msm_accumulate"), g("reduce_level"), g("msm_count"), g("msm_scatter"),
            g("bucket_sum") + g("bucket_sum_heavy")), flush=True)
        key.free()
    ctx.dev_free(d_s); ctx.dev_free(d_out)
