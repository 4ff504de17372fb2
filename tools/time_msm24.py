"""Exploration: jf_msm at 2^24 + 3 pairs, end to end with pageable host scalars, and its kernel breakdown."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpc_jellyfish_b200 as jf
ctx = jf.Context(0)
n = (1 << 24) + 3
key = ctx.generate_srs_for_testing("bn254", 0x1234567 + (7 << 200), n)
coeffs = np.random.default_rng(2).integers(0, 1 << 60, size=(n, 4), dtype=np.uint64)
for it in range(4):
    t0 = time.time(); xy, inf = ctx.msm(key, coeffs, montgomery=True); t1 = time.time()
    print("msm 2^24+3 c=%d call %d: %.1f ms e2e (pageable host scalars)" % (key.window_bits, it, (t1 - t0) * 1e3), flush=True)
ctx.profile(True)
ctx.msm(key, coeffs, montgomery=True)
for k, (c, ms) in sorted(ctx.profile_collect().items(), key=lambda kv: -kv[1][1])[:6]:
    print("   %-24s x%-3d %.3f ms" % (k, c, ms))
