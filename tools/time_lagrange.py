"""Timing target (not part of the product): the five wire commitments of the bench circuit at 2^LOG gates as ONE MSM group
(a) over the monomial key with random-looking coefficients, (b) over the Lagrange-basis key with the witness values
(0 .. 2^LOG, a constant-one column, two zero columns) -- per-kernel breakdown of both."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mpc_jellyfish_b200 as jf

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log_n
ctx = jf.Context(0)
key = ctx.generate_srs_for_testing("bn254", 0x1234567 + (7 << 200), n + 3)
t0 = time.perf_counter()
lag = key.lagrange(log_n, mask_points=True)
print("lagrange key: %.3f s" % (time.perf_counter() - t0))
rng = np.random.default_rng(3)
rand = [np.ascontiguousarray(rng.integers(0, 1 << 62, size=(n + 2, 4), dtype=np.uint64) >> np.uint64(2)) for _ in range(5)]
small = np.zeros((n + 2, 4), dtype=np.uint64)
small[:n, 0] = np.arange(n, dtype=np.uint64)
ones = np.zeros((n + 2, 4), dtype=np.uint64)
ones[:n, 0] = 1
zeros = np.zeros((n + 2, 4), dtype=np.uint64)
succ = small.copy()
succ[:n, 0] += 1
for v in (small, ones, zeros, succ):
    v[n:] = rand[0][n:]           # the two masking scalars are full-size
vals = [small, ones, zeros.copy(), zeros.copy(), succ]
for name, k, vecs in (("monomial key, random scalars", key, rand), ("lagrange key, witness values", lag, vals)):
    for _ in range(2):
        ctx.msm_batch(k, vecs, None, montgomery=False)
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.msm_batch(k, vecs, None, montgomery=False)
    ms = (time.perf_counter() - t0) * 1e3 / 5
    ctx.profile(True)
    ctx.msm_batch(k, vecs, None, montgomery=False)
    prof = ctx.profile_collect()
    ctx.profile(False)
    print("%s: %.3f ms per group of 5 (host scalars)" % (name, ms))
    print("   ", {kk: round(v[1], 3) for kk, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:10]})
