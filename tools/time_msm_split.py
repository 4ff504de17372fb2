"""Exploration: one 2^k MSM from page-locked host scalars as ONE jf_msm call against a jf_msm_batch of P equal parts with base
offsets (upload of part i+1 beside the kernels of part i, one shared bucket reduction); the parts' points are added on the host."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import mpc_jellyfish_b200 as jf
ctx = jf.Context(0)
rng = np.random.default_rng(3)
for log_n in [int(a) for a in sys.argv[1:]] or [20, 22]:
    n = 1 << log_n
    s = rng.integers(0, 1 << 60, size=(4, n, 4), dtype=np.uint64)
    pinned = torch.from_numpy(s.view(np.int64)).pin_memory()
    host = [pinned[k].numpy().view(np.uint64) for k in range(4)]
    key = ctx.generate_srs_for_testing("bn254", 0x1234567 + (7 << 200), n)
    for parts in (1, 2, 4):
        h = n // parts
        def run(v):
            if parts == 1:
                return ctx.msm(key, v)
            return ctx.msm_batch(key, [v[p * h:(p + 1) * h] for p in range(parts)], base_offsets=[p * h for p in range(parts)])
        for i in range(3):
            run(host[i % 4])
        K = 20
        t0 = time.perf_counter()
        for i in range(K):
            run(host[i % 4])
        dt = (time.perf_counter() - t0) / K * 1e3
        print("2^%d c=%d parts=%d: %.3f ms end to end" % (log_n, key.window_bits, parts, dt), flush=True)
    key.free()
