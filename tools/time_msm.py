"""Exploration: un-instrumented device time of one resident 2^log_n BN254 MSM (CUDA events around K back-to-back calls)."""
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
    s = rng.integers(0, 1 << 60, size=(8, n, 4), dtype=np.uint64)
    d = torch.from_numpy(s.view(np.int64)).cuda()
    d_out = ctx.dev_alloc(128)
    key = ctx.generate_srs_for_testing("bn254", 0x1234567 + (7 << 200), n)
    st = torch.cuda.Stream()
    ctx.set_stream(st.cuda_stream)
    for i in range(3):
        ctx.msm_device(key, d[i % 8].data_ptr(), n, d_out)
    ctx.sync()
    K = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(K):
        ctx.msm_device(key, d[i % 8].data_ptr(), n, d_out)
    e1.record(st)
    ctx.sync()
    print("2^%d c=%d: %.3f ms per MSM (device, %d rotating scalar sets, no per-kernel events)" % (log_n, key.window_bits, e0.elapsed_time(e1) / K, 8), flush=True)
    key.free(); ctx.dev_free(d_out)
