"""Exploration: 2^log_n-gate prove time against the commit key's window size c (the commitments of a round share one
bucket reduction, which shifts the single-MSM optimum of tools/sweep_c.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import mpc_jellyfish_b200 as jf
import bench_circuit as B

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cs = [int(x) for x in sys.argv[2:]] or [16, 17, 19, 20]
ctx = jf.Context(0)
arr = B.bench_circuit_arrays(ctx, log_n)
beta = 0x1D3C7A5B9E8F60412B7A6C5D4E3F20198A7B6C5D4E3F2A1B0C9D8E7F6A5B4C3 % B.BN254_FR_P
bl = np.random.default_rng(1).integers(0, 1 << 60, size=(17, 4), dtype=np.uint64)
ref = None
for c in cs:
    key = ctx.generate_srs_for_testing("bn254", beta, arr["n"] + 3, window_bits=c)
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [])
    for _ in range(2):
        pr = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
    ser = pr.serialize_compressed()
    ref = ref or ser
    assert ser == ref
    ctx.sync()
    K = 5
    t0 = time.time()
    for _ in range(K):
        jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
    print("2^%d gates, c=%d: prove %.2f ms" % (log_n, key.window_bits, (time.time() - t0) / K * 1e3), flush=True)
    pk.free(); key.free()
