import os, sys, time
ROOT = os.getcwd()
for p in (ROOT, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import mpc_jellyfish_b200 as jf
import bench_circuit as B
ctx = jf.Context(0)
for log_n in (14, 16):
    n = 1 << log_n
    arr = B.bench_circuit_arrays(ctx, log_n)
    key = ctx.generate_srs_for_testing("bn254", 0x1234567 + (7 << 200), n + 3)
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [])
    bl = np.random.default_rng(1).integers(0, 1 << 60, size=(17, 4), dtype=np.uint64)
    for _ in range(3):
        jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
    t0 = time.perf_counter()
    for _ in range(10):
        jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
    wall = (time.perf_counter() - t0) * 100
    l0 = ctx.launch_count
    ctx.profile(True)
    jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
    prof = ctx.profile_collect()
    ctx.profile(False)
    tot = sum(v[1] for v in prof.values())
    print("2^%d: wall %.2f ms, launches %d, sum of kernel times %.2f ms" % (log_n, wall, ctx.launch_count - l0, tot))
    print("   ", {k: (v[0], round(v[1], 3)) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:14]})
