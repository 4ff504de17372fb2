"""Exploration harness (not part of the product): kernel breakdown of jf_plonk_prove at 2^log_n gates."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import mpc_jellyfish_b200 as jf
import bench_circuit as B

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = jf.Context(0)
t0 = time.time(); arr = B.bench_circuit_arrays(ctx, log_n); print("circuit arrays %.2fs" % (time.time() - t0))
beta = 0x1D3C7A5B9E8F60412B7A6C5D4E3F20198A7B6C5D4E3F2A1B0C9D8E7F6A5B4C3 % B.BN254_FR_P
t0 = time.time(); key = ctx.generate_srs_for_testing("bn254", beta, arr["n"] + 3); print("srs %.2fs c=%d" % (time.time() - t0, key.window_bits))
rng = np.random.default_rng(1)
bl = rng.integers(0, 1 << 60, size=(17, 4), dtype=np.uint64)  # < 2^252 < r: valid Montgomery residues
for cache in (False, True):
    t0 = time.time()
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [],
                                     cache_coset_evals=cache)
    print("preprocess(cache=%s) %.2fs" % (cache, time.time() - t0))
    for _ in range(2):
        pr = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
    ctx.sync()
    K = 3
    t0 = time.time()
    for _ in range(K):
        pr = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
    wall = (time.time() - t0) / K * 1e3
    ctx.profile(True)
    pr = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
    prof = ctx.profile_collect(); ctx.profile(False)
    tot = sum(v[1] for v in prof.values())
    print("prove 2^%d cache=%s: wall %.2f ms; kernels %.2f ms" % (log_n, cache, wall, tot))
    for k, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:14]:
        print("   %-16s x%-4d %.3f ms" % (k, cnt, ms))
    pk.free()
