import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import mpc_jellyfish_b200 as jf, bench_circuit as B, plonk_util as U, plonk_ref as P, pyref, coracle as co
ctx = jf.Context(0)
cv, fr = pyref.BN254, pyref.BN254_FR
beta = 0x1D3C7A5B9E8F60412B7A6C5D4E3F20198A7B6C5D4E3F2A1B0C9D8E7F6A5B4C3 % fr.p
for log_n in (22,):
    t0 = time.time(); arr = B.bench_circuit_arrays(ctx, log_n); t1 = time.time()
    key = ctx.generate_srs_for_testing("bn254", beta, arr["n"] + 3); t2 = time.time()
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [], skip_zero_selectors=True); t3 = time.time()
    bl = np.random.default_rng(1).integers(0, 1 << 60, size=(17, 4), dtype=np.uint64)
    pr = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity"); t4 = time.time()
    pr = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity"); t5 = time.time()
    ok = P.verify(cv, U.vk_from_product(co, cv, pk, B.BN254_K), [], U.proof_to_oracle(co, cv, pr), beta, "solidity")
    print("2^%d gates: arrays %.1fs srs %.1fs (c=%d) preprocess %.1fs prove %.3fs / %.3fs verified=%s" % (log_n, t1-t0, t2-t1, key.window_bits, t3-t2, t4-t3, t5-t4, ok), flush=True)
    pk.free(); key.free()
# MSM 2^24 known-beta identity
n = (1 << 24) + 3
key = ctx.generate_srs_for_testing("bn254", beta, n)
coeffs = np.random.default_rng(2).integers(0, 1 << 60, size=(n, 4), dtype=np.uint64)
ctx.msm(key, coeffs, montgomery=True)  # first call: grows the workspaces (cudaMalloc of several GB)
t0 = time.time(); xy, inf = ctx.msm(key, coeffs, montgomery=True); t1 = time.time()
ev = co.poly_eval("bn254_fr", coeffs, co.ints_to_limbs([fr.to_mont(beta)], 4)[0])
want = co.fixed_base_mul("bn254", co.field_op("bn254_fr", "from_mont", ev[None, :]))[0]
print("msm 2^24+3 (c=%d): %.1f ms e2e, known-beta identity %s" % (key.window_bits, (t1-t0)*1e3, np.array_equal(xy, want)))
