"""Timing target (not part of the product): one TurboPlonk proof of the bench circuit (plonk/benches/bench.rs:29-46) per size,
2^LO .. 2^HI gates, default key and the key with every key-side option; each proof is checked by the restated verifier
(known-beta G1 form) before it is timed.  python tools/time_prove.py 18 23"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import mpc_jellyfish_b200 as jf
import bench_circuit as B
import coracle as co
import plonk_ref as P
import plonk_util as U
import pyref

lo, hi = int(sys.argv[1]), int(sys.argv[2])
ctx = jf.Context(0)
cv = pyref.BN254
beta = (0x1234567 + (7 << 200)) % cv.fr.p
for log_n in range(lo, hi + 1):
    n = 1 << log_n
    t0 = time.perf_counter()
    arr = B.bench_circuit_arrays(ctx, log_n)
    key = ctx.generate_srs_for_testing("bn254", beta, n + 3)
    t_setup = time.perf_counter() - t0
    bl = np.random.default_rng(1).integers(0, 1 << 60, size=(17, 4), dtype=np.uint64)
    row = {"log_n": log_n, "setup_s": round(t_setup, 2)}
    for name, kw in (("default", {}), ("all_options", dict(cache_coset_evals=True, skip_zero_selectors=True, lagrange_wire_commitments=True))):
        t0 = time.perf_counter()
        pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [], **kw)
        t_pre = time.perf_counter() - t0
        proof = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
        vk = U.vk_from_product(co, cv, pk, B.BN254_K)
        assert P.verify(cv, vk, [], U.proof_to_oracle(co, cv, proof), beta, "solidity"), "proof rejected"
        if name == "default":
            ref = proof.serialize_compressed()
        else:
            assert proof.serialize_compressed() == ref, "the key-side options changed the proof"
        jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
        steps = 3
        t0 = time.perf_counter()
        for _ in range(steps):
            jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
        row[name + "_ms"] = round((time.perf_counter() - t0) * 1e3 / steps, 2)
        row[name + "_preprocess_s"] = round(t_pre, 2)
        pk.free()
    key.free()
    print(row, flush=True)
