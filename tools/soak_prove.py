"""Soak (exploration): the same 2^20-gate proof 40 times through every key variant; all serialisations must be
identical (a race between the prover's streams would show up as differing bytes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import mpc_jellyfish_b200 as jf
import bench_circuit as B
ctx = jf.Context(0)
arr = B.bench_circuit_arrays(ctx, 20)
key = ctx.generate_srs_for_testing("bn254", 0x1234567 + (7 << 200), arr["n"] + 3)
bl = np.random.default_rng(1).integers(0, 1 << 60, size=(17, 4), dtype=np.uint64)
seen = set()
for cache, skip, full in ((False, False, False), (False, True, False), (True, True, False), (False, False, True)):
    # full: round 3 on the reference's literal 8n coset instead of six sub-cosets -- same quotient, same bytes
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [],
                                     cache_coset_evals=cache, skip_zero_selectors=skip, full_quotient_coset=full)
    for i in range(40):
        seen.add(jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "standard" if i % 2 else "solidity").serialize_compressed())
    pk.free()
print("distinct proofs:", len(seen), "(expected 2: one per transcript)")
assert len(seen) == 2
