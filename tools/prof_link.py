"""Profiling target (not part of the product): ONE 2^LOG-gate TurboPlonk proof of the bench circuit followed by ONE proof-linking
call (`jf_plonk_link_proofs`, coset-division path, group of 1024) on two synthetic hints that agree on the link domain -- the
calls bench.py's `prove` and `link_proofs` legs time.  Used under
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_link.py 20
  ncu --set full --clock-control none -k regex:'fold_kernel|link_roots_check|vmul|inv_combine|quotient_kernel|lincomb' python tools/prof_link.py 20"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import mpc_jellyfish_b200 as jf
import bench_circuit as B

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = jf.Context(0)
n = 1 << log_n
arr = B.bench_circuit_arrays(ctx, log_n)
key = ctx.generate_srs_for_testing("bn254", 0x1234567 + (7 << 200), n + 3)
pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [])
bl = np.random.default_rng(1).integers(0, 1 << 60, size=(17, 4), dtype=np.uint64)
l0 = ctx.launch_count
proof, hint = jf.PlonkKzgSnark.prove_with_link_hint(pk, arr["witness"], bl, "solidity")
l1 = ctx.launch_count
align = log_n - 1
a2 = hint.linking_wire_poly.copy()
c = np.array([[5, 0, 0, 0]], dtype=np.uint64)
a2[7:8] = ctx.field_op("bn254_fr", "sub", a2[7:8], c)                       # a2 = a1 + 5 X^7 (X^(2^align) - 1)
a2[7 + (1 << align):8 + (1 << align)] = ctx.field_op("bn254_fr", "add", a2[7 + (1 << align):8 + (1 << align)], c)
c2, i2 = ctx.msm(key, a2, montgomery=True)
l2 = ctx.launch_count
lp = jf.PlonkKzgSnark.link_proofs(ctx, key, hint, jf.LinkingHint(a2, c2, bool(i2)), jf.GroupLayout(align, 100, 1024), "solidity")
assert lp.path == 0
print("prof_link ok: proof %d launches, link %d launches" % (l1 - l0, ctx.launch_count - l2))
