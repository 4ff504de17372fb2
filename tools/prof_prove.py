"""Profiling target (not part of the product): preprocess + ONE 2^LOG-gate proof of the bench circuit
(plonk/benches/bench.rs:29-46), the same call bench.py's prove leg times.  Used under
  ncu --set full --clock-control none --import-source on -k regex:'ntt_pass|quotient|subcoset_solve' \
      --launch-skip 6 --launch-count 48 python tools/prof_prove.py 20
(the 6 skipped launches are the selector / sigma iNTT passes of preprocess)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import mpc_jellyfish_b200 as jf
import bench_circuit as B

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
flags = dict(full_quotient_coset="full" in sys.argv[2:])
ctx = jf.Context(0)
n = 1 << log_n
arr = B.bench_circuit_arrays(ctx, log_n)
key = ctx.generate_srs_for_testing("bn254", 0x1234567 + (7 << 200), n + 3)
pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [], **flags)
bl = np.random.default_rng(1).integers(0, 1 << 60, size=(17, 4), dtype=np.uint64)
proof = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
print("prof_prove ok: %d proof bytes, launches %d" % (len(proof.serialize_compressed()), ctx.launch_count))
