"""Profiling target (not part of the product): 2 MSMs of 2^20 BN254 pairs + 1 batched coset NTT
(2^22 x 4, BN254 Fr), the same kernels and geometry bench.py times.  Used under
`ncu --set full -k regex:'msm_accumulate|ntt_pass'` (B200_PROFILING.md recipe)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import mpc_jellyfish_b200 as jf
import coracle as co

ctx = jf.Context(0)
n = 1 << 20
key = ctx.generate_srs_for_testing("bn254", 0x1234567 + (7 << 200), n)
s = co.random_field_elems("bn254_fr", n, 5, False)
d_s = ctx.dev_alloc(32 * n); ctx.dev_upload(d_s, s)
d_out = ctx.dev_alloc(128)
for _ in range(2):
    ctx.msm_device(key, d_s, n, d_out)
ctx.sync()
lg, batch = 22, 4
x = co.random_field_elems("bn254_fr", 1 << lg, 3, True)
d = ctx.dev_alloc(32 * (1 << lg) * batch)
for b in range(batch):
    ctx.dev_upload(d + 32 * (1 << lg) * b, x)
off = co.field_op("bn254_fr", "to_mont", np.array([[5, 0, 0, 0]], dtype=np.uint64))[0]
ctx.ntt_device("bn254_fr", d, lg, False, off, batch=batch)
ctx.sync()
print("prof_small ok, launches", ctx.launch_count)
