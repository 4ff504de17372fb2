"""Exploration harness (not part of the product): kernel breakdown of the MSM / NTT at bench sizes."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import mpc_jellyfish_b200 as jf
import coracle as co

def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16]
    ctx = jf.Context(0)
    print("imad.wide/s %.3e   mont_mul/s %.3e" % (ctx.microbench(0), ctx.microbench(1)))
    n = 1 << log_n
    s = co.random_field_elems("bn254_fr", n, 5, False)
    d_s = ctx.dev_alloc(32 * n); ctx.dev_upload(d_s, s)
    d_out = ctx.dev_alloc(128)
    for c in cs:
        t0 = time.time()
        key = ctx.generate_srs_for_testing("bn254", 0x1234567 + (7 << 200), n, window_bits=c)
        t_srs = time.time() - t0
        for _ in range(2):
            ctx.msm_device(key, d_s, n, d_out)
        ctx.sync()
        ctx.profile(True)
        t0 = time.time(); K = 5
        for _ in range(K):
            ctx.msm_device(key, d_s, n, d_out)
        prof = ctx.profile_collect()
        wall = (time.time() - t0) / K * 1e3
        ctx.profile(False)
        tot = sum(v[1] for v in prof.values()) / K
        print("c=%d srs_build %.2fs  msm wall %.3f ms  kernels %.3f ms" % (c, t_srs, wall, tot))
        for k, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
            print("   %-16s x%-3d %.4f ms/msm" % (k, cnt // K, ms / K))
        key.free()
    # NTT
    for field in ("bn254_fr", "bls12_381_fr"):
        for lg, batch in ((22, 1), (22, 16), (18, 16), (24, 4)):
            nn = 1 << lg
            x = co.random_field_elems(field, nn, 3, True)
            d = ctx.dev_alloc(32 * nn * batch)
            for b in range(batch):
                ctx.dev_upload(d + 32 * nn * b, x)
            off = co.field_op(field, "to_mont", np.array([[7, 0, 0, 0]], dtype=np.uint64))[0]
            for inverse in (False, True):
                for _ in range(2):
                    ctx.ntt_device(field, d, lg, inverse, off, batch=batch)
                ctx.sync(); ctx.profile(True); K = 5
                for _ in range(K):
                    ctx.ntt_device(field, d, lg, inverse, off, batch=batch)
                prof = ctx.profile_collect(); ctx.profile(False)
                tot = sum(v[1] for v in prof.values()) / K
                print("%s ntt 2^%d x%d inv=%d: %.3f ms  -> %.0f Melem/s  (%s)" % (
                    field, lg, batch, inverse, tot, batch * nn / tot / 1e3,
                    " ".join("%s:%.3f" % (k, v[1] / K) for k, v in prof.items())))
            ctx.dev_free(d)
main()
