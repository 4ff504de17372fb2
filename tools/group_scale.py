"""One process, several GPUs (`jf_group`): ONE MSM of 2^log_n pairs with the key range-sharded over 1, 2, 4, 8 GPUs, end to end
from page-locked host scalars, checked against the known-beta identity; and the 16-polynomial coset NTT of BASELINE
configs[2] dealt out by polynomial.  Prints one JSON line per measurement (kept under profiles/)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import coracle as co
import mpc_jellyfish_b200 as jf

BETA = 0x1D3C7A5B9E8F60412B7A6C5D4E3F20198A7B6C5D4E3F2A1B0C9D8E7F6A5B4C3
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
ngpu = torch.cuda.device_count()
sizes = [int(a) for a in sys.argv[1:]] or [20, 24]
for log_n in sizes:
    n = 1 << log_n
    s = co.random_field_elems("bn254_fr", n, 99 + log_n, False)
    pinned = torch.from_numpy(s.view(np.int64)).pin_memory().numpy().view(np.uint64)
    beta_m = co.field_op("bn254_fr", "to_mont", co.ints_to_limbs([BETA % R], 4))[0]
    ev = co.poly_eval("bn254_fr", co.field_op("bn254_fr", "to_mont", s), beta_m)
    want = co.fixed_base_mul("bn254", co.field_op("bn254_fr", "from_mont", ev[None, :]))[0]
    base = None
    for g in (1, 2, 4, 8):
        if g > ngpu:
            break
        grp = jf.Group(list(range(g)))
        key = grp.generate_srs_for_testing("bn254", BETA % R, n)
        xy, inf = grp.msm(key, pinned)
        assert not inf and np.array_equal(xy, want), "group msm mismatch at %d GPUs" % g
        for _ in range(2):
            grp.msm(key, pinned)
        K = 10 if log_n <= 22 else 5
        t0 = time.perf_counter()
        for _ in range(K):
            grp.msm(key, pinned)
        ms = (time.perf_counter() - t0) / K * 1e3
        base = base or ms
        print(json.dumps({"what": "jf_group_msm end to end (pinned host scalars -> affine point), one process", "pairs": n, "gpus": g,
                          "ms": round(ms, 3), "speedup": round(base / ms, 2), "efficiency": round(base / ms / g, 3),
                          "known_beta_identity": True}), flush=True)
        key.free()
        grp.close()
# batched coset NTT, BLS12-381 Fr, 16 x 2^22, dealt out by polynomial (host buffers: PCIe bound, which is what spreads)
log_n, batch = 22, 16
x = co.random_field_elems("bls12_381_fr", 2 << log_n, 5, True).reshape(2, 1 << log_n, 4)
buf = torch.empty((batch, 1 << log_n, 4), dtype=torch.int64).pin_memory().numpy().view(np.uint64)
off = np.array([0x0000000efffffff1, 0x17e363d300189c0f, 0xff9c57876f8457b0, 0x351332208fc5a8c4], dtype=np.uint64)
base = None
for g in (1, 2, 4, 8):
    if g > ngpu:
        break
    grp = jf.Group(list(range(g)))
    for b in range(batch):
        buf[b] = x[b % 2]
    grp.ntt("bls12_381_fr", buf, log_n, False, off)
    want0 = co.ntt("bls12_381_fr", x[0].copy(), log_n, False, off)
    assert np.array_equal(buf[0], want0) and np.array_equal(buf[batch - 2], want0), "group ntt mismatch at %d GPUs" % g
    t0 = time.perf_counter()
    for _ in range(3):
        grp.ntt("bls12_381_fr", buf, log_n, False, off)
    ms = (time.perf_counter() - t0) / 3 * 1e3
    base = base or ms
    print(json.dumps({"what": "jf_group_ntt end to end (pinned host vectors in place), 16 x 2^22 BLS12-381 Fr forward coset", "gpus": g,
                      "ms": round(ms, 2), "Melem_per_s": round(batch * (1 << log_n) / ms / 1e3), "speedup": round(base / ms, 2)}), flush=True)
    grp.close()
