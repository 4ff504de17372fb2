"""The C ABI is called from rayon / tokio workers in the reference (`batch_commit`'s par_iter,
mod.rs:125-127; the coset-FFT par_iter, prover.rs:552-562): concurrent calls on ONE context must
serialise correctly, and separate contexts on the same GPU must not disturb each other."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_concurrent_calls_on_one_context_and_on_separate_contexts(ctx, co):
    import mpc_jellyfish_b200 as jf
    n = 3000
    ks = co.random_field_elems("bn254_fr", n, 1, False)
    pts = ctx.fixed_base_mul("bn254", ks)
    key = ctx.load_srs("bn254", pts)
    scalars = [co.random_field_elems("bn254_fr", n, 100 + t, False) for t in range(6)]
    want = [co.msm("bn254", pts, s) for s in scalars]
    polys = [co.random_field_elems("bn254_fr", 1 << 12, 200 + t, True) for t in range(6)]
    want_ntt = [co.ntt("bn254_fr", p, 12) for p in polys]
    errors = []

    def worker(t, c, k):
        try:
            for _ in range(5):
                xy, inf = c.msm(k, scalars[t])
                assert inf == want[t][1] and np.array_equal(xy, want[t][0])
                got = c.ntt("bn254_fr", polys[t].copy(), 12)
                assert np.array_equal(got, want_ntt[t])
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    # six threads hammering the session context
    th = [threading.Thread(target=worker, args=(t, ctx, key)) for t in range(6)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errors, errors
    # two more contexts with their own keys, all three used at once
    extra = [jf.Context(0) for _ in range(2)]
    keys = [c.load_srs("bn254", pts) for c in extra]
    th = [threading.Thread(target=worker, args=(t, c, k)) for t, (c, k) in enumerate(zip([ctx] + extra, [key] + keys))]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errors, errors
    for k in keys:
        k.free()
    for c in extra:
        c.close()
    key.free()
