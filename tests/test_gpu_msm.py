"""K2 parity: GPU Pippenger MSM vs the CPU oracle's restatement of `msm_bigint(..).into_affine()`:
identical affine (x, y, infinity), Montgomery limbs compared bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CURVES = ["bn254", "bls12_381"]


def _points(ctx, co, curve, n, seed):
    """n points with known discrete logs k_i (P_i = k_i G), from the library's fixed-base path,
    spot-checked against the oracle."""
    fr = "bn254_fr" if curve == "bn254" else "bls12_381_fr"
    ks = co.random_field_elems(fr, n, seed, False)
    pts = ctx.fixed_base_mul(curve, ks)
    idx = sorted(set([0, n - 1, n // 2] + list(range(min(n, 4)))))
    assert np.array_equal(pts[idx], co.fixed_base_mul(curve, ks[idx]))
    return ks, pts


def _check(ctx, co, curve, pts, scalars, window_bits=0, precompute=True, base_offset=0):
    key = ctx.load_srs(curve, pts, window_bits=window_bits, precompute=precompute)
    try:
        got_xy, got_inf = ctx.msm(key, scalars, base_offset=base_offset)
    finally:
        key.free()
    want_xy, want_inf = co.msm(curve, pts[base_offset:], scalars)
    assert got_inf == want_inf
    assert np.array_equal(got_xy, want_xy)


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("n", [1, 2, 3, 31, 33, 257, 1000])
@pytest.mark.parametrize("precompute", [True, False])
def test_msm_small_random(ctx, co, curve, n, precompute):
    fr = "bn254_fr" if curve == "bn254" else "bls12_381_fr"
    _, pts = _points(ctx, co, curve, n, 1000 + n)
    s = co.random_field_elems(fr, n, 2000 + n, False)
    _check(ctx, co, curve, pts, s, precompute=precompute)


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("c", [2, 5, 8, 13, 16])
@pytest.mark.parametrize("precompute", [True, False])
def test_msm_window_sizes(ctx, co, curve, c, precompute):
    fr = "bn254_fr" if curve == "bn254" else "bls12_381_fr"
    n = 300
    _, pts = _points(ctx, co, curve, n, 5)
    s = co.random_field_elems(fr, n, 6, False)
    _check(ctx, co, curve, pts, s, window_bits=c, precompute=precompute)


@pytest.mark.parametrize("curve", CURVES)
def test_msm_edge_scalars(ctx, co, py, curve):
    """SURVEY §8d edge suite: zeros, ones, r-1, small, mostly zero, all equal (one bucket piles up)."""
    cv = py.CURVES[curve]
    r = cv.fr.p
    n = 600
    _, pts = _points(ctx, co, curve, n, 8)
    rnd = py.random_field_elems(cv.fr, n, seed=9)
    suites = {
        "zero": [0] * n,
        "one": [1] * n,
        "r-1": [r - 1] * n,
        "small": [v & 0xFFFFFFFFFFFFFFFF for v in rnd],
        "sparse": [v if i % 10 == 0 else 0 for i, v in enumerate(rnd)],
        "equal": [rnd[0]] * n,
        "top": [r - 1 - i for i in range(n)],
        "halfwindow": [(1 << 15) + (1 << 31) + (1 << 47)] * n,
    }
    for name, vals in suites.items():
        s = co.ints_to_limbs(vals, 4)
        for pre in (True, False):
            _check(ctx, co, curve, pts, s, window_bits=16 if name != "equal" else 6, precompute=pre)


@pytest.mark.parametrize("curve", CURVES)
def test_msm_edge_points(ctx, co, py, curve):
    """Identity points in the key, duplicated points (P + P inside a bucket), P and -P pairs."""
    cv = py.CURVES[curve]
    L = cv.fq.limbs64
    fr = "bn254_fr" if curve == "bn254" else "bls12_381_fr"
    n = 64
    _, pts = _points(ctx, co, curve, n, 12)
    pts = pts.copy()
    pts[3] = 0                                  # identity
    pts[10] = pts[11]                           # duplicate
    neg = co.field_op(cv.fq.name, "neg", pts[20:21, L:])
    pts[21, :L] = pts[20, :L]
    pts[21, L:] = neg[0]                        # P and -P
    for vals in ([7] * n, [cv.fr.p - 1] * n, py.random_field_elems(cv.fr, n, seed=1)):
        s = co.ints_to_limbs(vals, 4)
        for c in (3, 8):
            _check(ctx, co, curve, pts, s, window_bits=c, precompute=True)
            _check(ctx, co, curve, pts, s, window_bits=c, precompute=False)
    # everything cancels -> identity result
    s = co.ints_to_limbs([0] * 20 + [5, 5] + [0] * (n - 22), 4)
    key = ctx.load_srs(curve, pts, window_bits=4)
    xy, inf = ctx.msm(key, s)
    key.free()
    assert inf and not xy.any()


def test_msm_truncates_to_min_len_and_offset(ctx, co):
    """`msm_bigint` uses min(len(bases), len(scalars)) pairs; `commit` offsets the key (mod.rs:110)."""
    _, pts = _points(ctx, co, "bn254", 100, 3)
    s = co.random_field_elems("bn254_fr", 150, 4, False)
    key = ctx.load_srs("bn254", pts)
    for off, m in ((0, 150), (0, 40), (7, 150), (99, 5), (100, 3)):
        xy, inf = ctx.msm(key, s[:m], base_offset=off)
        wxy, winf = co.msm("bn254", pts[off:], s[:m])
        assert inf == winf and np.array_equal(xy, wxy)
    key.free()


def test_msm_montgomery_scalars_and_range_check(ctx, co, py):
    import mpc_jellyfish_b200 as jf
    _, pts = _points(ctx, co, "bn254", 50, 3)
    s = co.random_field_elems("bn254_fr", 50, 4, False)
    sm = co.field_op("bn254_fr", "to_mont", s)
    key = ctx.load_srs("bn254", pts)
    a = ctx.msm(key, s)
    b = ctx.msm(key, sm, montgomery=True)
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]
    bad = s.copy()
    bad[7] = co.ints_to_limbs([py.BN254_FR.p], 4)[0]  # == r: not canonical
    with pytest.raises(jf.InvalidParameters):
        ctx.msm(key, bad)
    # the context stays usable afterwards
    c = ctx.msm(key, s)
    assert np.array_equal(a[0], c[0])
    key.free()


def test_msm_batch_matches_single(ctx, co):
    _, pts = _points(ctx, co, "bn254", 500, 3)
    key = ctx.load_srs("bn254", pts)
    vecs = [co.random_field_elems("bn254_fr", m, 40 + m, False) for m in (500, 1, 77, 0, 499)]
    offs = [0, 3, 100, 0, 1]
    out, infs = ctx.msm_batch(key, vecs, offs)
    for i, (v, o) in enumerate(zip(vecs, offs)):
        wxy, winf = co.msm("bn254", pts[o:], v) if len(v) else (np.zeros(8, np.uint64), True)
        assert infs[i] == winf and np.array_equal(out[i], wxy)
    key.free()


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("precompute", [True, False])
def test_msm_batch_spanning_several_groups(ctx, co, curve, precompute):
    """jf_msm_batch runs in groups of 8 that share one bucket reduction (one bucket set per member and window set):
    19 ragged vectors incl. empty ones at the group boundaries, all-equal scalars and a single-element vector."""
    fr = "bn254_fr" if curve == "bn254" else "bls12_381_fr"
    _, pts = _points(ctx, co, curve, 300, 9)
    key = ctx.load_srs(curve, pts, window_bits=6, precompute=precompute)
    lens = [0, 300, 1, 17, 299, 0, 64, 300, 0, 5, 300, 33, 0, 0, 128, 2, 300, 7, 0]
    vecs = [co.random_field_elems(fr, m, 700 + i, False) for i, m in enumerate(lens)]
    vecs[4][:] = vecs[4][0]  # every scalar equal: one bucket per window takes everything
    offs = [0, 0, 299, 5, 1, 0, 100, 0, 7, 295, 0, 20, 300, 0, 172, 298, 0, 50, 0]
    out, infs = ctx.msm_batch(key, vecs, offs)
    width = out.shape[1]
    for i, (v, o) in enumerate(zip(vecs, offs)):
        wxy, winf = co.msm(curve, pts[o:], v) if len(v) else (np.zeros(width, np.uint64), True)
        assert infs[i] == winf and np.array_equal(out[i], wxy), i
    key.free()


@pytest.mark.parametrize("curve,log_n", [("bn254", 16), ("bls12_381", 14)])
def test_msm_medium_vs_oracle_pippenger(ctx, co, curve, log_n):
    fr = "bn254_fr" if curve == "bn254" else "bls12_381_fr"
    n = (1 << log_n) + 3
    _, pts = _points(ctx, co, curve, n, 21)
    s = co.random_field_elems(fr, n, 22, False)
    _check(ctx, co, curve, pts, s)


@pytest.mark.parametrize("log_n", [18, 20])
def test_msm_large_known_beta_identity(ctx, co, py, log_n):
    """At BASELINE sizes the oracle MSM is too slow for a unit test; use the size-independent
    identity of a KZG key with known beta: commit(p) == p(beta) * G."""
    cv = py.BN254
    n = (1 << log_n) + 3
    beta = py.random_field_elems(cv.fr, 1, seed=99)[0]
    key = ctx.generate_srs_for_testing("bn254", beta, n)
    # spot-check the generated key against the oracle's gen_srs
    want = co.gen_srs("bn254", co.ints_to_limbs([beta], 4)[0], 4)
    assert np.array_equal(key.read(0, 4), want)
    i = n - 2
    bi = pow(beta, i, cv.fr.p)
    assert np.array_equal(key.read(i, 1), co.fixed_base_mul("bn254", co.ints_to_limbs([bi], 4)))
    coeffs = co.random_field_elems("bn254_fr", n, 5, True)
    xy, inf = ctx.msm(key, coeffs, montgomery=True)
    ev = co.poly_eval("bn254_fr", coeffs, co.ints_to_limbs([cv.fr.to_mont(beta)], 4)[0])
    ev_plain = co.field_op("bn254_fr", "from_mont", ev[None, :])
    want_pt = co.fixed_base_mul("bn254", ev_plain)[0]
    assert not inf and np.array_equal(xy, want_pt)
    # linearity, the property the MPC prover relies on (shares: s = s1 + s2)
    s1 = co.random_field_elems("bn254_fr", n, 6, True)
    s2 = co.field_op("bn254_fr", "sub", coeffs, s1)
    p1, _ = ctx.msm(key, s1, montgomery=True)
    p2, _ = ctx.msm(key, s2, montgomery=True)
    fq = cv.fq
    P1 = tuple(fq.from_mont(v) for v in co.limbs_to_ints(p1.reshape(2, 4)))
    P2 = tuple(fq.from_mont(v) for v in co.limbs_to_ints(p2.reshape(2, 4)))
    S = cv.add(P1, P2)
    assert S == tuple(fq.from_mont(v) for v in co.limbs_to_ints(xy.reshape(2, 4)))
    key.free()


def test_msm_two_half_split_of_large_copied_scalars(ctx, co, py):
    """jf_msm cuts an MSM of >= 2^21 scalars that must be staged (pageable source) into two halves over consecutive base ranges
    (api.cu: jf_msm); a one-vector jf_msm_batch never splits.  Both must give the same point, also for an odd length with an
    offset, and the known-beta identity commit(p) == p(beta) G must hold."""
    cv = py.BN254
    n = (1 << 21) + 5
    beta = py.random_field_elems(cv.fr, 1, seed=123)[0]
    key = ctx.generate_srs_for_testing("bn254", beta, n)
    coeffs = co.random_field_elems("bn254_fr", n, 17, True)
    xy, inf = ctx.msm(key, coeffs, montgomery=True)
    one, one_inf = ctx.msm_batch(key, [coeffs], montgomery=True)
    assert not inf and not one_inf[0] and np.array_equal(xy, one[0])
    ev = co.poly_eval("bn254_fr", coeffs, co.ints_to_limbs([cv.fr.to_mont(beta)], 4)[0])
    want_pt = co.fixed_base_mul("bn254", co.field_op("bn254_fr", "from_mont", ev[None, :]))[0]
    assert np.array_equal(xy, want_pt)
    xy2, inf2 = ctx.msm(key, coeffs[: n - 8], montgomery=True, base_offset=3)
    one2, one2_inf = ctx.msm_batch(key, [coeffs[: n - 8]], base_offsets=[3], montgomery=True)
    assert inf2 == one2_inf[0] and np.array_equal(xy2, one2[0])
    key.free()


def test_msm_2_20_random_points_vs_oracle_pippenger(ctx, co):
    """BASELINE config-2 size with random (non-KZG) points: the GPU result against the C restatement of
    ark-ec's Pippenger on all 2^20 + 3 pairs (the known-beta identity above only covers KZG-shaped keys)."""
    n = (1 << 20) + 3
    ks = co.random_field_elems("bn254_fr", n, 777, False)
    pts = ctx.fixed_base_mul("bn254", ks)
    s = co.random_field_elems("bn254_fr", n, 778, False)
    key = ctx.load_srs("bn254", pts)
    xy, inf = ctx.msm(key, s)
    key.free()
    wxy, winf = co.msm("bn254", pts, s)
    assert inf == winf and np.array_equal(xy, wxy)


@pytest.mark.parametrize("shape", ["all_ones", "all_equal", "zeros_but_two", "small_integers", "constant_plus_random_tail"])
@pytest.mark.parametrize("skew_key", [False, True])
def test_msm_piled_up_and_sparse_scalars_at_2_17(ctx, co, py, shape, skew_key):
    """Scalars as a Lagrange-basis key meets them (witness VALUES): one bucket holding the whole MSM (several segments of the heavy
    bucket sums), a handful of entries scattered over 65536 buckets (a thread's chunk spans thousands of empty buckets), small
    integers (eight buckets of the second window hold an eighth of the MSM each).  Checked with the known-beta identity
    commit(p) == p(beta) G on a plain key and on a key marked as skewed (the Lagrange-basis key of the same size: then the
    scalars are read as values on the domain, and the identity becomes commit == p_interp(beta) G)."""
    fr = py.BN254_FR
    p = fr.p
    log_n = 17
    n = 1 << log_n
    beta = 0x5DEECE66D1234567890ABCDEF % p
    key = ctx.generate_srs_for_testing("bn254", beta, n + 3)
    use = key.lagrange(log_n) if skew_key else key
    rng = np.random.default_rng(7)
    s = np.zeros((n, 4), dtype=np.uint64)
    if shape == "all_ones":
        s[:, 0] = 1
    elif shape == "all_equal":
        s[:] = co.random_field_elems("bn254_fr", 1, 5, False)[0]
    elif shape == "zeros_but_two":
        s[[17, n - 3]] = co.random_field_elems("bn254_fr", 2, 6, False)
    elif shape == "small_integers":
        s[:, 0] = np.arange(n, dtype=np.uint64)
    else:
        s[:, 0] = 3
        s[n - 2:] = co.random_field_elems("bn254_fr", 2, 8, False)
    xy, inf = ctx.msm(use, s)
    vals = co.limbs_to_ints(s)
    if skew_key:   # the scalars are values on H: p_interp(beta) = sum_j v_j L_j(beta), L_j(beta) = w^j (beta^n - 1) / (n (beta - w^j))
        w = py.Radix2Domain(fr, n).group_gen
        c = (pow(beta, n, p) - 1) * pow(n, -1, p) % p
        acc, wj = 0, 1
        for v in vals:
            if v:
                acc = (acc + v * wj % p * pow((beta - wj) % p, -1, p)) % p
            wj = wj * w % p
        want_scalar = acc * c % p
    else:
        want_scalar = py.poly_eval(fr, vals, beta)
    want = co.fixed_base_mul("bn254", co.ints_to_limbs([want_scalar], 4))[0]
    assert not inf and np.array_equal(xy, want), (shape, skew_key)
    if skew_key:
        use.free()
    key.free()
