"""CPU suite, part 1: pin the oracle.  The big-int model (oracle/pyref.py) is checked against the
public constants / known-answer points of tests/golden/constants.json, then the C restatement
(oracle/jf_oracle.c) against the big-int model on seeded inputs, edge cases included."""
import json
import os
import random

import numpy as np
import pytest

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "constants.json")))


def _int(x):
    return int(x, 0) if isinstance(x, str) else int(x)


@pytest.mark.parametrize("fname", ["bn254_fr", "bn254_fq", "bls12_381_fr", "bls12_381_fq"])
def test_field_constants_match_golden(py, fname):
    g, f = GOLDEN[fname], py.FIELDS[fname]
    assert f.p == _int(g["modulus"])
    assert f.inv64 == _int(g["inv64"]) and f.inv32 == _int(g["inv64"]) & 0xFFFFFFFF
    assert f.R == _int(g["R"]) and f.R2 == _int(g["R2"])
    if "two_adicity" in g:
        assert f.two_adicity == g["two_adicity"] and f.generator == g["generator"]
        assert (f.p - 1) % (1 << f.two_adicity) == 0 and ((f.p - 1) >> f.two_adicity) % 2 == 1
        w = f.two_adic_root
        assert w == _int(g["two_adic_root_of_unity"])
        assert pow(w, 1 << f.two_adicity, f.p) == 1 and pow(w, 1 << (f.two_adicity - 1), f.p) == f.p - 1


def test_curve_known_answers(py):
    g = GOLDEN["bn254_g1"]
    cv = py.BN254
    assert cv.b == g["b"] and cv.gen == tuple(_int(v) for v in g["generator"])
    assert cv.mul(2, cv.gen) == tuple(_int(v) for v in g["two_g"])
    assert cv.add(cv.gen, cv.gen) == cv.mul(2, cv.gen)
    assert cv.mul(cv.fr.p, cv.gen) is None and cv.mul(cv.fr.p - 1, cv.gen) == cv.neg(cv.gen)
    assert cv.serialize_compressed(cv.gen).hex() == g["generator_compressed"]
    assert cv.serialize_compressed(None).hex() == g["identity_compressed"]
    assert cv.serialize_compressed(cv.neg(cv.gen))[-1] & 0x80
    b = GOLDEN["bls12_381_g1"]
    cv = py.BLS12_381
    assert cv.gen == tuple(_int(v) for v in b["generator"]) and cv.b == b["b"]
    assert cv.is_on_curve(cv.gen) and cv.mul(cv.fr.p, cv.gen) is None
    # ark-bls12-381 0.4.0 serialises G1 in the ZCash / IETF form (x big-endian, flags in the top bits of byte 0)
    assert cv.serialize_compressed(cv.gen).hex() == b["generator_compressed"]
    assert cv.serialize_compressed(cv.mul(2, cv.gen)).hex() == b["two_g_compressed"]       # larger root: sort flag set
    assert cv.serialize_compressed(cv.mul(3, cv.gen)).hex() == b["three_g_compressed"]
    assert cv.serialize_compressed(None).hex() == b["identity_compressed"]
    neg = cv.serialize_compressed(cv.neg(cv.gen))
    assert neg[0] == 0xB7 and neg[1:] == cv.serialize_compressed(cv.gen)[1:]


@pytest.mark.parametrize("fname", ["bn254_fr", "bn254_fq", "bls12_381_fr", "bls12_381_fq"])
def test_c_oracle_field_ops_vs_bigint(py, co, fname):
    f = py.FIELDS[fname]
    rnd = random.Random(1)
    edge = [0, 1, f.p - 1, f.R, f.p >> 1]
    a = [x for x in edge for _ in edge] + [rnd.randrange(f.p) for _ in range(200)]
    b = [y for _ in edge for y in edge] + [rnd.randrange(f.p) for _ in range(200)]
    A, B = co.ints_to_limbs(a, f.limbs64), co.ints_to_limbs(b, f.limbs64)
    ri = pow(f.R, -1, f.p)
    assert co.limbs_to_ints(co.field_op(fname, "mul", A, B)) == [x * y * ri % f.p for x, y in zip(a, b)]
    assert co.limbs_to_ints(co.field_op(fname, "add", A, B)) == [(x + y) % f.p for x, y in zip(a, b)]
    assert co.limbs_to_ints(co.field_op(fname, "sub", A, B)) == [(x - y) % f.p for x, y in zip(a, b)]
    assert co.limbs_to_ints(co.field_op(fname, "neg", A)) == [(-x) % f.p for x in a]
    assert co.limbs_to_ints(co.field_op(fname, "to_mont", A)) == [f.to_mont(x) for x in a]
    assert co.limbs_to_ints(co.field_op(fname, "from_mont", A)) == [f.from_mont(x) for x in a]


@pytest.mark.parametrize("fname", ["bn254_fr", "bls12_381_fr"])
def test_c_oracle_ntt_vs_bigint(py, co, fname):
    f = py.FIELDS[fname]
    for log_n, in_len in ((0, 1), (1, 2), (2, 3), (3, 8), (6, 40), (9, 512), (10, 131)):
        n = 1 << log_n
        vals = py.random_field_elems(f, in_len, seed=7 + log_n)
        assert co.limbs_to_ints(co.random_field_elems(fname, in_len, 7 + log_n, False)) == vals
        x = np.zeros((n, 4), dtype=np.uint64)
        x[:in_len] = co.ints_to_limbs([f.to_mont(v) for v in vals], 4)
        for off in (None, f.generator):
            d = py.Radix2Domain(f, n, 1 if off is None else off)
            offl = None if off is None else co.ints_to_limbs([f.to_mont(off)], 4)[0]
            got = [f.from_mont(v) for v in co.limbs_to_ints(co.ntt(fname, x, log_n, False, offl, in_len=in_len))]
            assert got == d.fft(vals)
            if log_n <= 6:
                assert got == d.fft_naive(vals)
            full = py.random_field_elems(f, n, seed=99)
            X = co.ints_to_limbs([f.to_mont(v) for v in full], 4)
            got = [f.from_mont(v) for v in co.limbs_to_ints(co.ntt(fname, X, log_n, True, offl))]
            assert got == d.ifft(full)
            back = co.ntt(fname, co.ntt(fname, X, log_n, False, offl), log_n, True, offl)
            assert np.array_equal(back, X)


def test_domain_semantics(py):
    """Radix2EvaluationDomain::new rounds up, group_gen has exact order, 6n -> 8n (constants.rs:18-20)."""
    f = py.BN254_FR
    d = py.Radix2Domain(f, 6 * 1024)
    assert d.size == 8192 and pow(d.group_gen, 8192, f.p) == 1 and pow(d.group_gen, 4096, f.p) != 1
    with pytest.raises(ValueError):
        py.Radix2Domain(f, (1 << 28) + 1)
    c = d.get_coset(f.generator)
    assert c.element(3) == f.generator * pow(d.group_gen, 3, f.p) % f.p


@pytest.mark.parametrize("cname", ["bn254", "bls12_381"])
def test_c_oracle_msm_vs_bigint(py, co, cname):
    cv = py.CURVES[cname]
    L = cv.fq.limbs64
    for n in (1, 5, 31, 33, 120):
        s = py.random_field_elems(cv.fr, n, seed=n)
        if n > 4:
            s[0], s[1], s[2] = 0, cv.fr.p - 1, 1
        ks = py.random_field_elems(cv.fr, n, seed=1000 + n)
        pts = co.fixed_base_mul(cname, co.ints_to_limbs(ks, 4))
        ptsi = [cv.mul(k, cv.gen) for k in ks]
        got = list(zip((cv.fq.from_mont(v) for v in co.limbs_to_ints(pts[:, :L])),
                       (cv.fq.from_mont(v) for v in co.limbs_to_ints(pts[:, L:]))))
        assert got == ptsi
        xy, inf = co.msm(cname, pts, co.ints_to_limbs(s, 4))
        want = cv.msm_naive(s, ptsi)
        g = (cv.fq.from_mont(co.limbs_to_ints(xy[None, :L])[0]), cv.fq.from_mont(co.limbs_to_ints(xy[None, L:])[0]))
        assert (want is None and inf) or g == want
        assert cv.msm_pippenger(s, ptsi, 5) == want
    # identity result and identity points
    xy, inf = co.msm(cname, pts, co.ints_to_limbs([0] * n, 4))
    assert inf and not xy.any()


def test_kzg_restatement_known_beta(py, co):
    """commit == p(beta) G, open/verify identity (the reference's end_to_end test, mod.rs:407-443)."""
    cv, fr = py.BN254, py.BN254_FR
    beta = 987654321987654321
    srs = py.gen_srs_for_testing(cv, beta, 20)
    got = co.gen_srs("bn254", co.ints_to_limbs([beta], 4)[0], 21)
    assert [(cv.fq.from_mont(x), cv.fq.from_mont(y)) for x, y in
            zip(co.limbs_to_ints(got[:, :4]), co.limbs_to_ints(got[:, 4:]))] == srs
    coeffs = py.random_field_elems(fr, 18, seed=3)
    coeffs[0] = 0
    c = py.kzg_commit(cv, srs, coeffs)
    assert c == cv.mul(py.poly_eval(fr, coeffs, beta), cv.gen)
    proof, ev = py.kzg_open(cv, srs, coeffs, 4242)
    assert py.kzg_verify_known_beta(cv, beta, cv.gen, c, 4242, ev, proof)
    assert not py.kzg_verify_known_beta(cv, beta, cv.gen, c, 4242, (ev + 1) % fr.p, proof)
    with pytest.raises(ValueError):
        py.kzg_commit(cv, srs[:3], py.random_field_elems(fr, 10, seed=1))
