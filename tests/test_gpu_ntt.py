"""K3/K4 parity: GPU NTT vs the CPU oracle (arkworks Radix2EvaluationDomain semantics),
bit-exact Montgomery limbs, natural order in and out."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ["bn254_fr", "bls12_381_fr"]


def _offset(py, co, fname):
    f = py.FIELDS[fname]
    return co.ints_to_limbs([f.to_mont(f.generator)], 4)[0]


@pytest.mark.parametrize("fname", FIELDS)
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 18, 19])
def test_ntt_all_modes_vs_oracle(ctx, co, py, fname, log_n):
    n = 1 << log_n
    x = co.random_field_elems(fname, n, 100 + log_n, True)
    off = _offset(py, co, fname)
    for inverse in (False, True):
        for o in (None, off):
            want = co.ntt(fname, x, log_n, inverse, o)
            got = ctx.ntt(fname, x.copy(), log_n, inverse, o)
            assert np.array_equal(got, want), (fname, log_n, inverse, o is not None)


@pytest.mark.parametrize("fname", FIELDS)
def test_ntt_small_vs_bigint_definition(ctx, co, py, fname):
    """Against the O(n^2) definition out[i] = sum_j c_j (g w^i)^j in exact integers."""
    f = py.FIELDS[fname]
    for log_n, in_len in ((3, 8), (4, 11), (6, 64), (7, 3)):
        n = 1 << log_n
        vals = py.random_field_elems(f, in_len, seed=31 + log_n)
        x = np.zeros((n, 4), dtype=np.uint64)
        x[:in_len] = co.ints_to_limbs([f.to_mont(v) for v in vals], 4)
        d = py.Radix2Domain(f, n, f.generator)
        got = ctx.ntt(fname, x, log_n, False, _offset(py, co, fname), in_len=in_len)
        assert [f.from_mont(v) for v in co.limbs_to_ints(got)] == d.fft_naive(vals)


@pytest.mark.parametrize("fname", FIELDS)
@pytest.mark.parametrize("log_n,in_len", [(10, 131), (13, 1027), (15, 4099), (16, 1)])
def test_coset_fft_zero_padded_input(ctx, co, py, fname, log_n, in_len):
    """PLONK shape (prover.rs:552-567): <= n+3 coefficients evaluated over the 8n coset; entries
    past in_len must be ignored, whatever the buffer holds."""
    n = 1 << log_n
    x = co.random_field_elems(fname, n, 7, True)
    off = _offset(py, co, fname)
    want = co.ntt(fname, x, log_n, False, off, in_len=in_len)
    got = ctx.ntt(fname, x.copy(), log_n, False, off, in_len=in_len)
    assert np.array_equal(got, want)
    z = x.copy()
    z[in_len:] = 0
    assert np.array_equal(ctx.ntt(fname, z, log_n, False, off), want)


@pytest.mark.parametrize("fname", FIELDS)
def test_batched_ntt(ctx, co, py, fname):
    log_n, batch = 12, 5
    n = 1 << log_n
    x = co.random_field_elems(fname, n * batch, 3, True).reshape(batch, n, 4)
    off = _offset(py, co, fname)
    for inverse in (False, True):
        want = co.ntt(fname, x, log_n, inverse, off)
        got = ctx.ntt(fname, x.copy(), log_n, inverse, off)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("fname,log_n", [("bn254_fr", 22), ("bls12_381_fr", 22), ("bn254_fr", 23)])
def test_large_roundtrip_and_spot_check(ctx, co, py, fname, log_n):
    """BASELINE sizes: ifft(fft(x)) == x (the reference's own property, constraint_system.rs:2028-2034)
    plus Horner spot checks of individual outputs against exact integers."""
    f = py.FIELDS[fname]
    n = 1 << log_n
    in_len = n // 8 + 3
    x = np.zeros((n, 4), dtype=np.uint64)
    x[:in_len] = co.random_field_elems(fname, in_len, 77, True)
    off = _offset(py, co, fname)
    y = ctx.ntt(fname, x.copy(), log_n, False, off, in_len=in_len)
    d = py.Radix2Domain(f, n, f.generator)
    for i in (0, 1, 12345, n // 2 + 1, n - 1):
        pt = co.ints_to_limbs([f.to_mont(d.element(i))], 4)[0]
        assert np.array_equal(y[i], co.poly_eval(fname, x[:in_len], pt))
    back = ctx.ntt(fname, y, log_n, True, off)
    assert np.array_equal(back, x)


def test_domain_too_large_and_bad_args(ctx):
    import mpc_jellyfish_b200 as jf
    with pytest.raises(jf.DomainCreationError):
        ctx.ntt("bn254_fr", np.zeros((1, 4), dtype=np.uint64), 29)
    with pytest.raises(jf.DomainCreationError):
        jf.Radix2EvaluationDomain(ctx, "bn254_fr", (1 << 28) + 1)
    with pytest.raises(jf.InvalidParameters):
        ctx.ntt("bn254_fq", np.zeros((8, 4), dtype=np.uint64), 3)


@pytest.mark.parametrize("fname", FIELDS)
def test_domain_mirror(ctx, co, py, fname):
    """Radix2EvaluationDomain mirror: new / get_coset / fft / ifft / element."""
    import mpc_jellyfish_b200 as jf
    f = py.FIELDS[fname]
    dom = jf.Radix2EvaluationDomain(ctx, fname, 1000)
    assert dom.size == 1024
    coset = dom.get_coset(f.generator)
    ref = py.Radix2Domain(f, 1000, f.generator)
    assert coset.element(5) == ref.element(5) and dom.group_gen == ref.group_gen
    vals = py.random_field_elems(f, 700, seed=5)
    c = co.ints_to_limbs([f.to_mont(v) for v in vals], 4)
    ev = coset.fft(c)
    assert [f.from_mont(v) for v in co.limbs_to_ints(ev)] == ref.fft(vals)
    back = coset.ifft(ev)
    assert np.array_equal(back[:700], c) and not back[700:].any()


@pytest.mark.parametrize("fname,log_n,inverse", [("bn254_fr", 22, False), ("bls12_381_fr", 22, True), ("bn254_fr", 21, True)])
def test_full_size_output_equals_the_oracle(ctx, co, py, fname, log_n, inverse):
    """BASELINE config-3 sizes: every one of the 2^22 outputs against the C restatement (coset domain)."""
    f = py.FIELDS[fname]
    n = 1 << log_n
    x = co.random_field_elems(fname, n, 4242 + log_n, True)
    off = co.ints_to_limbs([f.to_mont(f.generator)], 4)[0]
    got = ctx.ntt(fname, x.copy(), log_n, inverse, off)
    assert np.array_equal(got, co.ntt(fname, x, log_n, inverse, off))


def _sub_offsets(py, co, fname, log_n, rows):
    """g * w_8n^r, r < rows: the sub-cosets of the prover's quotient coset g<w_8n> (prover.rs:545)."""
    f = py.FIELDS[fname]
    w8n = py.Radix2Domain(f, 8 << log_n).group_gen
    offs = [f.generator * pow(w8n, r, f.p) % f.p for r in range(rows)]
    return offs, co.ints_to_limbs([f.to_mont(o) for o in offs], 4)


@pytest.mark.parametrize("fname", FIELDS)
@pytest.mark.parametrize("log_n,in_len", [(3, 8), (3, 11), (5, 35), (9, 515), (10, 1024), (12, 4099), (14, 3), (18, (1 << 18) + 3)])
def test_ntt_cosets_vs_oracle(ctx, co, py, fname, log_n, in_len):
    """jf_ntt_cosets forward: each row is `get_coset(g w_8n^r).fft` of the polynomial reduced mod X^n - (g w_8n^r)^n."""
    f = py.FIELDS[fname]
    n = 1 << log_n
    rows, polys = 6, 3
    offs, off_limbs = _sub_offsets(py, co, fname, log_n, rows)
    x = co.random_field_elems(fname, polys * in_len, 900 + log_n, True).reshape(polys, in_len, 4)
    got = ctx.ntt_cosets(fname, x, log_n, off_limbs)
    assert got.shape == (polys, rows, n, 4)
    for p in range(polys):
        for r in range(rows):
            folded = np.zeros((n, 4), dtype=np.uint64)
            folded[:min(n, in_len)] = x[p, :n]
            if in_len > n:
                c = co.ints_to_limbs([f.to_mont(pow(offs[r], n, f.p))], 4)
                hi = x[p, n:]
                cc = np.repeat(c, hi.shape[0], axis=0)
                folded[:in_len - n] = co.field_op(fname, "add", folded[:in_len - n], co.field_op(fname, "mul", hi, cc))
            want = co.ntt(fname, folded, log_n, False, off_limbs[r])
            assert np.array_equal(got[p, r], want), (fname, log_n, p, r)


@pytest.mark.parametrize("fname", FIELDS)
@pytest.mark.parametrize("log_n", [3, 4, 8, 9, 10, 13, 18])
def test_intt_cosets_vs_oracle(ctx, co, py, fname, log_n):
    n = 1 << log_n
    rows, polys = 6, 2
    _, off_limbs = _sub_offsets(py, co, fname, log_n, rows)
    x = co.random_field_elems(fname, polys * rows * n, 77 + log_n, True).reshape(polys, rows, n, 4)
    got = ctx.intt_cosets(fname, x.copy(), log_n, off_limbs)
    for p in range(polys):
        for r in range(rows):
            assert np.array_equal(got[p, r], co.ntt(fname, x[p, r], log_n, True, off_limbs[r])), (fname, log_n, p, r)


@pytest.mark.parametrize("fname,log_n", [("bn254_fr", 12), ("bls12_381_fr", 10), ("bn254_fr", 19)])
def test_ntt_cosets_tile_the_8n_coset(ctx, co, py, fname, log_n):
    """The reference's own call: `quot_domain.get_coset(GENERATOR).fft` over 8n points (prover.rs:552-567).
    Row r of the sub-coset form is its outputs r, r + 8, r + 16, ..."""
    n = 1 << log_n
    in_len = n + 3
    x = co.random_field_elems(fname, in_len, 5151, True)
    _, off_limbs = _sub_offsets(py, co, fname, log_n, 8)
    rows = ctx.ntt_cosets(fname, x[None], log_n, off_limbs)[0]
    big = np.zeros((8 * n, 4), dtype=np.uint64)
    big[:in_len] = x
    full = ctx.ntt(fname, big, log_n + 3, False, _offset(py, co, fname), in_len=in_len)
    for r in range(8):
        assert np.array_equal(rows[r], full[r::8]), r


def test_ntt_cosets_bad_args(ctx):
    import mpc_jellyfish_b200 as jf
    one = np.zeros((1, 4), dtype=np.uint64)
    with pytest.raises(jf.InvalidParameters):
        ctx.ntt_cosets("bn254_fr", np.zeros((1, 17, 4), dtype=np.uint64), 3, one)   # in_len > 2n
    with pytest.raises(jf.InvalidParameters):
        ctx.ntt_cosets("bn254_fr", np.zeros((1, 4, 4), dtype=np.uint64), 2, one)    # domain below 8
    with pytest.raises(jf.DomainCreationError):
        ctx.ntt_cosets("bn254_fr", np.zeros((1, 4, 4), dtype=np.uint64), 29, one)
