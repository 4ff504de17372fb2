"""CPU suite: the algebra behind round 3 of the device prover ("Round 3 on sub-cosets"), stated with the
big-int oracle only.  The reference evaluates the quotient on the 8n-point coset g<w_8n>
(plonk/src/proof_system/prover.rs:545,552-567,672); the device prover uses six of its eight sub-cosets
(g w_8n^r)<w_n>.  These tests pin the three facts that make both forms produce the same polynomial."""
import random

import pytest


def _setup(py, fname, log_n):
    f = py.FIELDS[fname]
    n = 1 << log_n
    big = py.Radix2Domain(f, 8 * n, f.generator)                       # quot_domain.get_coset(GENERATOR)
    offs = [f.generator * pow(big.group_gen, r, f.p) % f.p for r in range(8)]
    return f, n, big, offs


def _solve_from_rows(py, f, n, offs, evals8):
    """Coefficients (6n of them) of the polynomial through the values of rows r < 6, as the device prover computes them:
    six size-n coset iNTTs, then the inverse of V[r][k] = c_r^k (c_r = (g w_8n^r)^n) applied per coefficient index."""
    cs = [pow(offs[r], n, f.p) for r in range(6)]
    T = [py.Radix2Domain(f, n, offs[r]).ifft(evals8[r::8]) for r in range(6)]
    V = [[pow(cs[r], k, f.p) for k in range(6)] + [int(r == k) for k in range(6)] for r in range(6)]
    for col in range(6):  # Gauss-Jordan elimination mod p
        piv = next(r for r in range(col, 6) if V[r][col])
        V[col], V[piv] = V[piv], V[col]
        inv = pow(V[col][col], -1, f.p)
        V[col] = [v * inv % f.p for v in V[col]]
        for r in range(6):
            if r != col and V[r][col]:
                m = V[r][col]
                V[r] = [(a - m * b) % f.p for a, b in zip(V[r], V[col])]
    out = [0] * (6 * n)
    for j in range(n):
        for k in range(6):
            out[k * n + j] = sum(V[k][6 + r] * T[r][j] for r in range(6)) % f.p
    return out


@pytest.mark.parametrize("fname", ["bn254_fr", "bls12_381_fr"])
@pytest.mark.parametrize("log_n", [3, 4])
def test_rows_of_the_sub_coset_form_are_the_residue_classes_of_the_8n_coset_fft(py, fname, log_n):
    """coset.fft(p)[8 i + r] == get_coset(g w_8n^r).fft(p mod (X^n - c_r))[i],  c_r = (g w_8n^r)^n."""
    f, n, big, offs = _setup(py, fname, log_n)
    rnd = random.Random(5 + log_n)
    p = [rnd.randrange(f.p) for _ in range(n + 3)]                      # a masked permutation polynomial has n + 3 coefficients
    full = big.fft(p)
    for r in range(8):
        c = pow(offs[r], n, f.p)
        folded = [(p[j] + (c * p[j + n] if j + n < len(p) else 0)) % f.p for j in range(n)]
        row = py.Radix2Domain(f, n, offs[r]).fft(folded)
        assert row == full[r::8]
        # w_n x stays in the row: the "next" evaluation z(w x) of the permutation argument is a shift by one
        nxt = [py.poly_eval(f, p, offs[r] * pow(big.group_gen, 8 * ((i + 1) % n), f.p) % f.p) for i in range(n)]
        assert nxt == row[1:] + row[:1]


@pytest.mark.parametrize("fname", ["bn254_fr", "bls12_381_fr"])
@pytest.mark.parametrize("log_n", [3, 4])
def test_six_sub_cosets_determine_the_quotient(py, fname, log_n):
    """t of degree 5n + 7 (quotient_polynomial_degree, prover.rs:1126-1128) from its values on rows r < 6: the row
    interpolants are T_r = sum_k c_r^k t_k (t = sum_k X^(k n) t_k), a 6 x 6 Vandermonde system per coefficient index;
    the result equals what the reference's 8n-point coset.ifft returns."""
    f, n, big, offs = _setup(py, fname, log_n)
    rnd = random.Random(11 + log_n)
    t = [rnd.randrange(f.p) for _ in range(5 * n + 8)]
    assert len(t) <= 6 * n                                              # n >= 8; smaller domains keep the 8n form
    evals8 = big.fft(t)
    assert py.poly_strip(big.ifft(evals8)) == t                         # the reference's path
    assert py.poly_strip(_solve_from_rows(py, f, n, offs, evals8)) == t


def test_values_that_are_not_a_low_degree_polynomial_fail_the_degree_check(py):
    """An unsatisfied witness makes the pointwise quotient values those of no polynomial of degree 5n + 7: the six-row
    interpolant then has non-zero coefficients above that degree (WrongQuotientPolyDegree, prover.rs:916-919)."""
    f, n, big, offs = _setup(py, "bn254_fr", 4)
    rnd = random.Random(3)
    t = [rnd.randrange(f.p) for _ in range(5 * n + 8)]
    evals8 = big.fft(t)
    evals8[5] = (evals8[5] + 1) % f.p                                   # one corrupted value in row 5
    got = _solve_from_rows(py, f, n, offs, evals8)
    deg = 5 * (n + 1) + 2
    assert len(got[deg + 1:]) == n - 8 and any(got[deg + 1:])           # degree_check_kernel's condition
    # n = 8 would leave no coefficient above the degree (6 n == 5 n + 8): the device prover keeps the 8n form below n = 16
    # and it is the interpolant: it reproduces the values of the six rows, the corrupted one included
    for r in range(6):
        for i in (0, 1, n - 1):
            x = offs[r] * pow(big.group_gen, 8 * i, f.p) % f.p
            assert py.poly_eval(f, got, x) == evals8[8 * i + r]
