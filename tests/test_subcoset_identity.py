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
    """Coefficients (6n of them) of the polynomial through the values of rows r < 6, as the device prover computes them."""
    got = _solve_from_rows(py, f, n, offs, evals8)
    assert py.poly_strip(got) == t


def test_values_that_are_not_a_low_degree_polynomial_fail_the_degree_check(py):
    """An unsatisfied witness makes the pointwise quotient values those of no polynomial of degree 5n + 7: the six-row
    interpolant then has non-zero coefficients above that degree (WrongQuotientPolyDegree, prover.rs:916-919)."""
    f, n, big, offs = _setup(py, "bn254_fr", 3)
    rnd = random.Random(3)
    t = [rnd.randrange(f.p) for _ in range(5 * n + 8)]
    evals8 = big.fft(t)
    evals8[5] = (evals8[5] + 1) % f.p                                   # one corrupted value in row 5
    got = _solve_from_rows(py, f, n, offs, evals8)
    deg = 5 * (n + 1) + 2
    assert any(got[deg + 1:])                                           # degree_check_kernel's condition
    # and it is the interpolant: it reproduces every value of the six rows, the corrupted one included
    for r in range(6):
        for i in (0, 1, n - 1):
            x = offs[r] * pow(big.group_gen, 8 * i, f.p) % f.p
            assert py.poly_eval(f, got, x) == evals8[8 * i + r]
