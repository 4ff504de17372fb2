"""CPU suite, part 3: the device field arithmetic's carry-chain sequence, executed through its
portable C fallback (same op list as the PTX, see csrc/gen_chains.py), against exact integers."""
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hfc(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("hfc") / "host_field_check")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "host_field_check.cpp")])

    def run(field, op, pairs):
        inp = "%s %s %d\n" % (field, op, len(pairs)) + "".join("%x %x\n" % p for p in pairs)
        out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split()
        return [int(x, 16) for x in out]
    return run


def test_generated_chains_are_current():
    csrc = os.path.join(ROOT, "mpc-jellyfish_b200", "csrc")
    out = subprocess.run(["python3", os.path.join(csrc, "gen_chains.py")], capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(csrc, "mont_chains.cuh")).read()


@pytest.mark.parametrize("fname", ["bn254_fr", "bn254_fq", "bls12_381_fr", "bls12_381_fq"])
def test_device_field_algorithm_on_host(py, hfc, fname):
    f = py.FIELDS[fname]
    c = hfc(fname, "consts", [])
    assert c == [f.R, f.R2, f.p, f.inv32]
    rnd = random.Random(3)
    edge = [0, 1, 2, f.p - 1, f.p - 2, f.R, f.R2, f.p >> 1, (1 << (f.bits - 1)) % f.p]
    vals = edge + [rnd.randrange(f.p) for _ in range(100)]
    pairs = [(a, b) for a in edge for b in edge] + [(rnd.choice(vals), rnd.choice(vals)) for _ in range(1500)]
    ri = pow(f.R, -1, f.p)
    assert hfc(fname, "mul", pairs) == [a * b * ri % f.p for a, b in pairs]
    # dedicated squaring (36 / 78 wide products; the doubled operand needs two spare bits in the top limb): operands
    # with saturated limbs exercise every carry the skipped products leave behind
    n32 = (f.bits + 31) // 32
    sat = [sum(((0xFFFFFFFF if rnd.random() < 0.8 else rnd.getrandbits(32)) << (32 * i)) for i in range(n32)) % f.p for _ in range(600)]
    sq = [(a, 0) for a in vals + sat + [rnd.randrange(f.p) for _ in range(3000)]]
    assert hfc(fname, "sqr", sq) == [a * a * ri % f.p for a, _ in sq]
    # a b + c d under one reduction (mul_add; c = a + b, d = a - b derived in the harness)
    mp = pairs + [(a, rnd.choice(sat)) for a in sat]
    assert hfc(fname, "madd", mp) == [(a * b + ((a + b) % f.p) * ((a - b) % f.p)) * ri % f.p for a, b in mp]
    assert hfc(fname, "add", pairs) == [(a + b) % f.p for a, b in pairs]
    assert hfc(fname, "sub", pairs) == [(a - b) % f.p for a, b in pairs]
    assert hfc(fname, "neg", pairs) == [(-a) % f.p for a, _ in pairs]
    assert hfc(fname, "to_mont", pairs[:40]) == [f.to_mont(a) for a, _ in pairs[:40]]
    iv = hfc(fname, "inv", pairs[:20])
    assert all(a == 0 or f.from_mont(a) * f.from_mont(i) % f.p == 1 for (a, _), i in zip(pairs[:20], iv))
