"""GPU suite: the multi-GPU forms of the MSM and of batched NTTs (csrc/comm.cu).

On one GPU: a single-rank `jf_comm` (both transports) and a `jf_group` that lists device 0 twice (two contexts, two slices,
the same code path as two GPUs: there is no kernel in it that waits for another).  With >= 2 GPUs visible: the group on two
devices and the one-process-per-GPU run under torchrun + NCCL (tests/multi/sharded_worker.py)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BETA = 0xDEADBEEF1234567890ABCDEF
R_BN = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _known_beta(co, curve, fr, scalars, beta):
    beta_m = co.field_op(fr, "to_mont", co.ints_to_limbs([beta], 4))[0]
    ev = co.poly_eval(fr, co.field_op(fr, "to_mont", scalars), beta_m)
    return co.fixed_base_mul(curve, co.field_op(fr, "from_mont", ev[None, :]))[0]


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_single_rank_comm_matches_jf_msm(ctx, co, transport):
    import mpc_jellyfish_b200 as jf
    comm = jf.Comm(ctx, 0, 1, jf.Comm.unique_id(), transport)
    assert comm.transport == transport
    n = 5000
    key = ctx.generate_srs_for_testing("bn254", BETA, n)
    s = co.random_field_elems("bn254_fr", n, 11, False)
    want = ctx.msm(key, s)
    for _ in range(3):
        got = comm.msm(key, s)
        assert got[1] == want[1] and np.array_equal(got[0], want[0])
    assert np.array_equal(want[0], _known_beta(co, "bn254", "bn254_fr", s, BETA))
    comm.close()
    key.free()


def _group_checks(co, devices):
    import mpc_jellyfish_b200 as jf
    g = jf.Group(devices)
    assert len(g) == len(devices)
    for curve, fr, n in (("bn254", "bn254_fr", 6001), ("bls12_381", "bls12_381_fr", 2048), ("bn254", "bn254_fr", 1)):
        mod = R_BN if curve == "bn254" else 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
        key = g.generate_srs_for_testing(curve, BETA % mod, n)
        s = co.random_field_elems(fr, n, 21 + n, False)
        xy, inf = g.msm(key, s)
        assert not inf and np.array_equal(xy, _known_beta(co, curve, fr, s, BETA % mod))
        # base_offset / shorter scalar vectors: `msm_bigint(&powers_of_g[off..], scalars)` semantics
        if n > 100:
            off, m = 37, n // 2
            xy2, inf2 = g.msm(key, s[:m], base_offset=off)
            shifted = np.concatenate([np.zeros((off, 4), dtype=np.uint64), s[:m]])
            assert not inf2 and np.array_equal(xy2, _known_beta(co, curve, fr, shifted, BETA % mod))
            xy3, inf3 = g.msm(key, np.concatenate([s, s]))        # more scalars than points: truncated like arkworks
            assert np.array_equal(xy3, xy)
        z, zinf = g.msm(key, np.zeros((n, 4), dtype=np.uint64))
        assert zinf and not z.any()
        with pytest.raises(jf.InvalidParameters):
            bad = s.copy()
            bad[n - 1] = np.array([0xFFFFFFFFFFFFFFFF] * 4, dtype=np.uint64)
            g.msm(key, bad)
        key.free()
    # loaded (not generated) key: slices of the caller's point array
    n = 3000
    ks = co.random_field_elems("bn254_fr", n, 5, False)
    pts = co.fixed_base_mul("bn254", ks)
    key = g.load_srs("bn254", pts)
    s = co.random_field_elems("bn254_fr", n, 6, False)
    xy, inf = g.msm(key, s)
    wxy, winf = co.msm("bn254", pts, s)
    assert inf == winf and np.array_equal(xy, wxy)
    key.free()
    # batched coset NTT dealt out by polynomial == the oracle's transform of every vector
    for batch in (1, 5, 16):
        log_n = 12
        x = co.random_field_elems("bls12_381_fr", batch << log_n, 31 + batch, True).reshape(batch, 1 << log_n, 4)
        off = co.field_op("bls12_381_fr", "to_mont", np.array([[7, 0, 0, 0]], dtype=np.uint64))[0]
        y = g.ntt("bls12_381_fr", x.copy(), log_n, False, off)
        for b in range(batch):
            assert np.array_equal(y[b], co.ntt("bls12_381_fr", x[b].copy(), log_n, False, off)), "vector %d" % b
        back = g.ntt("bls12_381_fr", y, log_n, True, off)
        assert np.array_equal(back, x)
    g.close()


def test_group_on_one_device_listed_twice(co):
    _group_checks(co, [0, 0])


def test_group_on_three_contexts_of_one_device(co):
    _group_checks(co, [0, 0, 0])


def test_group_on_two_devices(co):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _group_checks(co, [0, 1])


def test_sharded_msm_one_process_per_gpu_nccl_and_p2p():
    n = _ngpu()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    world = 2 if n < 4 else 4
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "multi", "sharded_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for k in range(world):
        assert "rank %d ok" % k in r.stdout


def test_sharded_prover_single_rank(ctx, co, py):
    """`jf_plonk_pk_shard_commits` with one rank (slice == the whole key): the exchange path runs, the proof is unchanged"""
    import random
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mpc_jellyfish_b200 as jf
    import plonk_ref as P
    import plonk_util as U
    cv, fr = py.BN254, py.BN254_FR
    cs = P.gen_circuit_for_test(20, 1)
    arr = U.arrays_from_oracle_circuit(co, py, cs)
    key = ctx.generate_srs_for_testing("bn254", BETA, cs.n + 3)
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], arr["pub_gate_ids"])
    rnd = random.Random(5)
    bl = co.ints_to_limbs([fr.to_mont(rnd.randrange(fr.p)) for _ in range(17)], 4)
    want = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity").serialize_compressed()
    for transport in ("p2p", "nccl"):
        comm = jf.Comm(ctx, 0, 1, jf.Comm.unique_id(), transport)
        pk.shard_commits(comm, key, 0)
        for _ in range(2):
            assert jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity").serialize_compressed() == want
        pk.shard_commits(None, None)
        comm.close()
    assert jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity").serialize_compressed() == want
    pk.free()
    key.free()


def test_sharded_prover_one_process_per_gpu():
    n = _ngpu()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    world = 2 if n < 4 else 4
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "multi", "prove_worker.py")],
                       capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for k in range(world):
        assert "rank %d ok" % k in r.stdout


def test_sharded_prover_one_process_one_thread_per_gpu(co, py):
    """The form a single Rust prover process needs: ONE process, one thread (context + `jf_comm` over NCCL) per GPU, every thread
    calling the same prove with the same inputs.  Ranks of one process reach each other's mailbox through peer access (CUDA IPC
    only maps memory of ANOTHER process), so the peer-memory transport is available here too; both transports are run.
    Bytes == the CPU restatement on every thread."""
    import random
    import threading
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mpc_jellyfish_b200 as jf
    import plonk_ref as P
    import plonk_util as U
    from mpc_jellyfish_b200.sharded import shard_range
    cv, fr = py.BN254, py.BN254_FR
    world = 2
    cs = P.gen_circuit_for_bench(1 << 10)
    arr = U.arrays_from_oracle_circuit(co, py, cs)
    beta = BETA % fr.p
    opk = P.preprocess(cv, P.gen_srs(cv, beta, cs.n + 2), cs)
    rnd = random.Random(9)
    ints = [rnd.randrange(fr.p) for _ in range(17)]
    bl = co.ints_to_limbs([fr.to_mont(v) for v in ints], 4)
    want = P.serialize_proof(cv, P.prove(cv, cs, opk, ints, "solidity"))
    results, errors = [None] * world, []

    def worker(rank, uid, transport):
        try:
            ctx = jf.Context(rank)
            comm = jf.Comm(ctx, rank, world, uid, transport)
            key = ctx.generate_srs_for_testing("bn254", beta, cs.n + 3)
            a, b = shard_range(cs.n + 3, world, rank)
            key_slice = ctx.generate_srs_for_testing("bn254", beta, b - a, first_power=a)
            pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                             arr["pub_gate_ids"])
            pk.shard_commits(comm, key_slice, a)
            out = [jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity").serialize_compressed() for _ in range(3)]
            results[rank] = (comm.transport, out)
            pk.shard_commits(None, None)
            pk.free()
            key_slice.free()
            key.free()
            comm.close()
            ctx.close()
        except Exception as e:  # noqa: BLE001
            errors.append((rank, repr(e)))

    for transport, expect in (("auto", "p2p"), ("nccl", "nccl")):
        uid = jf.Comm.unique_id()
        threads = [threading.Thread(target=worker, args=(r, uid, transport)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=600)
        assert not errors, errors
        for r in range(world):
            assert results[r] is not None and results[r][0] == expect, (transport, results[r] and results[r][0])
            assert all(o == want for o in results[r][1]), "rank %d: sharded proof differs (%s)" % (r, transport)
            results[r] = None
