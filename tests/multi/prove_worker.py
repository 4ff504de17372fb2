"""One rank of the sharded-prover parity run (launched by torchrun, one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/multi/prove_worker.py

`jf_plonk_pk_shard_commits`: every rank runs the same prove call and commits only its slice of every polynomial against its slice
of the key; with shard_round3 the sub-cosets of round 3 are dealt out over the ranks too (interpolants broadcast before the solve).  Checked on EVERY rank: the proof bytes equal the CPU restatement's (TurboPlonk, UltraPlonk, a batch of two, both
transports, ragged slices incl. an empty one), and equal the one-GPU proof of the same key after un-sharding.
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

import coracle as co
import mpc_jellyfish_b200 as jf
import plonk_ref as P
import plonk_util as U
import pyref
from mpc_jellyfish_b200.sharded import Comm, shard_range

BETA = 0x1D3C7A5B9E8F60412B7A6C5D4E3F20198A7B6C5D4E3F2A1B0C9D8E7F6A5B4C3


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = jf.Context(local)
    cv, fr = pyref.BN254, pyref.BN254_FR
    beta = BETA % fr.p
    checks = 0
    for transport in ("p2p", "nccl"):
        comm = Comm.from_torch_distributed(ctx, transport=transport)
        for name, make, ultra in (("test_m20", lambda: P.gen_circuit_for_test(20, 1), False),
                                  ("tiny_n4", lambda: _tiny(1), False),          # n + 3 = 7 points over `world` ranks
                                  ("bench_2^10", lambda: P.gen_circuit_for_bench(1 << 10), False),
                                  ("ultra_m6", lambda: P.gen_circuit_for_test(6, 1, ultra=True), True)):
            cs = make()
            n = cs.n
            arr = U.arrays_from_oracle_circuit(co, pyref, cs)
            key = ctx.generate_srs_for_testing("bn254", beta, n + 3)
            a, b = shard_range(n + 3, world, rank)
            if name == "tiny_n4" and world > 1:   # a deliberately ragged split: the last rank holds nothing
                cut = [min(i * 4, n + 3) for i in range(world)] + [n + 3]
                a, b = cut[rank], cut[rank + 1]
            key_slice = ctx.generate_srs_for_testing("bn254", beta, b - a, first_power=a)
            if ultra:
                pk = jf.PlonkKzgSnark.preprocess_ultra(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                                       arr["pub_gate_ids"], arr["range_bit_len"], arr["table_key"], arr["table_dom_sep"],
                                                       arr["q_dom_sep"])
            else:
                # one key also carries the Lagrange-basis form: it is switched off while the key is sharded and back on afterwards
                pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                                 arr["pub_gate_ids"], lagrange_wire_commitments=(name == "bench_2^10"))
            opk = P.preprocess(cv, P.gen_srs(cv, beta, n + 2), cs)
            rnd = random.Random(31)
            ints = [rnd.randrange(fr.p) for _ in range(P.num_blinders(cs))]
            bl = co.ints_to_limbs([fr.to_mont(v) for v in ints], 4)
            prove = jf.PlonkKzgSnark.prove_ultra if ultra else jf.PlonkKzgSnark.prove
            for kind in ("solidity", "standard"):
                want = P.serialize_proof(cv, P.prove(cv, cs, opk, ints, kind))
                for rows in (False, True):                    # commitments only / commitments + round 3 by sub-coset
                    pk.shard_commits(comm, key_slice, a, shard_round3=rows)
                    for it in range(3):                       # back-to-back: both mailbox parities
                        got = prove(pk, arr["witness"], bl, kind).serialize_compressed()
                        assert got == want, "%s %s %s sharded proof differs (rows=%s, call %d)" % (transport, name, kind, rows, it)
                pk.shard_commits(None, None)
                assert prove(pk, arr["witness"], bl, kind).serialize_compressed() == want, "one-GPU proof after un-sharding"
            if name == "test_m20":                        # a batch of two instances through the sharded commitments
                cs2 = P.gen_circuit_for_test(20, 5)
                arr2 = U.arrays_from_oracle_circuit(co, pyref, cs2)
                pk2 = jf.PlonkKzgSnark.preprocess(ctx, key, arr2["selectors"], arr2["sigmas"], arr2["k"], arr2["wire_vars"], arr2["num_vars"],
                                                  arr2["pub_gate_ids"])
                opk2 = P.preprocess(cv, P.gen_srs(cv, beta, n + 2), cs2)
                ints2 = [rnd.randrange(fr.p) for _ in range(P.batch_num_blinders([cs, cs2]))]
                bl2 = co.ints_to_limbs([fr.to_mont(v) for v in ints2], 4)
                want = P.serialize_batch_proof(cv, P.batch_prove(cv, [cs, cs2], [opk, opk2], ints2, "solidity"))
                pk.shard_commits(comm, key_slice, a)
                pk2.shard_commits(comm, key_slice, a)
                got = jf.PlonkKzgSnark.batch_prove([pk, pk2], [arr["witness"], arr2["witness"]], bl2, "solidity")
                assert got.serialize_compressed() == want, "%s sharded batch proof differs" % transport
                pk2.free()
            # an unsatisfied witness fails on every rank alike (nobody is left waiting in an exchange)
            if name == "bench_2^10":
                pk.shard_commits(comm, key_slice, a)
                badw = arr["witness"].copy()
                badw[7] = arr["witness"][8]
                try:
                    prove(pk, badw, bl, "solidity")
                    raise AssertionError("expected WrongQuotientPolyDegree")
                except jf.WrongQuotientPolyDegree:
                    pass
                assert prove(pk, arr["witness"], bl, "solidity").serialize_compressed() == \
                    P.serialize_proof(cv, P.prove(cv, cs, opk, ints, "solidity"))
            pk.free()
            key_slice.free()
            key.free()
            checks += 1
        comm.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.write("rank %d ok (%d circuits x 2 transports)\n" % (rank, checks // 2))
    sys.stdout.flush()


def _tiny(adds):
    cs = P.PlonkCircuit()
    a = cs.create_public_variable(5)
    for _ in range(adds):
        a = cs.add(a, cs.one())
    cs.finalize_for_arithmetization()
    return cs


if __name__ == "__main__":
    main()
