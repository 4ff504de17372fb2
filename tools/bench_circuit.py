"""Caller-side circuit data for the prover entry points (what `relation::PlonkCircuit` hands to
`preprocess` / `prove` after `finalize_for_arithmetization`), built with numpy so that 2^20-gate
circuits take seconds.  Not part of the product: the reference's `relation` crate is out of scope.

    arrays = bench_circuit_arrays(ctx, log_n)      # plonk/benches/bench.rs:29-46, num_gates = 2^log_n
    arrays = arrays_from_columns(ctx, ...)          # any circuit given as selector / wire-variable columns

Field conversions go through the library's own element-wise kernels (`ctx.field_op`, `ctx.ntt`).
"""
import numpy as np

BN254_FR_P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
# compute_coset_representatives::<ark_bn254::Fr>(5, _) (relation/src/constants.rs:30-79); public constants,
# pinned in tests/golden/transcript_vectors.json
BN254_K = [1,
           0x2f8dd1f1a7583c42c4e12a44e110404c73ca6c94813f85835da4fb7bb1301d4a,
           0x1ee678a0470a75a6eaa8fe837060498ba828a3703b311d0f77f010424afeb025,
           0x2042a587a90c187b0a087c03e29c968b950b1db26d5c82d666905a6895790c0a,
           0x2e2b91456103698adf57b799969dea1c8f739da5d8d40dd3eb9222db7c81e881]
# sixth representative of compute_coset_representatives::<ark_bn254::Fr>(6, _): the same ChaCha20 stream, one more draw
# (oracle/plonk_ref.py computes it; no published constant exists for it)
BN254_K5 = 0x1f20f5b0adb417179d42df7ddd4410a330afdb03e5c28949665b55adf7d7922d
NW, NSEL = 5, 13


def small_ints_to_limbs(vals: np.ndarray) -> np.ndarray:
    """non-negative int64 values -> (len, 4) canonical limbs"""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    out[:, 0] = vals.astype(np.uint64)
    return out


def int_to_limbs(v: int) -> np.ndarray:
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def wire_permutation(wire_vars: np.ndarray) -> np.ndarray:
    """compute_wire_permutation (constraint_system.rs:743-778) on flattened slots w * n + g: the slots of
    one variable, in ascending order, form a cycle.  Returns next_slot[slot]."""
    flat = wire_vars.reshape(-1).astype(np.int64)
    order = np.argsort(flat, kind="stable")
    sv = flat[order]
    first_of_group = np.concatenate(([True], sv[1:] != sv[:-1]))
    group_start = np.maximum.accumulate(np.where(first_of_group, np.arange(len(sv)), 0))
    last_of_group = np.concatenate((sv[1:] != sv[:-1], [True]))
    nxt_pos = np.where(last_of_group, group_start, np.arange(len(sv)) + 1)
    perm = np.empty(len(sv), dtype=np.int64)
    perm[order] = order[nxt_pos]
    return perm


def arrays_from_columns(ctx, field, log_n, selector_cols_small, wire_vars, witness_small, pub_gate_ids, k_ints, p):
    """selector_cols_small: (13 or 14, n) int64 with negative values meaning p - |v|; witness_small: int64 values;
    wire_vars: (5 or 6, n).  Everything is converted to Montgomery limbs with the library's kernels."""
    n = 1 << log_n
    NSEL, NW = selector_cols_small.shape[0], wire_vars.shape[0]
    to_mont = lambda a: ctx.field_op(field, "to_mont", np.ascontiguousarray(a))  # noqa: E731
    sel = np.zeros((NSEL, n, 4), dtype=np.uint64)
    for s in range(NSEL):
        col = selector_cols_small[s]
        pos = small_ints_to_limbs(np.abs(col))
        m = to_mont(pos)
        if (col < 0).any():
            neg = ctx.field_op(field, "neg", m)
            m[col < 0] = neg[col < 0]
        sel[s] = m
    # domain elements g^j = NTT of the polynomial X
    x_poly = np.zeros((n, 4), dtype=np.uint64)
    x_poly[1] = to_mont(int_to_limbs(1)[None, :])[0]
    elems = ctx.ntt(field, x_poly, log_n, False)
    k_m = to_mont(np.stack([int_to_limbs(k % p) for k in k_ints]))
    ext_id = np.zeros((NW * n, 4), dtype=np.uint64)
    for i in range(NW):
        ext_id[i * n:(i + 1) * n] = ctx.field_op(field, "mul", elems, np.repeat(k_m[i][None, :], n, axis=0))
    perm = wire_permutation(wire_vars)
    sigma = ext_id[perm].reshape(NW, n, 4)
    witness = to_mont(small_ints_to_limbs(witness_small))
    return {"selectors": sel, "sigmas": np.ascontiguousarray(sigma), "k": k_m, "wire_vars": wire_vars.astype(np.uint32),
            "witness": witness, "pub_gate_ids": list(pub_gate_ids), "num_vars": len(witness_small), "n": n, "log_n": log_n}


def bench_circuit_arrays(ctx, log_n, field="bn254_fr", k_ints=BN254_K, p=BN254_FR_P, ultra=False):
    """gen_circuit_for_bench(num_gates = 2^log_n, TurboPlonk | UltraPlonk): two constant gates (variables 0 and 1), then
    2^log_n - 10 additions a <- a + 1; padded with PaddingGate / variable 0.  UltraPlonk
    (`new_ultra_plonk(8)`, bench.rs:33-37): a sixth wire column (all variable 0: the circuit has no range gates), an all-zero
    q_lookup selector and all-zero table columns; the domain is max(gates, 2^8 + 1) rounded up = 2^log_n for log_n >= 9."""
    n = 1 << log_n
    adds = n - 10
    nw, nsel = (6, 14) if ultra else (NW, NSEL)
    if ultra:
        assert log_n >= 9 and k_ints is BN254_K
        k_ints = BN254_K + [BN254_K5]
    wire_vars = np.zeros((nw, n), dtype=np.int64)
    sel = np.zeros((nsel, n), dtype=np.int64)
    # gates 0, 1: ConstantGate on wires [0,0,0,0,var] with q_c = value, q_o = 1
    wire_vars[4, 0], wire_vars[4, 1] = 0, 1
    sel[10, 0] = sel[10, 1] = 1
    sel[11, 1] = 1
    t = np.arange(adds)
    g = 2 + t
    wire_vars[0, g] = np.where(t == 0, 0, 1 + t)
    wire_vars[1, g] = 1
    wire_vars[4, g] = 2 + t
    sel[0, g] = 1
    sel[1, g] = 1
    sel[10, g] = 1
    witness = np.concatenate(([0, 1], t + 1)).astype(np.int64)
    arr = arrays_from_columns(ctx, field, log_n, sel, wire_vars, witness, [], k_ints, p)
    if ultra:
        zero = np.zeros((n, 4), dtype=np.uint64)
        arr.update({"range_bit_len": 8, "table_key": zero, "table_dom_sep": zero, "q_dom_sep": zero})
    return arr
