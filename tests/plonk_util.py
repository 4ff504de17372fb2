"""Shared helpers of the prover tests: oracle circuit -> product arrays, product proof -> oracle dict."""
import numpy as np


def limbs(co, vals):
    return co.ints_to_limbs(list(vals), 4)


def arrays_from_oracle_circuit(co, py, cs):
    """What `relation::PlonkCircuit` hands to preprocess / prove, as Montgomery limb arrays."""
    f = cs.f
    n = cs.n
    mont = lambda vs: limbs(co, [f.to_mont(v) for v in vs])  # noqa: E731
    sel = np.stack([mont(col) for col in cs.selector_evals()])
    ext = cs.extended_permutation()
    nw = getattr(cs, "nw", 5)
    sig = np.stack([mont(ext[i * n:(i + 1) * n]) for i in range(nw)])
    out = {"selectors": sel, "sigmas": sig, "k": mont(cs.k), "wire_vars": np.array(cs.wire_variables[:nw], dtype=np.uint32),
           "witness": mont(cs.witness), "pub_gate_ids": list(cs.pub_input_gate_ids), "num_vars": cs.num_vars(), "n": n,
           "log_n": n.bit_length() - 1}
    if getattr(cs, "ultra", False):   # the three per-gate Plookup columns (constraint_system.rs:873-888)
        out.update({"range_bit_len": cs.range_bit_len, "table_key": mont(cs.table_key_vec()),
                    "table_dom_sep": mont(cs.table_dom_sep_vec()), "q_dom_sep": mont(cs.q_dom_sep())})
    return out


def point_to_affine(co, cv, xy, inf):
    if inf:
        return None
    L = cv.fq.limbs64
    x, y = co.limbs_to_ints(np.asarray(xy, dtype=np.uint64).reshape(2, L))
    return (cv.fq.from_mont(x), cv.fq.from_mont(y))


def proof_to_oracle(co, cv, pr):
    fr = cv.fr
    ev = lambda a: [fr.from_mont(v) for v in co.limbs_to_ints(np.asarray(a).reshape(-1, 4))]  # noqa: E731
    nw = len(pr.wires_poly_comms)
    out = {
        "wires_poly_comms": [point_to_affine(co, cv, pr.wires_poly_comms[i], pr.wires_inf[i]) for i in range(nw)],
        "prod_perm_poly_comm": point_to_affine(co, cv, pr.prod_perm_poly_comm, pr.prod_perm_inf),
        "split_quot_poly_comms": [point_to_affine(co, cv, pr.split_quot_poly_comms[i], pr.split_inf[i]) for i in range(nw)],
        "opening_proof": point_to_affine(co, cv, pr.opening_proof, pr.opening_inf),
        "shifted_opening_proof": point_to_affine(co, cv, pr.shifted_opening_proof, pr.shifted_opening_inf),
        "wires_evals": ev(pr.wires_evals), "wire_sigma_evals": ev(pr.wire_sigma_evals),
        "perm_next_eval": ev(pr.perm_next_eval)[0], "plookup_proof": None,
    }
    if pr.h_poly_comms is not None:
        from mpc_jellyfish_b200.plonk import PLOOKUP_EVAL_FIELDS
        out["plookup_proof"] = {
            "h_poly_comms": [point_to_affine(co, cv, pr.h_poly_comms[i], pr.h_inf[i]) for i in range(2)],
            "prod_lookup_poly_comm": point_to_affine(co, cv, pr.prod_lookup_poly_comm, pr.prod_lookup_inf),
            "poly_evals": dict(zip(PLOOKUP_EVAL_FIELDS, ev(pr.plookup_evals)))}
    return out


def vk_from_product(co, cv, pk, k_ints):
    vk = {"domain_size": pk.n, "num_inputs": pk.num_inputs,
          "selector_comms": [point_to_affine(co, cv, pk.selector_comms[i], pk.selector_inf[i]) for i in range(len(pk.selector_comms))],
          "sigma_comms": [point_to_affine(co, cv, pk.sigma_comms[i], pk.sigma_inf[i]) for i in range(len(pk.sigma_comms))],
          "k": list(k_ints), "plookup": None}
    if getattr(pk, "ultra", False):
        names = ["range_table_comm", "key_table_comm", "table_dom_sep_comm", "q_dom_sep_comm"]
        vk["plookup"] = {nm: point_to_affine(co, cv, pk.plookup_comms[i], pk.plookup_inf[i]) for i, nm in enumerate(names)}
    return vk
