"""ctypes binding of libjf_b200.so -- exactly the symbols include/jf_b200.h declares.

There is no fallback: if the shared library is missing or a symbol is absent, importing
anything that computes raises.  (Loading the library does not need a GPU; creating a
context does.)
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# JF_B200_LIB selects another build of the same library (A/B timing of a compile-time variant: csrc/Makefile VARIANT=...)
LIB_PATH = os.environ.get("JF_B200_LIB") or os.path.join(_HERE, "libjf_b200.so")

c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)

JF_OK = 0
JF_ERR_INVALID_ARG = -1
JF_ERR_CUDA = -2
JF_ERR_DOMAIN_TOO_LARGE = -3
JF_ERR_SCALAR_RANGE = -4
JF_ERR_NOMEM = -5
JF_ERR_QUOTIENT_DEGREE = -6
JF_ERR_COMM = -7
JF_COMM_ID_BYTES = 128

CURVES = {"bn254": 0, "bls12_381": 1}
FIELDS = {"bn254_fr": 0, "bn254_fq": 1, "bls12_381_fr": 2, "bls12_381_fq": 3}
FIELD_LIMBS = {"bn254_fr": 4, "bn254_fq": 4, "bls12_381_fr": 4, "bls12_381_fq": 6}
CURVE_FQ_LIMBS = {"bn254": 4, "bls12_381": 6}
CURVE_FR = {"bn254": "bn254_fr", "bls12_381": "bls12_381_fr"}
FIELD_OPS = {"mul": 0, "add": 1, "sub": 2, "sqr": 3, "inv": 4, "to_mont": 5, "from_mont": 6, "neg": 7}

c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_u8p = ctypes.POINTER(ctypes.c_uint8)


class PlonkProofStruct(ctypes.Structure):
    """`jf_plonk_proof` of include/jf_b200.h."""
    _fields_ = [
        ("curve", ctypes.c_int),
        ("wires_poly_comms", ctypes.c_uint64 * 60), ("wires_inf", ctypes.c_int * 5),
        ("prod_perm_poly_comm", ctypes.c_uint64 * 12), ("prod_perm_inf", ctypes.c_int),
        ("split_quot_poly_comms", ctypes.c_uint64 * 60), ("split_inf", ctypes.c_int * 5),
        ("opening_proof", ctypes.c_uint64 * 12), ("opening_inf", ctypes.c_int),
        ("shifted_opening_proof", ctypes.c_uint64 * 12), ("shifted_opening_inf", ctypes.c_int),
        ("wires_evals", ctypes.c_uint64 * 20), ("wire_sigma_evals", ctypes.c_uint64 * 16),
        ("perm_next_eval", ctypes.c_uint64 * 4), ("challenges", ctypes.c_uint64 * 20),
    ]


class UltraPlonkProofStruct(ctypes.Structure):
    """`jf_ultraplonk_proof` of include/jf_b200.h."""
    _fields_ = [
        ("curve", ctypes.c_int),
        ("wires_poly_comms", ctypes.c_uint64 * 72), ("wires_inf", ctypes.c_int * 6),
        ("prod_perm_poly_comm", ctypes.c_uint64 * 12), ("prod_perm_inf", ctypes.c_int),
        ("split_quot_poly_comms", ctypes.c_uint64 * 72), ("split_inf", ctypes.c_int * 6),
        ("opening_proof", ctypes.c_uint64 * 12), ("opening_inf", ctypes.c_int),
        ("shifted_opening_proof", ctypes.c_uint64 * 12), ("shifted_opening_inf", ctypes.c_int),
        ("wires_evals", ctypes.c_uint64 * 24), ("wire_sigma_evals", ctypes.c_uint64 * 20),
        ("perm_next_eval", ctypes.c_uint64 * 4),
        ("h_poly_comms", ctypes.c_uint64 * 24), ("h_inf", ctypes.c_int * 2),
        ("prod_lookup_poly_comm", ctypes.c_uint64 * 12), ("prod_lookup_inf", ctypes.c_int),
        ("plookup_evals", ctypes.c_uint64 * 60), ("challenges", ctypes.c_uint64 * 24),
    ]


class LinkProofStruct(ctypes.Structure):
    """`jf_link_proof` of include/jf_b200.h."""
    _fields_ = [
        ("curve", ctypes.c_int),
        ("quotient_commitment", ctypes.c_uint64 * 12), ("quotient_inf", ctypes.c_int),
        ("opening_proof", ctypes.c_uint64 * 12), ("opening_inf", ctypes.c_int),
        ("eta", ctypes.c_uint64 * 4), ("path", ctypes.c_int),
    ]


# name -> (restype, argtypes); must list every function of include/jf_b200.h
SIGNATURES = {
    "jf_ctx_create": (ctypes.c_int, [ctypes.c_int, c_void_pp]),
    "jf_ctx_destroy": (None, [ctypes.c_void_p]),
    "jf_ctx_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "jf_ctx_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "jf_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "jf_ctx_launch_count": (ctypes.c_uint64, [ctypes.c_void_p]),
    "jf_srs_load": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                   ctypes.c_long, ctypes.c_int, ctypes.c_int, c_void_pp]),
    "jf_srs_generate_for_testing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_u64p, ctypes.c_size_t, ctypes.c_size_t,
                                                   ctypes.c_int, ctypes.c_int, c_void_pp]),
    "jf_srs_lagrange": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_int, c_void_pp]),
    "jf_srs_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, c_u64p]),
    "jf_srs_len": (ctypes.c_size_t, [ctypes.c_void_p]),
    "jf_srs_window_bits": (ctypes.c_int, [ctypes.c_void_p]),
    "jf_srs_free": (None, [ctypes.c_void_p, ctypes.c_void_p]),
    "jf_msm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, c_u64p, ctypes.c_size_t, ctypes.c_int,
                              c_u64p, ctypes.POINTER(ctypes.c_int)]),
    "jf_msm_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(c_u64p),
                                    ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t), ctypes.c_size_t,
                                    ctypes.c_int, c_u64p, ctypes.POINTER(ctypes.c_int)]),
    "jf_kzg_open": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(c_u64p), ctypes.POINTER(ctypes.c_size_t),
                                   ctypes.c_size_t, c_u64p, c_u64p, ctypes.POINTER(ctypes.c_int), c_u64p]),
    "jf_msm_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_int, ctypes.c_void_p]),
    "jf_msm_combine": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_u64p, ctypes.c_size_t, c_u64p,
                                      ctypes.POINTER(ctypes.c_int)]),
    "jf_comm_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "jf_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, c_void_pp]),
    "jf_comm_destroy": (None, [ctypes.c_void_p]),
    "jf_comm_transport": (ctypes.c_int, [ctypes.c_void_p]),
    "jf_msm_sharded": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, c_u64p, ctypes.c_size_t,
                                      ctypes.c_int, c_u64p, ctypes.POINTER(ctypes.c_int)]),
    "jf_msm_sharded_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                             ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "jf_group_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int, c_void_pp]),
    "jf_group_destroy": (None, [ctypes.c_void_p]),
    "jf_group_size": (ctypes.c_int, [ctypes.c_void_p]),
    "jf_group_ctx": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    "jf_group_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "jf_group_srs_load": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                         ctypes.c_long, ctypes.c_int, ctypes.c_int, c_void_pp]),
    "jf_group_srs_generate_for_testing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_u64p, ctypes.c_size_t, ctypes.c_int,
                                                         ctypes.c_int, c_void_pp]),
    "jf_group_srs_free": (None, [ctypes.c_void_p, ctypes.c_void_p]),
    "jf_group_msm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, c_u64p, ctypes.c_size_t, ctypes.c_int,
                                    c_u64p, ctypes.POINTER(ctypes.c_int)]),
    "jf_group_ntt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_u64p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_int, c_u64p,
                                    ctypes.c_size_t, ctypes.c_size_t]),
    "jf_ntt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_u64p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_int, c_u64p,
                              ctypes.c_size_t, ctypes.c_size_t]),
    "jf_ntt_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint,
                                     ctypes.c_int, c_u64p, ctypes.c_size_t, ctypes.c_size_t]),
    "jf_ntt_cosets": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                     ctypes.c_uint, ctypes.c_int, c_u64p, ctypes.c_int, c_u64p]),
    "jf_dev_alloc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, c_void_pp]),
    "jf_dev_free": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "jf_dev_upload": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "jf_dev_download": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "jf_host_alloc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, c_void_pp]),
    "jf_host_free": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "jf_field_op": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_u64p, c_u64p, c_u64p, ctypes.c_size_t]),
    "jf_fixed_base_mul": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_u64p, ctypes.c_size_t, c_u64p]),
    "jf_profile_enable": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "jf_profile_collect": (ctypes.c_long, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]),
    "jf_microbench": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "jf_plonk_preprocess": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, c_u64p, c_u64p, c_u64p, c_u32p,
                                           ctypes.c_size_t, c_u32p, ctypes.c_size_t, ctypes.c_int, c_void_pp]),
    "jf_ultraplonk_preprocess": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, c_u64p, c_u64p, c_u64p, c_u32p,
                                                ctypes.c_size_t, c_u32p, ctypes.c_size_t, ctypes.c_uint, c_u64p, c_u64p, c_u64p,
                                                ctypes.c_int, c_void_pp]),
    "jf_ultraplonk_prove": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_u64p, c_u64p, ctypes.c_int, ctypes.c_char_p,
                                           ctypes.c_size_t, ctypes.POINTER(UltraPlonkProofStruct)]),
    "jf_ultraplonk_proof_serialize": (ctypes.c_long, [ctypes.POINTER(UltraPlonkProofStruct), ctypes.c_char_p, ctypes.c_size_t]),
    "jf_plonk_batch_prove": (ctypes.c_int, [ctypes.c_void_p, c_void_pp, ctypes.c_size_t, ctypes.POINTER(c_u64p), c_u64p, ctypes.c_int,
                                            ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(PlonkProofStruct)]),
    "jf_ultraplonk_batch_prove": (ctypes.c_int, [ctypes.c_void_p, c_void_pp, ctypes.c_size_t, ctypes.POINTER(c_u64p), c_u64p, ctypes.c_int,
                                                 ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(UltraPlonkProofStruct)]),
    "jf_plonk_batch_proof_serialize": (ctypes.c_long, [ctypes.POINTER(PlonkProofStruct), ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]),
    "jf_ultraplonk_batch_proof_serialize": (ctypes.c_long, [ctypes.POINTER(UltraPlonkProofStruct), ctypes.c_size_t, ctypes.c_char_p,
                                                            ctypes.c_size_t]),
    "jf_plonk_vk_commitments": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_u64p, ctypes.POINTER(ctypes.c_int)]),
    "jf_plonk_pk_free": (None, [ctypes.c_void_p, ctypes.c_void_p]),
    "jf_plonk_prove": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_u64p, c_u64p, ctypes.c_int, ctypes.c_char_p,
                                      ctypes.c_size_t, ctypes.POINTER(PlonkProofStruct)]),
    "jf_plonk_proof_serialize": (ctypes.c_long, [ctypes.POINTER(PlonkProofStruct), ctypes.c_char_p, ctypes.c_size_t]),
    "jf_plonk_pk_shard_commits": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                 ctypes.c_int]),
    "jf_plonk_link_hint": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_u64p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "jf_plonk_link_proofs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_u64p, ctypes.c_size_t, c_u64p, ctypes.c_int, c_u64p,
                                            ctypes.c_size_t, c_u64p, ctypes.c_int, ctypes.c_uint, ctypes.c_size_t, ctypes.c_size_t,
                                            ctypes.c_int, ctypes.c_int, ctypes.POINTER(LinkProofStruct)]),
    "jf_plonk_link_proofs_resident": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(PlonkProofStruct), ctypes.c_void_p,
                                                     ctypes.POINTER(PlonkProofStruct), ctypes.c_uint, ctypes.c_size_t, ctypes.c_size_t,
                                                     ctypes.c_int, ctypes.c_int, ctypes.POINTER(LinkProofStruct)]),
    "jf_poly_div_link_domain": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(c_u64p), ctypes.POINTER(ctypes.c_size_t),
                                               ctypes.c_size_t, ctypes.c_uint, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                               ctypes.POINTER(c_u64p)]),
    "jf_link_proof_serialize": (ctypes.c_long, [ctypes.POINTER(LinkProofStruct), ctypes.c_char_p, ctypes.c_size_t]),
    "jf_keccak256": (None, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]),
    "jf_transcript_new": (ctypes.c_void_p, [ctypes.c_int, ctypes.c_char_p]),
    "jf_transcript_free": (None, [ctypes.c_void_p]),
    "jf_transcript_append": (None, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]),
    "jf_transcript_challenge": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, c_u64p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load libjf_b200.so (built by `__graft_entry__.build()` / `make -C mpc-jellyfish_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libjf_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the symbol is missing: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib
