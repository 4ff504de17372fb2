//! Raw bindings of `include/jf_b200.h`.  One declaration per exported symbol; see the header for the
//! contract of each call and the reference call site it replaces.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_long, c_uint, c_void};

#[repr(C)] pub struct jf_ctx { _p: [u8; 0] }
#[repr(C)] pub struct jf_srs { _p: [u8; 0] }
#[repr(C)] pub struct jf_plonk_pk { _p: [u8; 0] }
#[repr(C)] pub struct jf_comm { _p: [u8; 0] }
#[repr(C)] pub struct jf_group { _p: [u8; 0] }
#[repr(C)] pub struct jf_group_srs { _p: [u8; 0] }

pub const JF_OK: c_int = 0;
pub const JF_ERR_INVALID_ARG: c_int = -1;
pub const JF_ERR_CUDA: c_int = -2;
pub const JF_ERR_DOMAIN_TOO_LARGE: c_int = -3;
pub const JF_ERR_SCALAR_RANGE: c_int = -4;
pub const JF_ERR_NOMEM: c_int = -5;
pub const JF_ERR_QUOTIENT_DEGREE: c_int = -6;
pub const JF_ERR_COMM: c_int = -7;
pub const JF_COMM_ID_BYTES: usize = 128;
pub const JF_BN254: c_int = 0;
pub const JF_BLS12_381: c_int = 1;
pub const JF_BN254_FR: c_int = 0;
pub const JF_BLS12_381_FR: c_int = 2;

/// `jf_plonk_proof`: points are x || y with 2 L limbs each, packed (L = 4 BN254, 6 BLS12-381).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct jf_plonk_proof {
    pub curve: c_int,
    pub wires_poly_comms: [u64; 60], pub wires_inf: [c_int; 5],
    pub prod_perm_poly_comm: [u64; 12], pub prod_perm_inf: c_int,
    pub split_quot_poly_comms: [u64; 60], pub split_inf: [c_int; 5],
    pub opening_proof: [u64; 12], pub opening_inf: c_int,
    pub shifted_opening_proof: [u64; 12], pub shifted_opening_inf: c_int,
    pub wires_evals: [u64; 20], pub wire_sigma_evals: [u64; 16], pub perm_next_eval: [u64; 4],
    pub challenges: [u64; 20],
}

/// `jf_link_proof`: `LinkingProof<E>` (plonk/src/proof_system/proof_linking.rs:32-39) plus diagnostics.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct jf_link_proof {
    pub curve: c_int,
    pub quotient_commitment: [u64; 12], pub quotient_inf: c_int,
    pub opening_proof: [u64; 12], pub opening_inf: c_int,
    pub eta: [u64; 4],
    pub path: c_int,
}

/// `jf_ultraplonk_proof`: an UltraPlonk `Proof<E>` with `plookup_proof: Some(..)`.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct jf_ultraplonk_proof {
    pub curve: c_int,
    pub wires_poly_comms: [u64; 72], pub wires_inf: [c_int; 6],
    pub prod_perm_poly_comm: [u64; 12], pub prod_perm_inf: c_int,
    pub split_quot_poly_comms: [u64; 72], pub split_inf: [c_int; 6],
    pub opening_proof: [u64; 12], pub opening_inf: c_int,
    pub shifted_opening_proof: [u64; 12], pub shifted_opening_inf: c_int,
    pub wires_evals: [u64; 24], pub wire_sigma_evals: [u64; 20], pub perm_next_eval: [u64; 4],
    pub h_poly_comms: [u64; 24], pub h_inf: [c_int; 2],
    pub prod_lookup_poly_comm: [u64; 12], pub prod_lookup_inf: c_int,
    pub plookup_evals: [u64; 60],
    pub challenges: [u64; 24],
}

extern "C" {
    pub fn jf_ctx_create(device: c_int, out: *mut *mut jf_ctx) -> c_int;
    pub fn jf_ctx_destroy(ctx: *mut jf_ctx);
    pub fn jf_ctx_set_stream(ctx: *mut jf_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn jf_ctx_sync(ctx: *mut jf_ctx) -> c_int;
    pub fn jf_last_error(ctx: *const jf_ctx) -> *const c_char;
    pub fn jf_ctx_launch_count(ctx: *const jf_ctx) -> u64;

    pub fn jf_srs_load(ctx: *mut jf_ctx, curve: c_int, affine_pts: *const c_void, n: usize, stride_bytes: usize,
                       inf_flag_offset: c_long, window_bits: c_int, precompute: c_int, out: *mut *mut jf_srs) -> c_int;
    pub fn jf_srs_generate_for_testing(ctx: *mut jf_ctx, curve: c_int, beta: *const u64, first_power: usize, n: usize,
                                       window_bits: c_int, precompute: c_int, out: *mut *mut jf_srs) -> c_int;
    pub fn jf_srs_lagrange(ctx: *mut jf_ctx, srs: *const jf_srs, log_n: c_uint, mask_points: c_int, out: *mut *mut jf_srs) -> c_int;
    pub fn jf_srs_read(ctx: *mut jf_ctx, srs: *const jf_srs, first: usize, count: usize, out_xy: *mut u64) -> c_int;
    pub fn jf_srs_len(srs: *const jf_srs) -> usize;
    pub fn jf_srs_window_bits(srs: *const jf_srs) -> c_int;
    pub fn jf_srs_free(ctx: *mut jf_ctx, srs: *mut jf_srs);

    pub fn jf_msm(ctx: *mut jf_ctx, srs: *const jf_srs, base_offset: usize, scalars: *const u64, n: usize,
                  scalars_in_montgomery: c_int, out_xy: *mut u64, out_infinity: *mut c_int) -> c_int;
    pub fn jf_msm_batch(ctx: *mut jf_ctx, srs: *const jf_srs, scalars: *const *const u64, lens: *const usize,
                        base_offsets: *const usize, batch: usize, scalars_in_montgomery: c_int, out_xy: *mut u64,
                        out_infinity: *mut c_int) -> c_int;
    pub fn jf_kzg_open(ctx: *mut jf_ctx, srs: *const jf_srs, polys: *const *const u64, lens: *const usize, batch: usize,
                       points: *const u64, out_proof_xy: *mut u64, out_infinity: *mut c_int, out_evals: *mut u64) -> c_int;
    pub fn jf_msm_device(ctx: *mut jf_ctx, srs: *const jf_srs, base_offset: usize, d_scalars: *const c_void, n: usize,
                         scalars_in_montgomery: c_int, d_out_xyzz: *mut c_void) -> c_int;
    pub fn jf_msm_combine(ctx: *mut jf_ctx, curve: c_int, xyzz_parts: *const u64, parts: usize, out_xy: *mut u64,
                          out_infinity: *mut c_int) -> c_int;

    // multi-GPU, one process per GPU
    pub fn jf_comm_unique_id(id: *mut u8) -> c_int;
    pub fn jf_comm_init(ctx: *mut jf_ctx, rank: c_int, nranks: c_int, id: *const u8, transport: c_int, out: *mut *mut jf_comm) -> c_int;
    pub fn jf_comm_destroy(comm: *mut jf_comm);
    pub fn jf_comm_transport(comm: *const jf_comm) -> c_int;
    pub fn jf_msm_sharded(ctx: *mut jf_ctx, comm: *mut jf_comm, key_slice: *const jf_srs, base_offset: usize, scalars: *const u64,
                          n_local: usize, scalars_in_montgomery: c_int, out_xy: *mut u64, out_infinity: *mut c_int) -> c_int;
    pub fn jf_msm_sharded_device(ctx: *mut jf_ctx, comm: *mut jf_comm, key_slice: *const jf_srs, base_offset: usize,
                                 d_scalars: *const c_void, n_local: usize, scalars_in_montgomery: c_int,
                                 d_out_parts: *mut c_void) -> c_int;
    // multi-GPU, one process driving several GPUs
    pub fn jf_group_create(devices: *const c_int, n_dev: c_int, out: *mut *mut jf_group) -> c_int;
    pub fn jf_group_destroy(g: *mut jf_group);
    pub fn jf_group_size(g: *const jf_group) -> c_int;
    pub fn jf_group_ctx(g: *mut jf_group, i: c_int) -> *mut jf_ctx;
    pub fn jf_group_last_error(g: *const jf_group) -> *const c_char;
    pub fn jf_group_srs_load(g: *mut jf_group, curve: c_int, affine_pts: *const c_void, n: usize, stride_bytes: usize,
                             inf_flag_offset: c_long, window_bits: c_int, precompute: c_int, out: *mut *mut jf_group_srs) -> c_int;
    pub fn jf_group_srs_generate_for_testing(g: *mut jf_group, curve: c_int, beta: *const u64, n: usize, window_bits: c_int,
                                             precompute: c_int, out: *mut *mut jf_group_srs) -> c_int;
    pub fn jf_group_srs_free(g: *mut jf_group, srs: *mut jf_group_srs);
    pub fn jf_group_msm(g: *mut jf_group, srs: *const jf_group_srs, base_offset: usize, scalars: *const u64, n: usize,
                        scalars_in_montgomery: c_int, out_xy: *mut u64, out_infinity: *mut c_int) -> c_int;
    pub fn jf_group_ntt(g: *mut jf_group, field: c_int, data: *mut u64, in_len: usize, log_n: c_uint, inverse: c_int,
                        coset_offset: *const u64, batch: usize, batch_stride: usize) -> c_int;

    pub fn jf_ntt(ctx: *mut jf_ctx, field: c_int, data: *mut u64, in_len: usize, log_n: c_uint, inverse: c_int,
                  coset_offset: *const u64, batch: usize, batch_stride: usize) -> c_int;
    pub fn jf_ntt_device(ctx: *mut jf_ctx, field: c_int, d_data: *mut c_void, in_len: usize, log_n: c_uint, inverse: c_int,
                         coset_offset: *const u64, batch: usize, batch_stride: usize) -> c_int;
    pub fn jf_ntt_cosets(ctx: *mut jf_ctx, field: c_int, polys_in: *const u64, in_len: usize, in_stride: usize, polys: usize,
                         log_n: c_uint, inverse: c_int, offsets: *const u64, rows: c_int, out: *mut u64) -> c_int;

    pub fn jf_plonk_preprocess(ctx: *mut jf_ctx, srs: *const jf_srs, log_n: c_uint, selector_evals: *const u64,
                               sigma_evals: *const u64, k: *const u64, wire_variables: *const u32, num_vars: usize,
                               pub_input_gate_ids: *const u32, num_inputs: usize, flags: c_int,
                               out: *mut *mut jf_plonk_pk) -> c_int;
    pub fn jf_ultraplonk_preprocess(ctx: *mut jf_ctx, srs: *const jf_srs, log_n: c_uint, selector_evals: *const u64,
                                    sigma_evals: *const u64, k: *const u64, wire_variables: *const u32, num_vars: usize,
                                    pub_input_gate_ids: *const u32, num_inputs: usize, range_bit_len: c_uint,
                                    table_key_evals: *const u64, table_dom_sep_evals: *const u64, q_dom_sep_evals: *const u64,
                                    flags: c_int, out: *mut *mut jf_plonk_pk) -> c_int;
    pub fn jf_ultraplonk_prove(ctx: *mut jf_ctx, pk: *mut jf_plonk_pk, witness: *const u64, blinders: *const u64,
                               transcript_kind: c_int, extra_msg: *const u8, extra_len: usize,
                               out: *mut jf_ultraplonk_proof) -> c_int;
    pub fn jf_ultraplonk_proof_serialize(proof: *const jf_ultraplonk_proof, out: *mut u8, cap: usize) -> c_long;
    pub fn jf_plonk_batch_prove(ctx: *mut jf_ctx, pks: *const *mut jf_plonk_pk, count: usize, witnesses: *const *const u64,
                                blinders: *const u64, transcript_kind: c_int, extra_msg: *const u8, extra_len: usize,
                                out: *mut jf_plonk_proof) -> c_int;
    pub fn jf_ultraplonk_batch_prove(ctx: *mut jf_ctx, pks: *const *mut jf_plonk_pk, count: usize, witnesses: *const *const u64,
                                     blinders: *const u64, transcript_kind: c_int, extra_msg: *const u8, extra_len: usize,
                                     out: *mut jf_ultraplonk_proof) -> c_int;
    pub fn jf_plonk_batch_proof_serialize(proofs: *const jf_plonk_proof, count: usize, out: *mut u8, cap: usize) -> c_long;
    pub fn jf_ultraplonk_batch_proof_serialize(proofs: *const jf_ultraplonk_proof, count: usize, out: *mut u8, cap: usize) -> c_long;
    pub fn jf_plonk_vk_commitments(ctx: *mut jf_ctx, pk: *const jf_plonk_pk, out_xy: *mut u64, out_inf: *mut c_int) -> c_int;
    pub fn jf_plonk_pk_free(ctx: *mut jf_ctx, pk: *mut jf_plonk_pk);
    pub fn jf_plonk_prove(ctx: *mut jf_ctx, pk: *mut jf_plonk_pk, witness: *const u64, blinders: *const u64,
                          transcript_kind: c_int, extra_msg: *const u8, extra_len: usize, out: *mut jf_plonk_proof) -> c_int;
    pub fn jf_plonk_proof_serialize(proof: *const jf_plonk_proof, out: *mut u8, cap: usize) -> c_long;

    pub fn jf_plonk_pk_shard_commits(ctx: *mut jf_ctx, pk: *mut jf_plonk_pk, comm: *mut jf_comm, key_slice: *const jf_srs,
                                     slice_start: usize, shard_round3: c_int) -> c_int;
    pub fn jf_plonk_link_hint(ctx: *mut jf_ctx, pk: *const jf_plonk_pk, out_poly: *mut u64, cap: usize, out_len: *mut usize) -> c_int;
    pub fn jf_plonk_link_proofs(ctx: *mut jf_ctx, srs: *const jf_srs, a1: *const u64, len1: usize, a1_comm_xy: *const u64, a1_inf: c_int,
                                a2: *const u64, len2: usize, a2_comm_xy: *const u64, a2_inf: c_int, alignment: c_uint, offset: usize,
                                size: usize, transcript_kind: c_int, flags: c_int, out: *mut jf_link_proof) -> c_int;
    pub fn jf_plonk_link_proofs_resident(ctx: *mut jf_ctx, lhs: *const jf_plonk_pk, lhs_proof: *const jf_plonk_proof,
                                         rhs: *const jf_plonk_pk, rhs_proof: *const jf_plonk_proof, alignment: c_uint, offset: usize,
                                         size: usize, transcript_kind: c_int, flags: c_int, out: *mut jf_link_proof) -> c_int;
    pub fn jf_poly_div_link_domain(ctx: *mut jf_ctx, field: c_int, polys: *const *const u64, lens: *const usize, batch: usize,
                                   alignment: c_uint, offset: usize, size: usize, flags: c_int, out_quotients: *const *mut u64) -> c_int;
    pub fn jf_link_proof_serialize(proof: *const jf_link_proof, out: *mut u8, cap: usize) -> c_long;

    pub fn jf_keccak256(data: *const u8, len: usize, out: *mut u8);
    pub fn jf_transcript_new(kind: c_int, label: *const c_char) -> *mut c_void;
    pub fn jf_transcript_free(t: *mut c_void);
    pub fn jf_transcript_append(t: *mut c_void, label: *const c_char, msg: *const u8, len: usize);
    pub fn jf_transcript_challenge(t: *mut c_void, field: c_int, label: *const c_char, out_montgomery: *mut u64) -> c_int;

    pub fn jf_dev_alloc(ctx: *mut jf_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn jf_dev_free(ctx: *mut jf_ctx, ptr: *mut c_void) -> c_int;
    pub fn jf_dev_upload(ctx: *mut jf_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn jf_dev_download(ctx: *mut jf_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn jf_host_alloc(ctx: *mut jf_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn jf_host_free(ctx: *mut jf_ctx, ptr: *mut c_void) -> c_int;
    pub fn jf_field_op(ctx: *mut jf_ctx, field: c_int, op: c_int, a: *const u64, b: *const u64, out: *mut u64, n: usize) -> c_int;
    pub fn jf_fixed_base_mul(ctx: *mut jf_ctx, curve: c_int, scalars: *const u64, n: usize, out_xy: *mut u64) -> c_int;
    pub fn jf_profile_enable(ctx: *mut jf_ctx, on: c_int) -> c_int;
    pub fn jf_profile_collect(ctx: *mut jf_ctx, buf: *mut c_char, cap: usize) -> c_long;
    pub fn jf_microbench(ctx: *mut jf_ctx, kind: c_int, out_rate: *mut f64) -> c_int;
}
