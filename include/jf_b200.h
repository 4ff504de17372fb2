/* jf_b200.h -- C ABI of the B200-native MSM / NTT core for the mpc-jellyfish PLONK prover.
 *
 * The reference (100 % Rust) has no FFI seam of its own: its hot path is the set of static
 * trait calls out of the workspace into arkworks.  Each entry point below
 * names the call(s) it replaces.  A `jf-b200-sys` crate binds exactly these symbols
 * (see INTEGRATION.md for the bindgen/`extern "C"` block and the patched call sites).
 *
 * Conventions
 *   - Field elements: ark-ff's in-memory layout, little-endian u64 limbs (4 for BN254 Fr/Fq and
 *     BLS12-381 Fr, 6 for BLS12-381 Fq), MONTGOMERY form unless a parameter says otherwise.
 *     A `&[Fr]` / `&mut [Fr]` can be passed as is.
 *   - Affine G1 point: x || y (2*L limbs).  The identity is (0, 0) plus, on output, an int
 *     flag (ark-ec's `Affine { x: 0, y: 0, infinity: true }`).
 *   - Every function returns a jf_status; nothing throws or aborts across the boundary.
 *     jf_last_error() gives a human-readable reason for the last failure on that context.
 *   - Contexts are thread-safe (rayon / tokio workers may call concurrently: calls on one
 *     context are serialised on its stream).  Host buffers are caller-owned and only read
 *     (scalars, points) or overwritten in place (jf_ntt) during the call; device memory is
 *     owned by the library behind opaque handles.
 *   - There is NO CPU fallback: without a usable CUDA device jf_ctx_create fails.
 */
#ifndef JF_B200_H
#define JF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jf_ctx jf_ctx; /* one CUDA device + stream + workspace */
typedef struct jf_srs jf_srs; /* device-resident commit key (`UnivariateProverParam::powers_of_g`) */
typedef struct jf_plonk_pk jf_plonk_pk; /* device-resident `ProvingKey` + per-proof workspace */
typedef struct jf_comm jf_comm;   /* this rank's end of a one-process-per-GPU group (range-sharded MSM) */
typedef struct jf_group jf_group; /* several GPUs driven by ONE process: contexts + peer access */
typedef struct jf_group_srs jf_group_srs; /* a commit key split by point range over the GPUs of a jf_group */

typedef enum {
    JF_OK = 0,
    JF_ERR_INVALID_ARG = -1,      /* -> PCSError::InvalidParameters (primitives/src/pcs/errors.rs:17-34) */
    JF_ERR_CUDA = -2,             /* -> PCSError::UpstreamError */
    JF_ERR_DOMAIN_TOO_LARGE = -3, /* log_n > two-adicity -> PlonkError::DomainCreationError (plonk/src/errors.rs:16-49) */
    JF_ERR_SCALAR_RANGE = -4,     /* a scalar was >= the group order (arkworks BigInts from into_bigint never are) */
    JF_ERR_NOMEM = -5,
    JF_ERR_QUOTIENT_DEGREE = -6,  /* -> SnarkError::WrongQuotientPolyDegree (prover.rs:916-919): witness does not satisfy the circuit */
    JF_ERR_COMM = -7              /* NCCL unavailable / failed, or a peer did not deliver its partial sum in time -> PCSError::UpstreamError */
} jf_status;

typedef enum { JF_BN254 = 0, JF_BLS12_381 = 1 } jf_curve;
typedef enum { JF_BN254_FR = 0, JF_BN254_FQ = 1, JF_BLS12_381_FR = 2, JF_BLS12_381_FQ = 3 } jf_field;

/* ---- context ------------------------------------------------------------------------ */
int jf_ctx_create(int device, jf_ctx **out);
void jf_ctx_destroy(jf_ctx *ctx);
/* Run on an externally owned CUDA stream (a `cudaStream_t` passed as void*); NULL = the context's own. */
int jf_ctx_set_stream(jf_ctx *ctx, void *cuda_stream);
int jf_ctx_sync(jf_ctx *ctx);
const char *jf_last_error(const jf_ctx *ctx);
/* Number of kernels this context has launched since creation (bench.py's `gpu_launches`). */
uint64_t jf_ctx_launch_count(const jf_ctx *ctx);

/* ---- commit key ---------------------------------------------------------------------
 * jf_srs_load uploads `powers_of_g` once (it is immutable across proofs:
 * plonk/src/proof_system/structs.rs:583 `ProvingKey.commit_key`).  Records are `stride_bytes`
 * apart, each starting with x || y in Montgomery form; if `inf_flag_offset` >= 0 the byte at
 * that offset inside a record is ark-ec's `infinity: bool`, otherwise (0,0) means identity.
 * `window_bits` = 0 picks the signed-digit window from n; `precompute` != 0 additionally
 * stores 2^(window_bits*t) * P_i for every window t so that all windows share one bucket set.
 */
int jf_srs_load(jf_ctx *ctx, int curve, const void *affine_pts, size_t n, size_t stride_bytes,
                long inf_flag_offset, int window_bits, int precompute, jf_srs **out);
/* `gen_srs_for_testing` (primitives/src/pcs/univariate_kzg/srs.rs:118-153) with a caller-chosen,
 * KNOWN beta (4 canonical limbs) and g = the curve generator: key[i] = beta^(first_power + i) * g
 * (first_power > 0 yields the slice of the key one GPU holds in a range-sharded MSM). */
int jf_srs_generate_for_testing(jf_ctx *ctx, int curve, const uint64_t *beta, size_t first_power, size_t n,
                                int window_bits, int precompute, jf_srs **out);
/* The commit key in the LAGRANGE basis of the size-2^log_n domain: out[j] = [L_j(beta)] G, so that
 * `commit(p) = sum_j p(w^j) * out[j]` is an MSM over a polynomial's VALUES on the domain (zero and small values then cost nothing or
 * little).  Computed from `srs` alone -- beta stays unknown -- as the inverse DFT of its first 2^log_n points taken in the group
 * (n/2 log n scalar multiplications: about a second at 2^20, once per key and domain).  mask_points != 0 appends P_n - P_0 and
 * P_(n+1) - P_1, the commit-key side of the prover's masking terms (b_0 + b_1 X)(X^n - 1) (prover.rs:463-486). */
int jf_srs_lagrange(jf_ctx *ctx, const jf_srs *srs, unsigned log_n, int mask_points, jf_srs **out);
/* Copy `count` affine points starting at `first` back to the host (x || y each). */
int jf_srs_read(jf_ctx *ctx, const jf_srs *srs, size_t first, size_t count, uint64_t *out_xy);
size_t jf_srs_len(const jf_srs *srs);
int jf_srs_window_bits(const jf_srs *srs);
void jf_srs_free(jf_ctx *ctx, jf_srs *srs);

/* ---- MSM ----------------------------------------------------------------------------
 * jf_msm == `E::G1::msm_bigint(&powers_of_g[base_offset..], scalars).into_affine()`
 * (primitives/src/pcs/univariate_kzg/mod.rs:108-111 in `commit`, :151-155 in `open`).
 * Uses min(n, srs_len - base_offset) pairs like arkworks.  Scalars: 4 limbs each; canonical
 * BigInts when scalars_in_montgomery == 0 (what `into_bigint`, mod.rs:392, produces), or the
 * polynomial's own Montgomery-form coefficients when != 0 (the conversion then happens on the
 * GPU and `convert_to_bigints`, mod.rs:390-395, can be dropped by the caller).
 * Result: out_xy = affine x || y (Montgomery), *out_infinity = 1 for the identity. */
int jf_msm(jf_ctx *ctx, const jf_srs *srs, size_t base_offset, const uint64_t *scalars, size_t n,
           int scalars_in_montgomery, uint64_t *out_xy, int *out_infinity);
/* jf_msm_batch == `batch_commit`'s par_iter over polynomials (mod.rs:119-131) as ONE call: the vectors are uploaded on a
 * copy stream while their predecessors are being sorted and accumulated, and groups of up to 8 MSMs share one bucket
 * reduction (its ~20 dependent levels cost the same latency for one bucket set or many).  Empty vectors give the identity. */
int jf_msm_batch(jf_ctx *ctx, const jf_srs *srs, const uint64_t *const *scalars, const size_t *lens,
                 const size_t *base_offsets, size_t batch, int scalars_in_montgomery, uint64_t *out_xy,
                 int *out_infinity);
/* jf_kzg_open == `UnivariateKzgPCS::open` (mod.rs:135-161) for `batch` (polynomial, point) pairs in one call:
 * the witness polynomial p / (X - z) (remainder dropped, as ark-poly's `/`), its commitment and p(z) are
 * computed on the GPU (division = power scaling + one suffix scan; evaluation = blocked Horner).
 * polys[i]: lens[i] Montgomery coefficients, low degree first; points: batch x 4 Montgomery limbs.
 * Outputs: affine proofs (x || y each), infinity flags, evaluations (batch x 4 Montgomery limbs).  With the
 * share and the MAC polynomial of an authenticated share as a batch of two this is also the local part of
 * `MultiproverKZG::open` (plonk/src/multiprover/primitives/multiprover_kzg.rs:171-197). */
int jf_kzg_open(jf_ctx *ctx, const jf_srs *srs, const uint64_t *const *polys, const size_t *lens, size_t batch,
                const uint64_t *points, uint64_t *out_proof_xy, int *out_infinity, uint64_t *out_evals);
/* Device-resident form: scalars already in HBM (device pointer), result left in HBM as one
 * XYZZ point (4*L limbs: X, Y, ZZ, ZZZ; x = X/ZZ, y = Y/ZZZ; ZZ = 0 <=> identity).  This is the
 * per-GPU partial sum of a range-sharded MSM; it is all-gathered and fed to jf_msm_combine. */
int jf_msm_device(jf_ctx *ctx, const jf_srs *srs, size_t base_offset, const void *d_scalars, size_t n,
                  int scalars_in_montgomery, void *d_out_xyzz);
/* Sum `parts` XYZZ partial results (host memory) and normalise: the tail of a sharded MSM.
 * Pure host code; ctx may be NULL. */
int jf_msm_combine(jf_ctx *ctx, int curve, const uint64_t *xyzz_parts, size_t parts, uint64_t *out_xy,
                   int *out_infinity);

/* ---- multi-GPU MSM, sharded by point range ------------------------------------
 * `commit` is a sum over (coefficient, key point) pairs (mod.rs:106-111), so GPU g can hold key[start_g, end_g) resident and
 * receive the matching scalar slice; the partial sums (one XYZZ point = 128 B BN254 / 192 B BLS12-381 per GPU) are joined by
 * ONE exchange step.  Two forms:
 *
 * (1) one process per GPU (torchrun / MPI style; what bench.py --gpus N runs).  Rank 0 calls jf_comm_unique_id and hands the
 *     128 bytes to the other ranks by whatever channel the caller has; every rank then calls jf_comm_init (collective).
 *     NCCL (libnccl.so.2, loaded at run time: single-GPU users need not have it) bootstraps the group and is one of the two
 *     transports.  transport: 1 = `ncclAllGather` of the partials; 2 = peer-memory mailboxes: every rank maps the other
 *     ranks' mailbox through CUDA IPC (ranks that are threads of ONE process: through peer access), and ONE kernel at the tail of the MSM stores its partial into all peers' HBM over
 *     NVLink and waits for theirs (no collective launch; the wait gives up after 2 s with JF_ERR_COMM); 0 = 2 when all peers
 *     are reachable, else 1 (env JF_COMM_TRANSPORT=nccl|p2p overrides).
 *     jf_msm_sharded(_device) are collective: every rank calls them in the same order with its own key slice and scalar
 *     slice and obtains the same result.  The final G-1 additions and the `into_affine` inversion run on the host (a lone
 *     GPU thread needs ~0.2 ms for a Fermat inversion, a CPU core ~20 us). */
#define JF_COMM_ID_BYTES 128
int jf_comm_unique_id(uint8_t id[JF_COMM_ID_BYTES]);
int jf_comm_init(jf_ctx *ctx, int rank, int nranks, const uint8_t id[JF_COMM_ID_BYTES], int transport, jf_comm **out);
void jf_comm_destroy(jf_comm *comm);
int jf_comm_transport(const jf_comm *comm); /* 1 nccl, 2 p2p */
/* scalars: this rank's slice in host memory; out_xy / out_infinity as jf_msm */
int jf_msm_sharded(jf_ctx *ctx, jf_comm *comm, const jf_srs *key_slice, size_t base_offset, const uint64_t *scalars,
                   size_t n_local, int scalars_in_montgomery, uint64_t *out_xy, int *out_infinity);
/* device form: scalars in HBM; leaves all `nranks` XYZZ partials (rank order, 4*L limbs each) in d_out_parts on every rank;
 * asynchronous on the context's stream.  jf_msm_combine finishes on the host. */
int jf_msm_sharded_device(jf_ctx *ctx, jf_comm *comm, const jf_srs *key_slice, size_t base_offset, const void *d_scalars,
                          size_t n_local, int scalars_in_montgomery, void *d_out_parts);
/* (2) ONE process driving several GPUs -- the drop-in for a Rust prover process (rayon threads, one address space).  The
 *     group owns one context per device and enables peer access; the key is split by point range at load time; jf_group_msm
 *     uploads each GPU's scalar slice, runs the slices concurrently and joins the partials with event-ordered peer reads (no
 *     NCCL, no spinning kernel).  The same device may be listed more than once (used by the single-GPU tests). */
int jf_group_create(const int *devices, int n_dev, jf_group **out);
void jf_group_destroy(jf_group *g);
int jf_group_size(const jf_group *g);
jf_ctx *jf_group_ctx(jf_group *g, int i);
const char *jf_group_last_error(const jf_group *g);
int jf_group_srs_load(jf_group *g, int curve, const void *affine_pts, size_t n, size_t stride_bytes, long inf_flag_offset,
                      int window_bits, int precompute, jf_group_srs **out);
int jf_group_srs_generate_for_testing(jf_group *g, int curve, const uint64_t *beta, size_t n, int window_bits, int precompute,
                                      jf_group_srs **out);
void jf_group_srs_free(jf_group *g, jf_group_srs *srs);
/* == jf_msm over the whole key: `msm_bigint(&powers_of_g[base_offset..], scalars).into_affine()` */
int jf_group_msm(jf_group *g, const jf_group_srs *srs, size_t base_offset, const uint64_t *scalars, size_t n,
                 int scalars_in_montgomery, uint64_t *out_xy, int *out_infinity);
/* batched transforms sharded by polynomial (vector b runs on GPU b mod size; no exchange step): == jf_ntt */
int jf_group_ntt(jf_group *g, int field, uint64_t *data, size_t in_len, unsigned log_n, int inverse,
                 const uint64_t *coset_offset, size_t batch, size_t batch_stride);

/* ---- NTT ----------------------------------------------------------------------------
 * In-place radix-2 transforms with `Radix2EvaluationDomain` semantics (ark-poly 0.4.2), natural
 * order in and out, `batch` vectors `batch_stride` ELEMENTS apart, field = BN254 Fr or BLS12-381 Fr.
 *   inverse == 0: `domain.fft_in_place` / `coset.fft` (plonk/src/proof_system/prover.rs:552-567):
 *       the first in_len entries are coefficients (the rest is treated as zero, i.e. arkworks'
 *       resize), out[i] = sum_j c[j] * (offset * w^i)^j.
 *   inverse != 0: `domain.ifft_in_place` (relation/src/constraint_system.rs:1172-1257) /
 *       `coset.ifft` (prover.rs:672): c[j] = offset^-j * n^-1 * sum_i e[i] * w^(-ij).
 * coset_offset: NULL for the plain domain, else 4 Montgomery limbs (`get_coset(Fr::GENERATOR)`,
 * prover.rs:545).  log_n above the field's two-adicity -> JF_ERR_DOMAIN_TOO_LARGE. */
int jf_ntt(jf_ctx *ctx, int field, uint64_t *data, size_t in_len, unsigned log_n, int inverse,
           const uint64_t *coset_offset, size_t batch, size_t batch_stride);
/* Same on a device pointer (data stays in HBM; used by the device-resident prover rows). */
int jf_ntt_device(jf_ctx *ctx, int field, void *d_data, size_t in_len, unsigned log_n, int inverse,
                  const uint64_t *coset_offset, size_t batch, size_t batch_stride);
/* Several cosets of the size-2^log_n domain in one batch.  The 8n-point quotient coset g<w_8n> of the prover
 * (plonk/src/proof_system/prover.rs:545,552-567) is the union of the cosets (g w_8n^r)<w_n>, r < 8, and the
 * quotient (degree 5n + 7, prover.rs:1126-1128) is already determined by six of them, so round 3 evaluates its
 * 25 polynomials on rows of n points instead of one 8n-point transform each.
 *   inverse == 0: `polys` coefficient vectors of in_len <= 2n entries, in_stride ELEMENTS apart, are read from
 *       polys_in; out[(p * rows + r) * n + i] = poly_p(offsets[r] * w_n^i)   (`get_coset(offsets[r]).fft`).
 *   inverse != 0: polys_in is ignored; `out` holds polys x rows rows of n values on those cosets and is
 *       overwritten in place by the n coefficients of each row's interpolant (`get_coset(offsets[r]).ifft`).
 * offsets: rows x 4 Montgomery limbs, 1 <= rows <= 16; 3 <= log_n <= two-adicity. */
int jf_ntt_cosets(jf_ctx *ctx, int field, const uint64_t *polys_in, size_t in_len, size_t in_stride, size_t polys,
                  unsigned log_n, int inverse, const uint64_t *offsets, int rows, uint64_t *out);

/* ---- TurboPlonk prover rounds around the two kernels -----------------
 * One TurboPlonk instance, 5 wire types, no Plookup.  Circuit construction stays with the caller
 * (relation/src/constraint_system.rs): it passes the selector columns `all_selectors()` (13 x n, order
 * q_lc0-3, q_mul0-1, q_hash0-3, q_o, q_c, q_ecc; :890-905), the extended permutation
 * `compute_extended_permutation()` (5 x n; :929-953), the coset representatives k (relation/src/constants.rs:30-79),
 * the wire -> variable map `wire_variables` (5 x n) and the io gate ids.  All field elements are 4-limb
 * Montgomery.
 *
 * jf_plonk_preprocess == `PlonkKzgSnark::preprocess` (plonk/src/proof_system/snark.rs:529-611): 18 iNTTs,
 * 18 commitments; the selector / sigma polynomials stay resident (the `ProvingKey`).  flags & 1: also
 * keep their 8n coset evaluations resident (18 of the 25 coset NTTs of round 3 then happen once per
 * key instead of once per proof; +4.5 GiB at n = 2^20).  flags & 2: selector columns that are identically
 * zero (e.g. q_hash / q_ecc in circuits without Rescue or ECC gates) are recognised once and their coset NTTs and
 * quotient terms are skipped; the proof is unchanged.  Round 3 evaluates the quotient on six sub-cosets
 * (g w_8n^r)<w_n>, r < 6, of the reference's 8n-point coset (6n points determine a polynomial of degree 5n + 7;
 * the coefficients follow from six size-n inverse transforms and a 6 x 6 Vandermonde solve per index), which
 * yields the same quotient polynomial with 29 % less transform work; flags & 8: the wire polynomials are committed in the Lagrange
 * basis (jf_srs_lagrange is run once for this key: the five wire commitments of a proof become MSMs over the witness values; same
 * commitments; ignored while the key is sharded over several GPUs); flags & 4 keeps the reference's form (one
 * 8n-point coset transform per polynomial, prover.rs:552-567,672), as do domains below 16 (at n = 8 six rows hold
 * exactly the quotient's 48 coefficients and nothing would be left for the WrongQuotientPolyDegree check).  `srs` must outlive the key and hold >= n + 3 points.  2 <= log_n and
 * log_n + 3 <= two-adicity (JF_ERR_DOMAIN_TOO_LARGE otherwise: `Prover::new`, prover.rs:54-62). */
int jf_plonk_preprocess(jf_ctx *ctx, const jf_srs *srs, unsigned log_n, const uint64_t *selector_evals,
                        const uint64_t *sigma_evals, const uint64_t *k, const uint32_t *wire_variables, size_t num_vars,
                        const uint32_t *pub_input_gate_ids, size_t num_inputs, int flags, jf_plonk_pk **out);
/* selector_comms (13) then sigma_comms (5) of the `VerifyingKey`: affine x || y each, plus infinity flags */
int jf_plonk_vk_commitments(jf_ctx *ctx, const jf_plonk_pk *pk, uint64_t *out_xy, int *out_inf);
void jf_plonk_pk_free(jf_ctx *ctx, jf_plonk_pk *pk);

/* `Proof<E>` (plonk/src/proof_system/structs.rs:62-84); points are x || y with 2*L limbs each, packed. */
typedef struct {
    int curve;
    uint64_t wires_poly_comms[5 * 12];
    int wires_inf[5];
    uint64_t prod_perm_poly_comm[12];
    int prod_perm_inf;
    uint64_t split_quot_poly_comms[5 * 12];
    int split_inf[5];
    uint64_t opening_proof[12];
    int opening_inf;
    uint64_t shifted_opening_proof[12];
    int shifted_opening_inf;
    uint64_t wires_evals[5 * 4];       /* ProofEvaluations, Montgomery */
    uint64_t wire_sigma_evals[4 * 4];
    uint64_t perm_next_eval[4];
    uint64_t challenges[5 * 4];        /* beta, gamma, alpha, zeta, v (Montgomery): diagnostics */
} jf_plonk_proof;

/* jf_plonk_prove == `PlonkKzgSnark::prove` -> `batch_prove_internal` for one instance (snark.rs:201-469)
 * with all five `Prover` rounds (prover.rs).  witness: num_vars Montgomery elements (`cs.witness`).
 * blinders: the 17 field elements the reference draws from its prng, in its consumption order
 * (5 x 2 wire masks, 3 for the permutation product, 4 split-quotient randomizers; prover.rs:463-486,946-957):
 * the caller keeps control of the randomness, which is also what makes proofs comparable byte for byte.
 * transcript_kind: 0 `SolidityTranscript` (Keccak-256), 1 `StandardTranscript` (Merlin).
 * extra_msg: `extra_transcript_init_msg` or NULL.  Fails with JF_ERR_QUOTIENT_DEGREE when the witness does
 * not satisfy the circuit (the reference's WrongQuotientPolyDegree). */
int jf_plonk_prove(jf_ctx *ctx, jf_plonk_pk *pk, const uint64_t *witness, const uint64_t *blinders, int transcript_kind,
                   const uint8_t *extra_msg, size_t extra_len, jf_plonk_proof *out);
/* ark-serialize `serialize_compressed` of the proof; returns the byte count (769 for BN254) or < 0. */
long jf_plonk_proof_serialize(const jf_plonk_proof *proof, uint8_t *out, size_t cap);

/* ---- UltraPlonk: the same prover with the Plookup argument (6 wire types) ---------------------------------------------------
 * `PlonkCircuit::new_ultra_plonk(range_bit_len)` circuits (relation/src/constraint_system.rs:234-240): range gates on a sixth
 * wire, key-value tables (`create_table_and_lookup_variables`), the q_lookup selector.  The caller passes what the
 * `Arithmetization` impl exposes after `finalize_for_arithmetization`: `all_selectors()` (14 x n: the 13 above, then q_lookup),
 * the extended permutation (6 x n), k (6), `wire_variables` (6 x n, the range wire last) and the three per-gate columns
 * `table_key_vec()`, `table_dom_sep_vec()`, `q_dom_sep()` (n each; :873-888); the range table {0 .. 2^range_bit_len - 1, 0 ..}
 * is built on the device.  jf_plonk_vk_commitments then yields 14 selector, 6 sigma and the 4 `PlookupVerifyingKey`
 * commitments (range table, key table, table dom sep, q dom sep; snark.rs:573-594), in that order.
 * Round 3 evaluates the quotient (degree 6 n + 8) on SEVEN sub-cosets of n points (from n = 16), like the six of the TurboPlonk
 * prover; flags as for jf_plonk_preprocess: 1 (resident coset evaluations of the 14 selector, 6 sigma and 4 table polynomials),
 * 2 (skip zero selectors), 4 (the reference's 8n-point coset instead), 8 (Lagrange-basis wire commitments). */
int jf_ultraplonk_preprocess(jf_ctx *ctx, const jf_srs *srs, unsigned log_n, const uint64_t *selector_evals,
                             const uint64_t *sigma_evals, const uint64_t *k, const uint32_t *wire_variables, size_t num_vars,
                             const uint32_t *pub_input_gate_ids, size_t num_inputs, unsigned range_bit_len,
                             const uint64_t *table_key_evals, const uint64_t *table_dom_sep_evals, const uint64_t *q_dom_sep_evals,
                             int flags, jf_plonk_pk **out);

/* `Proof<E>` with `plookup_proof: Some(PlookupProof)` (structs.rs:62-84,255-265,496-541) */
typedef struct {
    int curve;
    uint64_t wires_poly_comms[6 * 12];
    int wires_inf[6];
    uint64_t prod_perm_poly_comm[12];
    int prod_perm_inf;
    uint64_t split_quot_poly_comms[6 * 12];
    int split_inf[6];
    uint64_t opening_proof[12];
    int opening_inf;
    uint64_t shifted_opening_proof[12];
    int shifted_opening_inf;
    uint64_t wires_evals[6 * 4];
    uint64_t wire_sigma_evals[5 * 4];
    uint64_t perm_next_eval[4];
    uint64_t h_poly_comms[2 * 12];      /* PlookupProof */
    int h_inf[2];
    uint64_t prod_lookup_poly_comm[12];
    int prod_lookup_inf;
    uint64_t plookup_evals[15 * 4];     /* PlookupEvaluations in declaration order: range_table, key_table, table_dom_sep, q_dom_sep,
                                           h_1, q_lookup, prod_next, range_table_next, key_table_next, table_dom_sep_next, h_1_next,
                                           h_2_next, q_lookup_next, w_3_next, w_4_next */
    uint64_t challenges[6 * 4];         /* tau, beta, gamma, alpha, zeta, v (Montgomery): diagnostics */
} jf_ultraplonk_proof;

/* == `PlonkKzgSnark::prove` for one UltraPlonk instance.  blinders: the 29 field elements the reference draws, in its order:
 * 6 x 2 wire masks, 3 + 3 for h1 / h2 (`mask_polynomial(.., 2)`, prover.rs:113-114), 3 for the permutation product, 3 for the
 * lookup product, 5 split-quotient randomizers.  JF_ERR_INVALID_ARG when a lookup value is not in the table
 * ("The sorted vector has wrong length", constraint_system.rs:1411-1413), JF_ERR_QUOTIENT_DEGREE for an unsatisfied gate. */
int jf_ultraplonk_prove(jf_ctx *ctx, jf_plonk_pk *pk, const uint64_t *witness, const uint64_t *blinders, int transcript_kind,
                        const uint8_t *extra_msg, size_t extra_len, jf_ultraplonk_proof *out);
long jf_ultraplonk_proof_serialize(const jf_ultraplonk_proof *proof, uint8_t *out, size_t cap);

/* ---- several instances in one proof -----------------------------------------------------------------------------------------
 * == `PlonkKzgSnark::batch_prove` -> `batch_prove_internal` (snark.rs:201-469): `count` instances over the same domain size and the
 * same commit key share one transcript, ONE quotient polynomial (instance i enters with alpha_base_i = (alpha^3 | alpha^7)^i,
 * prover.rs:661-669) and ONE pair of opening proofs.  pks: one proving key per instance (each key owns its device workspace, so
 * the same circuit twice needs two keys; all built with the same flags).  blinders, in the order the reference's single prng is
 * consumed: the wire masks of every instance (2 x 5|6 each), [UltraPlonk: the h1 / h2 masks of every instance (6 each),] the
 * permutation-product masks (3 each), [the lookup-product masks (3 each),] then the 4|5 split-quotient randomizers:
 * count x 13 + 4 (TurboPlonk) or count x 24 + 5 (UltraPlonk) field elements.  out: `count` proof records; record i holds instance
 * i's commitments and evaluations, the shared parts (split quotient commitments, opening proofs) are written to every record.
 * The batch serialisers emit `BatchProof<E>` (structs.rs:271-292); a batch of one equals jf_plonk_prove. */
int jf_plonk_batch_prove(jf_ctx *ctx, jf_plonk_pk *const *pks, size_t count, const uint64_t *const *witnesses, const uint64_t *blinders,
                         int transcript_kind, const uint8_t *extra_msg, size_t extra_len, jf_plonk_proof *out);
int jf_ultraplonk_batch_prove(jf_ctx *ctx, jf_plonk_pk *const *pks, size_t count, const uint64_t *const *witnesses,
                              const uint64_t *blinders, int transcript_kind, const uint8_t *extra_msg, size_t extra_len,
                              jf_ultraplonk_proof *out);
long jf_plonk_batch_proof_serialize(const jf_plonk_proof *proofs, size_t count, uint8_t *out, size_t cap);
long jf_ultraplonk_batch_proof_serialize(const jf_ultraplonk_proof *proofs, size_t count, uint8_t *out, size_t cap);

/* ---- ONE proof on several GPUs (one process per GPU, jf_comm) ----------------------------------------------------------------
 * Every commitment of a proof is a sum over (coefficient, key point) pairs, so it splits by point range like jf_msm_sharded.  After
 * jf_plonk_pk_shard_commits every rank runs the SAME jf_plonk_prove / jf_plonk_batch_prove / jf_ultraplonk_prove call (same circuit,
 * witness, blinders, transcript: the polynomial algebra and the transforms are replicated, they hold identical polynomials in HBM),
 * but commits only coefficients [slice_start, slice_start + jf_srs_len(key_slice)) against its slice of the key; the XYZZ partials
 * cross NVLink (peer-memory mailboxes or ncclAllGather) before each commitment is normalised, so every rank sees the same
 * commitments, squeezes the same challenges and returns the same proof.  The 13 MSMs of a TurboPlonk proof then cost 1/nranks of
 * their accumulation each.  `pk` was built with the FULL key (jf_plonk_preprocess is not collective); comm == NULL returns the key to
 * one-GPU operation.  Calls are collective: the ranks must issue them in the same order.
 * shard_round3 != 0 additionally deals out round 3 by sub-coset: the quotient is evaluated on 6 (UltraPlonk: 7) cosets of n points
 * that do not interact until the final Vandermonde solve, so rank r mod nranks transforms the 25 (35) polynomials onto coset r only,
 * evaluates the quotient there and transforms it back; the n-coefficient interpolants of the rows are then broadcast (NCCL over
 * NVLink: 32 n bytes per row) and every rank solves for the same quotient polynomial.  Ignored for keys built with flags & 1
 * (resident coset evaluations) or flags & 4 / n < 16 (8n-point coset). */
int jf_plonk_pk_shard_commits(jf_ctx *ctx, jf_plonk_pk *pk, jf_comm *comm, const jf_srs *key_slice, size_t slice_start, int shard_round3);

/* ---- proof linking ------------------------------------------------------------------------------------------------------------
 * `PlonkKzgSnark::link_proofs` (plonk/src/proof_system/proof_linking.rs:79-216): two TurboPlonk proofs whose circuits placed the
 * same link group (`GroupLayout { alignment, offset, size }`, relation/src/proof_linking/mod.rs:17-53) carry the group's values in
 * their first wire polynomial at the points g^(offset + i), g the 2^alignment-th root of unity.  The linking proof is the commitment
 * of q = (a1 - a2) / Z_D (Z_D the vanishing polynomial of those points; remainder dropped like ark-poly's `/`) and a KZG opening
 * (`UnivariateKzgPCS::open`, mod.rs:135-161) of a1 - a2 - q Z_D(eta) at the challenge eta squeezed from a fresh transcript over
 * (a1's commitment, a2's commitment, q's commitment).  On the device: a1 - a2 is evaluated on the 2^alignment-th roots of unity; when
 * it vanishes on the link domain the quotient is a pointwise ratio on a coset (three transforms, independent of `size`); otherwise
 * -- the two proofs are NOT linked and the verifier will reject -- the remainder (a1 - a2) mod Z_D is taken off first (Z_D divides
 * X^(2^alignment) - 1, so it is a fold and a short schoolbook reduction) and the same division yields the floor quotient; groups
 * aligned above 2^12 fall back to `size` divisions by a linear factor.  flags & 1 forces that last form.  The circuit-side layout (`LinkableCircuit`) stays with the caller. */
typedef struct {
    int curve;
    uint64_t quotient_commitment[12];  /* `LinkingProof::quotient_commitment` (proof_linking.rs:32-39) */
    int quotient_inf;
    uint64_t opening_proof[12];        /* `LinkingProof::opening_proof` */
    int opening_inf;
    uint64_t eta[4];                   /* the opening challenge (Montgomery): diagnostics */
    int path;                          /* 0: exact division on a coset, 2: the same after taking the remainder off, 1: successive linear
                                          divisions: diagnostics */
} jf_link_proof;
/* `LinkingHint::linking_wire_poly` (structs.rs:88-97; snark.rs:96-100) of the LAST proof made with `pk`: the first wire polynomial
 * after masking, n + 2 Montgomery coefficients (cap >= n + 2 elements).  The hint's commitment is wires_poly_comms[0] of that proof. */
int jf_plonk_link_hint(jf_ctx *ctx, const jf_plonk_pk *pk, uint64_t *out_poly, size_t cap, size_t *out_len);
/* == `link_proofs(lhs_link_hint, rhs_link_hint, group_layout, commit_key)` with the hints in host memory: a1 / a2 Montgomery
 * coefficients (low degree first), their commitments as affine x || y (+ infinity flag).  transcript_kind as jf_plonk_prove. */
int jf_plonk_link_proofs(jf_ctx *ctx, const jf_srs *srs, const uint64_t *a1, size_t len1, const uint64_t *a1_comm_xy, int a1_inf,
                         const uint64_t *a2, size_t len2, const uint64_t *a2_comm_xy, int a2_inf, unsigned alignment, size_t offset,
                         size_t size, int transcript_kind, int flags, jf_link_proof *out);
/* The same with both wire polynomials still in HBM: the workspaces of the two proving keys hold them after jf_plonk_prove
 * (the keys must share one commit key; no polynomial crosses PCIe). */
int jf_plonk_link_proofs_resident(jf_ctx *ctx, const jf_plonk_pk *lhs, const jf_plonk_proof *lhs_proof, const jf_plonk_pk *rhs,
                                  const jf_plonk_proof *rhs_proof, unsigned alignment, size_t offset, size_t size, int transcript_kind,
                                  int flags, jf_link_proof *out);
/* floor(p / Z_D) for `batch` polynomials (host memory, lens[i] Montgomery coefficients; out_quotients[i] receives
 * max(lens[i] - size, 0)).  Division by the PUBLIC vanishing polynomial is linear, so the collaborative prover's
 * `compute_linking_quotient` (plonk/src/multiprover/proof_system/proof_linking.rs:127-138) is this call on the share, the MAC and the
 * public-modifier vector of a1 - a2; the shares' remainders do not vanish one by one, so they go through the remainder step above.  flags as jf_plonk_link_proofs.  field: JF_BN254_FR / JF_BLS12_381_FR. */
int jf_poly_div_link_domain(jf_ctx *ctx, int field, const uint64_t *const *polys, const size_t *lens, size_t batch, unsigned alignment,
                            size_t offset, size_t size, int flags, uint64_t *const *out_quotients);
/* ark-serialize `serialize_compressed` of `LinkingProof<E>`: 64 bytes (BN254) / 96 (BLS12-381); returns the count or < 0 */
long jf_link_proof_serialize(const jf_link_proof *proof, uint8_t *out, size_t cap);

/* Host-only pieces of the transcripts (no GPU needed): sha3 `Keccak256`, and `PlonkTranscript`
 * new / append_message / get_and_append_challenge (plonk/src/transcript/{solidity,standard}.rs). */
void jf_keccak256(const uint8_t *data, size_t len, uint8_t out[32]);
void *jf_transcript_new(int kind, const char *label);
void jf_transcript_free(void *t);
void jf_transcript_append(void *t, const char *label, const uint8_t *msg, size_t len);
int jf_transcript_challenge(void *t, int field, const char *label, uint64_t *out_montgomery);

/* ---- device buffers (plumbing for callers that keep polynomials resident) ------------ */
int jf_dev_alloc(jf_ctx *ctx, size_t bytes, void **out);
int jf_dev_free(jf_ctx *ctx, void *ptr);
int jf_dev_upload(jf_ctx *ctx, void *dst, const void *src, size_t bytes);
int jf_dev_download(jf_ctx *ctx, void *dst, const void *src, size_t bytes);
/* Page-locked host staging buffers (what the shim crate uses for coefficient vectors). */
int jf_host_alloc(jf_ctx *ctx, size_t bytes, void **out);
int jf_host_free(jf_ctx *ctx, void *ptr);

/* ---- element-wise field kernels (K1 parity tests and the "next" rows) ----------------
 * op: 0 mul, 1 add, 2 sub, 3 sqr, 4 inv, 5 to_mont, 6 from_mont, 7 neg; host arrays of n elements. */
int jf_field_op(jf_ctx *ctx, int field, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
/* out[i] = scalars[i] * G (canonical 4-limb scalars) as affine x || y: fixed-base helper used by
 * the tests to build point sets with known discrete logs. */
int jf_fixed_base_mul(jf_ctx *ctx, int curve, const uint64_t *scalars, size_t n, uint64_t *out_xy);

/* ---- measurement hooks (bench.py) ------------------------------------------------------
 * Per-kernel CUDA-event timing on the context's stream.  While enabled (on = 1) every kernel launch is
 * bracketed by two events (this costs ~0.2 ms per 2^20 MSM: events between the ~30 dependent launches); on = 2
 * brackets only the dominant kernels (msm_accumulate, ntt_pass), which leaves the step time unchanged; jf_profile_collect synchronises, writes one line per kernel name
 * ("name launches total_ms\n") into buf and resets the log.  Returns the number of bytes
 * written or a negative jf_status. */
int jf_profile_enable(jf_ctx *ctx, int on);
long jf_profile_collect(jf_ctx *ctx, char *buf, size_t cap);
/* Integer-pipe micro-benchmarks that give the MSM / NTT kernels their compute roof:
 * kind 0: independent IMAD.WIDE.U32 chains -> 32x32->64 multiply-adds per second;
 * kind 1: dependent 256-bit Montgomery multiplications (BN254 Fq) in registers -> field muls/s.
 * *out_rate = operations per second over the whole GPU. */
int jf_microbench(jf_ctx *ctx, int kind, double *out_rate);

#ifdef __cplusplus
}
#endif
#endif /* JF_B200_H */
