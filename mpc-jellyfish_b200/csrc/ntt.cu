// ntt.cu -- plan bookkeeping and the field dispatch of the NTT; the kernels live in ntt_impl.cuh and are
// instantiated per field in ntt_bn254.cu / ntt_bls12381.cu (two translation units compile in parallel).
#include "ntt_impl.cuh"

namespace jf {

int ntt_run_bn254(jf_ctx *ctx, int field, void *d_data, void *d_out, size_t in_len, unsigned log_n, int inverse,
                  const uint64_t *coset_offset, size_t batch, size_t batch_stride);
int ntt_run_bls12381(jf_ctx *ctx, int field, void *d_data, void *d_out, size_t in_len, unsigned log_n, int inverse,
                     const uint64_t *coset_offset, size_t batch, size_t batch_stride);

int ntt_run_cosets_bn254(jf_ctx *ctx, int field, const void *d_src, size_t src_stride, size_t in_len, void *d_dst, unsigned log_n,
                         int inverse, const uint64_t *offsets, int rows, size_t polys);
int ntt_run_cosets_bls12381(jf_ctx *ctx, int field, const void *d_src, size_t src_stride, size_t in_len, void *d_dst, unsigned log_n,
                            int inverse, const uint64_t *offsets, int rows, size_t polys);

void ntt_free_plans(jf_ctx *ctx) {
    for (auto &kv : ctx->ntt_plans) free_plan(kv.second);
    ctx->ntt_plans.clear();
}

int ntt_run(jf_ctx *ctx, int field, void *d_data, void *d_out, size_t in_len, unsigned log_n, int inverse,
            const uint64_t *coset_offset, size_t batch, size_t batch_stride) {
    if (field == JF_BN254_FR)
        return ntt_run_bn254(ctx, field, d_data, d_out, in_len, log_n, inverse, coset_offset, batch, batch_stride);
    if (field == JF_BLS12_381_FR)
        return ntt_run_bls12381(ctx, field, d_data, d_out, in_len, log_n, inverse, coset_offset, batch, batch_stride);
    return fail(ctx, JF_ERR_INVALID_ARG, "ntt: field must be BN254 Fr or BLS12-381 Fr");
}

int ntt_run_cosets(jf_ctx *ctx, int field, const void *d_src, size_t src_stride, size_t in_len, void *d_dst, unsigned log_n,
                   int inverse, const uint64_t *offsets, int rows, size_t polys) {
    if (field == JF_BN254_FR) return ntt_run_cosets_bn254(ctx, field, d_src, src_stride, in_len, d_dst, log_n, inverse, offsets, rows, polys);
    if (field == JF_BLS12_381_FR)
        return ntt_run_cosets_bls12381(ctx, field, d_src, src_stride, in_len, d_dst, log_n, inverse, offsets, rows, polys);
    return fail(ctx, JF_ERR_INVALID_ARG, "ntt: field must be BN254 Fr or BLS12-381 Fr");
}

}  // namespace jf
