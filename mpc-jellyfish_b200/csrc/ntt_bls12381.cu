// ntt_bls12381.cu -- the NTT kernels of ntt_impl.cuh instantiated for Bls12381Fr.
#include "ntt_impl.cuh"

namespace jf {

int ntt_run_bls12381(jf_ctx *ctx, int field, void *d_data, void *d_out, size_t in_len, unsigned log_n, int inverse,
               const uint64_t *coset_offset, size_t batch, size_t batch_stride) {
    return ntt_run_t<Bls12381Fr>(ctx, field, d_data, d_out, in_len, log_n, inverse, coset_offset, batch, batch_stride);
}

int ntt_run_cosets_bls12381(jf_ctx *ctx, int field, const void *d_src, size_t src_stride, size_t in_len, void *d_dst, unsigned log_n,
                         int inverse, const uint64_t *offsets, int rows, size_t polys) {
    return ntt_run_cosets_t<Bls12381Fr>(ctx, field, d_src, src_stride, in_len, d_dst, log_n, inverse, offsets, rows, polys);
}

}  // namespace jf
