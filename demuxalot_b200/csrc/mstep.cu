// M-step: variant-segmented accumulation of posterior-weighted counts (demux.py:113-118).
//
//   addition[v, g] = float32( sum_{rows r of variant v, ascending}^{float64} ( post[cb_r, g] * (1 - e_r) ) ^ power ),  g < G
//
// Rows come in the reference's own order (CSC: ascending variant, then barcode), so every float64 sum is taken in
// exactly the order np.bincount uses and the result is bit-exact given identical posteriors (power == 2).
// One warp per variant, lanes over genotypes (coalesced 128-byte reads of the singlet-posterior rows, which are
// L2-resident: B x G x 4 bytes), rows unrolled by 4 for memory-level parallelism.  HBM / L2-gather bound:
// 8 bytes of row records + 4G bytes of gathered posteriors per row, 4G bytes written per variant.  No atomics.
#include "common.cuh"

namespace dmx {

template <int SLOTS, bool SQUARE>
__global__ void __launch_bounds__(256) mstep_kernel(const int64_t* __restrict__ offsets,
                                                    const int32_t* __restrict__ cb_arr,
                                                    const float* __restrict__ e_arr, const float* __restrict__ post,
                                                    int64_t ld_post, int n_genotypes, float power,
                                                    float* __restrict__ addition, int64_t ld_add,
                                                    double* __restrict__ addition64, int64_t ld_add64,
                                                    int64_t variant_lo, int64_t variant_hi) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;

    for (int64_t v = variant_lo + warp_global; v < variant_hi; v += n_warps) {
        const int64_t lo = offsets[v], hi = offsets[v + 1];
        double acc[SLOTS];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) acc[s] = 0.0;

        int64_t r = lo;
        for (; r + 4 <= hi; r += 4) {
            int32_t cb[4];
            float w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                cb[u] = cb_arr[r + u];
                w[u] = __fsub_rn(1.f, e_arr[r + u]);
            }
            float x[4][SLOTS];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int g = lane + 32 * s;
                    x[u][s] = (g < n_genotypes) ? __ldg(post + (int64_t)cb[u] * ld_post + g) : 0.f;
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)  // rows strictly in order: float64 addition is not associative
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const float c = __fmul_rn(x[u][s], w[u]);
                    acc[s] += (double)(SQUARE ? __fmul_rn(c, c) : powf(c, power));
                }
        }
        for (; r < hi; ++r) {
            const int32_t cb = cb_arr[r];
            const float w = __fsub_rn(1.f, e_arr[r]);
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int g = lane + 32 * s;
                const float x = (g < n_genotypes) ? __ldg(post + (int64_t)cb * ld_post + g) : 0.f;
                const float c = __fmul_rn(x, w);
                acc[s] += (double)(SQUARE ? __fmul_rn(c, c) : powf(c, power));
            }
        }
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int g = lane + 32 * s;
            if (g < n_genotypes) {
                if (addition) addition[v * ld_add + g] = (float)acc[s];
                if (addition64) addition64[v * ld_add64 + g] = acc[s];
            }
        }
    }
}

template <bool SQUARE>
static int launch_mstep(int slots, int grid, cudaStream_t stream, const int64_t* offsets, const int32_t* cb,
                        const float* e, const float* post, int64_t ld_post, int G, float power, float* addition,
                        int64_t ld_add, double* addition64, int64_t ld_add64, int64_t v_lo, int64_t v_hi) {
#define DMX_MSTEP_CASE(S)                                                                                          \
    case S:                                                                                                        \
        mstep_kernel<S, SQUARE><<<grid, 256, 0, stream>>>(offsets, cb, e, post, ld_post, G, power, addition, ld_add, \
                                                          addition64, ld_add64, v_lo, v_hi);                      \
        break;
    switch (slots) {
        DMX_MSTEP_CASE(1)
        DMX_MSTEP_CASE(2)
        DMX_MSTEP_CASE(4)
        DMX_MSTEP_CASE(8)
        DMX_MSTEP_CASE(16)
        default:
            set_error("M-step supports up to 512 genotypes");
            return -2;
    }
#undef DMX_MSTEP_CASE
    DMX_LAUNCH_CHECK();
    return 0;
}

}  // namespace dmx

extern "C" {

int dmx_mstep(const int64_t* variant_offsets, const int32_t* csc_cb, const float* csc_e,
              const float* singlet_posteriors, int64_t ld_singlet, int32_t n_genotypes, double power,
              float* addition, int64_t ld_addition, double* addition64, int64_t ld_addition64, int64_t variant_lo,
              int64_t variant_hi, void* stream_) {
    using namespace dmx;
    if (variant_hi <= variant_lo || n_genotypes <= 0) return 0;
    DMX_REQUIRE(addition || addition64, "no output buffer");
    const int slots_needed = (int)ceil_div(n_genotypes, 32);
    int slots = 1;
    while (slots < slots_needed) slots *= 2;
    const int64_t n_variants = variant_hi - variant_lo;
    const int warps_per_block = 8;
    int64_t blocks = ceil_div(n_variants, warps_per_block);
    const int64_t cap = (int64_t)sm_count() * 8 * 4;  // 8 resident CTAs per SM x 4 waves, grid-stride beyond
    if (blocks > cap) blocks = cap;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (power == 2.0)
        return launch_mstep<true>(slots, (int)blocks, stream, variant_offsets, csc_cb, csc_e, singlet_posteriors,
                                  ld_singlet, n_genotypes, 2.f, addition, ld_addition, addition64, ld_addition64,
                                  variant_lo, variant_hi);
    return launch_mstep<false>(slots, (int)blocks, stream, variant_offsets, csc_cb, csc_e, singlet_posteriors,
                               ld_singlet, n_genotypes, (float)power, addition, ld_addition, addition64,
                               ld_addition64, variant_lo, variant_hi);
}

}  // extern "C"
