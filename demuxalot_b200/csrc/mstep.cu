// M-step: variant-segmented accumulation of posterior-weighted counts (demux.py:113-118).
//
//   addition[v, g] = float32( sum_{rows r of variant v}^{float64} ( post[cb_r, g] * (1 - e_r) ) ^ power ),  g < G
//
// Rows come in the reference's own order (CSC: ascending variant, then barcode).  One warp per variant, lanes over
// genotypes.  A warp reads the row records of 32 rows with one coalesced load per array, then walks them with
// shuffles: every row is one coalesced 4G-byte gather of the singlet posteriors (B x G x 4 bytes, L2 resident),
// eight gathers in flight per warp.  Terms are added in row order in float64, i.e. exactly the order np.bincount
// uses, so the result is bit-exact given identical posteriors (power == 2).  Variants with more than HEAVY_ROWS
// rows (expression skew: a few variants are seen in most barcodes) are set aside and processed by all warps of
// the CTA together, partial sums combined in warp order -- still deterministic, no atomics.
// HBM / L2-gather bound: 8 bytes of row records + 4G gathered bytes per row, 4G bytes written per variant.
#include "common.cuh"

namespace dmx {

constexpr int MSTEP_WARPS = 8;
constexpr int VARIANTS_PER_CTA = 64;
constexpr int HEAVY_ROWS = 2048;

template <int SLOTS, bool SQUARE>
__device__ __forceinline__ void accumulate_batch(double (&acc)[SLOTS], const int32_t* __restrict__ cb_arr,
                                                 const float* __restrict__ e_arr, int64_t base, int n,
                                                 const float* __restrict__ post, int64_t ld_post, int n_genotypes,
                                                 float power, int lane) {
    const int32_t my_cb = lane < n ? __ldg(cb_arr + base + lane) : 0;
    const float my_w = lane < n ? __fsub_rn(1.f, __ldg(e_arr + base + lane)) : 0.f;
    int k = 0;
    for (; k + 8 <= n; k += 8) {
        float x[8][SLOTS], w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int32_t cb = __shfl_sync(0xffffffffu, my_cb, k + u);
            w[u] = __shfl_sync(0xffffffffu, my_w, k + u);
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int g = lane + 32 * s;
                x[u][s] = (g < n_genotypes) ? __ldg(post + (int64_t)cb * ld_post + g) : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)  // rows strictly in order: float64 addition is not associative
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const float c = __fmul_rn(x[u][s], w[u]);
                acc[s] += (double)(SQUARE ? __fmul_rn(c, c) : powf(c, power));
            }
    }
    for (; k < n; ++k) {
        const int32_t cb = __shfl_sync(0xffffffffu, my_cb, k);
        const float w = __shfl_sync(0xffffffffu, my_w, k);
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int g = lane + 32 * s;
            const float x = (g < n_genotypes) ? __ldg(post + (int64_t)cb * ld_post + g) : 0.f;
            const float c = __fmul_rn(x, w);
            acc[s] += (double)(SQUARE ? __fmul_rn(c, c) : powf(c, power));
        }
    }
}

template <int SLOTS, bool SQUARE>
__global__ void __launch_bounds__(MSTEP_WARPS * 32) mstep_kernel(
    const int64_t* __restrict__ offsets, const int32_t* __restrict__ cb_arr, const float* __restrict__ e_arr,
    const float* __restrict__ post, int64_t ld_post, int n_genotypes, float power, float* __restrict__ addition,
    int64_t ld_add, double* __restrict__ addition64, int64_t ld_add64, int64_t variant_lo, int64_t variant_hi) {
    __shared__ int heavy_list[VARIANTS_PER_CTA];
    __shared__ int heavy_count;
    __shared__ double partial[MSTEP_WARPS][SLOTS * 32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) heavy_count = 0;
    __syncthreads();

    auto store = [&](int64_t v, const double (&acc)[SLOTS]) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int g = lane + 32 * s;
            if (g < n_genotypes) {
                if (addition) addition[v * ld_add + g] = (float)acc[s];
                if (addition64) addition64[v * ld_add64 + g] = acc[s];
            }
        }
    };

    const int64_t v0 = variant_lo + (int64_t)blockIdx.x * VARIANTS_PER_CTA;
    for (int k = warp; k < VARIANTS_PER_CTA; k += MSTEP_WARPS) {
        const int64_t v = v0 + k;
        if (v >= variant_hi) break;
        const int64_t lo = offsets[v], hi = offsets[v + 1];
        if (hi - lo > HEAVY_ROWS) {
            if (lane == 0) heavy_list[atomicAdd(&heavy_count, 1)] = k;  // order of the list does not matter
            continue;
        }
        double acc[SLOTS];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) acc[s] = 0.0;
        for (int64_t base = lo; base < hi; base += 32) {
            const int n = (int)(hi - base < 32 ? hi - base : 32);
            accumulate_batch<SLOTS, SQUARE>(acc, cb_arr, e_arr, base, n, post, ld_post, n_genotypes, power, lane);
        }
        store(v, acc);
    }
    __syncthreads();

    // heavy variants: the CTA's warps take contiguous slices of the rows; partial sums are added in warp order
    const int n_heavy = heavy_count;
    for (int h = 0; h < n_heavy; ++h) {
        const int64_t v = v0 + heavy_list[h];
        const int64_t lo = offsets[v], hi = offsets[v + 1];
        const int64_t n_batches = (hi - lo + 31) / 32;
        const int64_t per_warp = (n_batches + MSTEP_WARPS - 1) / MSTEP_WARPS;
        const int64_t b_lo = warp * per_warp;
        const int64_t b_hi = b_lo + per_warp < n_batches ? b_lo + per_warp : n_batches;
        double acc[SLOTS];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) acc[s] = 0.0;
        for (int64_t b = b_lo; b < b_hi; ++b) {
            const int64_t base = lo + 32 * b;
            const int n = (int)(hi - base < 32 ? hi - base : 32);
            accumulate_batch<SLOTS, SQUARE>(acc, cb_arr, e_arr, base, n, post, ld_post, n_genotypes, power, lane);
        }
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) partial[warp][s * 32 + lane] = acc[s];
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                double sum = partial[0][s * 32 + lane];
                for (int w = 1; w < MSTEP_WARPS; ++w) sum += partial[w][s * 32 + lane];
                acc[s] = sum;
            }
            store(v, acc);
        }
        __syncthreads();
    }
}

template <bool SQUARE>
static int launch_mstep(int slots, unsigned grid, cudaStream_t stream, const int64_t* offsets, const int32_t* cb,
                        const float* e, const float* post, int64_t ld_post, int G, float power, float* addition,
                        int64_t ld_add, double* addition64, int64_t ld_add64, int64_t v_lo, int64_t v_hi) {
#define DMX_MSTEP_CASE(S)                                                                                       \
    case S:                                                                                                     \
        mstep_kernel<S, SQUARE><<<grid, MSTEP_WARPS * 32, 0, stream>>>(offsets, cb, e, post, ld_post, G, power, \
                                                                       addition, ld_add, addition64, ld_add64,  \
                                                                       v_lo, v_hi);                             \
        break;
    switch (slots) {
        DMX_MSTEP_CASE(1)
        DMX_MSTEP_CASE(2)
        DMX_MSTEP_CASE(4)
        DMX_MSTEP_CASE(8)
        default:
            set_error("M-step supports up to 256 genotypes");
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
    const int64_t blocks = ceil_div(variant_hi - variant_lo, VARIANTS_PER_CTA);
    DMX_REQUIRE(blocks < (1ll << 31), "grid too large");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (power == 2.0)
        return launch_mstep<true>(slots, (unsigned)blocks, stream, variant_offsets, csc_cb, csc_e, singlet_posteriors,
                                  ld_singlet, n_genotypes, 2.f, addition, ld_addition, addition64, ld_addition64,
                                  variant_lo, variant_hi);
    return launch_mstep<false>(slots, (unsigned)blocks, stream, variant_offsets, csc_cb, csc_e, singlet_posteriors,
                               ld_singlet, n_genotypes, (float)power, addition, ld_addition, addition64,
                               ld_addition64, variant_lo, variant_hi);
}

}  // extern "C"
