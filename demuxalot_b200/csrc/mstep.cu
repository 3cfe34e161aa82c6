// M-step: variant-segmented accumulation of posterior-weighted counts (demux.py:113-118).
//
//   addition[v, g] = float32( sum_{rows r of variant v}^{float64} ( post[cb_r, g] * (1 - e_r) ) ^ power ),  g < G
//
// Rows come in the reference's own order (CSC: ascending variant, then barcode).  One warp per variant.  A warp
// reads the row records of 32 rows with one coalesced load per array and walks them with shuffles.  Each lane owns
// 4 consecutive genotypes (one 128-bit load of the singlet-posterior row, which is L2 resident: B x G x 4 bytes),
// so a row needs LPR = G/4 lanes and a warp works on 32/LPR rows at once, four such waves in flight.  Every lane adds its
// rows in ascending order in float64; the 32/LPR row groups are then combined in a fixed order, i.e. the result is
// deterministic and equals the reference's sequential float64 sum up to float64 re-association (a difference is
// only visible if the exact sum lies within 1e-16 relative of a float32 rounding boundary).  Variants with more
// than HEAVY_ROWS rows (expression skew: a few variants are seen in most barcodes) are processed by a second launch
// in which all warps of a CTA share one variant.  No atomics anywhere.
// HBM / L2-gather bound: 8 bytes of row records + 4G gathered bytes per row, 4G bytes written per variant.
#include "common.cuh"

namespace dmx {

constexpr int MSTEP_WARPS = 8;
constexpr int HEAVY_ROWS = 4096;
constexpr int WAVES_IN_FLIGHT = 4;  // row waves whose gathers are issued before any is consumed (measured: 2-4 best)

template <int LPR, int SLOTS, bool SQUARE, bool FULL>
struct RowWalker {
    static constexpr int RGW = 32 / LPR;        // row groups per warp
    static constexpr int WAVES = WAVES_IN_FLIGHT;
    double acc[SLOTS][4];
    int lane, sub, rgw;

    __device__ __forceinline__ void init(int lane_) {
        lane = lane_;
        sub = lane % LPR;
        rgw = lane / LPR;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[s][c] = 0.0;
    }

    // rows [base, base + n), n <= 32.  Rows past n are clamped to row 0 with weight 0 (their term is exactly 0
    // for the squared contribution; the general power masks them explicitly because 0^0 = 1).
    __device__ __forceinline__ void batch(const int32_t* __restrict__ cb_arr, const float* __restrict__ e_arr,
                                          int64_t base, int n, const float* __restrict__ post, int64_t ld_post,
                                          int n_quads, float power) {
        const int32_t my_cb = __ldg(cb_arr + base + (lane < n ? lane : 0));
        const float my_w = lane < n ? __fsub_rn(1.f, __ldg(e_arr + base + lane)) : 0.f;
        for (int k0 = 0; k0 < n; k0 += RGW * WAVES) {
            float4 x[WAVES][SLOTS];
            float w[WAVES];
#pragma unroll
            for (int u = 0; u < WAVES; ++u) {
                const int k = k0 + u * RGW + rgw;  // k >= n reads lane k's clamped record: weight 0
                const int32_t cb = __shfl_sync(0xffffffffu, my_cb, k & 31);
                w[u] = __shfl_sync(0xffffffffu, my_w, k & 31);
                const float4* row = reinterpret_cast<const float4*>(post + (int64_t)cb * ld_post);
                const bool wave_has_rows = k0 + u * RGW < n;  // warp-uniform: whole waves past the end are skipped
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int q = sub + LPR * s;
                    x[u][s] = (wave_has_rows && (FULL || q < n_quads)) ? __ldg(row + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < WAVES; ++u)
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const float v[4] = {x[u][s].x, x[u][s].y, x[u][s].z, x[u][s].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float t = __fmul_rn(v[c], w[u]);
                        if (SQUARE) {
                            acc[s][c] += (double)__fmul_rn(t, t);
                        } else {
                            const bool valid = k0 + u * RGW + rgw < n;
                            acc[s][c] += (double)(valid ? powf(t, power) : 0.f);
                        }
                    }
                }
        }
    }

    // fixed-order combination of the row groups; afterwards every lane of group 0 holds the totals
    __device__ __forceinline__ void combine() {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                double total = acc[s][c];
#pragma unroll
                for (int g = 1; g < RGW; ++g) total += __shfl_sync(0xffffffffu, acc[s][c], (sub + g * LPR) & 31);
                acc[s][c] = total;
            }
    }

    __device__ __forceinline__ void store(int64_t v, int n_genotypes, float* __restrict__ addition, int64_t ld_add,
                                          double* __restrict__ addition64, int64_t ld_add64) const {
        if (rgw != 0) return;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int g = 4 * (sub + LPR * s) + c;
                if (g < n_genotypes) {
                    if (addition) addition[v * ld_add + g] = (float)acc[s][c];
                    if (addition64) addition64[v * ld_add64 + g] = acc[s][c];
                }
            }
    }
};

// one warp per variant; variants above HEAVY_ROWS are left to mstep_heavy_kernel
template <int LPR, int SLOTS, bool SQUARE, bool FULL>
__global__ void __launch_bounds__(MSTEP_WARPS * 32) mstep_kernel(
    const int64_t* __restrict__ offsets, const int32_t* __restrict__ cb_arr, const float* __restrict__ e_arr,
    const float* __restrict__ post, int64_t ld_post, int n_genotypes, float power, float* __restrict__ addition,
    int64_t ld_add, double* __restrict__ addition64, int64_t ld_add64, int64_t variant_lo, int64_t variant_hi) {
    const int lane = threadIdx.x & 31;
    const int64_t v = variant_lo + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (v >= variant_hi) return;
    const int64_t lo = offsets[v], hi = offsets[v + 1];
    if (hi - lo > HEAVY_ROWS) return;
    const int n_quads = (int)((ld_post + 3) / 4);
    RowWalker<LPR, SLOTS, SQUARE, FULL> walker;
    walker.init(lane);
    for (int64_t base = lo; base < hi; base += 32)
        walker.batch(cb_arr, e_arr, base, (int)(hi - base < 32 ? hi - base : 32), post, ld_post, n_quads, power);
    walker.combine();
    walker.store(v, n_genotypes, addition, ld_add, addition64, ld_add64);
}

// heavy variants: every CTA scans a block of variants; the (rare) ones above HEAVY_ROWS are processed by all warps
// of the CTA on contiguous slices of the rows, partial sums combined in warp order
template <int LPR, int SLOTS, bool SQUARE, bool FULL>
__global__ void __launch_bounds__(MSTEP_WARPS * 32) mstep_heavy_kernel(
    const int64_t* __restrict__ offsets, const int32_t* __restrict__ cb_arr, const float* __restrict__ e_arr,
    const float* __restrict__ post, int64_t ld_post, int n_genotypes, float power, float* __restrict__ addition,
    int64_t ld_add, double* __restrict__ addition64, int64_t ld_add64, int64_t variant_lo, int64_t variant_hi) {
    __shared__ double partial[MSTEP_WARPS][SLOTS * 4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_quads = (int)((ld_post + 3) / 4);
    __shared__ int heavy_list[MSTEP_WARPS * 32];
    __shared__ int heavy_count;
    const int64_t v0 = variant_lo + (int64_t)blockIdx.x * (MSTEP_WARPS * 32);
    if (threadIdx.x == 0) heavy_count = 0;
    __syncthreads();
    {   // every thread looks at one variant; the order of the list does not influence any result
        const int64_t v = v0 + threadIdx.x;
        if (v < variant_hi && offsets[v + 1] - offsets[v] > HEAVY_ROWS) heavy_list[atomicAdd(&heavy_count, 1)] = threadIdx.x;
    }
    __syncthreads();
    const int n_heavy = heavy_count;
    for (int h = 0; h < n_heavy; ++h) {
        const int64_t v = v0 + heavy_list[h];
        const int64_t lo = offsets[v], hi = offsets[v + 1];
        const int64_t n_batches = (hi - lo + 31) / 32;
        const int64_t per_warp = (n_batches + MSTEP_WARPS - 1) / MSTEP_WARPS;
        const int64_t b_lo = warp * per_warp;
        const int64_t b_hi = b_lo + per_warp < n_batches ? b_lo + per_warp : n_batches;
        RowWalker<LPR, SLOTS, SQUARE, FULL> walker;
        walker.init(lane);
        for (int64_t b = b_lo; b < b_hi; ++b) {
            const int64_t base = lo + 32 * b;
            walker.batch(cb_arr, e_arr, base, (int)(hi - base < 32 ? hi - base : 32), post, ld_post, n_quads, power);
        }
        walker.combine();
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) partial[warp][s * 4 + c][lane] = walker.acc[s][c];
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    double sum = partial[0][s * 4 + c][lane];
                    for (int w = 1; w < MSTEP_WARPS; ++w) sum += partial[w][s * 4 + c][lane];
                    walker.acc[s][c] = sum;
                }
            walker.store(v, n_genotypes, addition, ld_add, addition64, ld_add64);
        }
        __syncthreads();
    }
}

template <int LPR, int SLOTS, bool SQUARE, bool FULL>
static int launch_pair(cudaStream_t stream, const int64_t* offsets, const int32_t* cb, const float* e,
                       const float* post, int64_t ld_post, int G, float power, float* addition, int64_t ld_add,
                       double* addition64, int64_t ld_add64, int64_t v_lo, int64_t v_hi) {
    const int64_t n = v_hi - v_lo;
    const int64_t blocks = ceil_div(n, MSTEP_WARPS);
    const int64_t heavy_blocks = ceil_div(n, MSTEP_WARPS * 32);
    DMX_REQUIRE(blocks < (1ll << 31), "grid too large");
    mstep_kernel<LPR, SLOTS, SQUARE, FULL><<<(unsigned)blocks, MSTEP_WARPS * 32, 0, stream>>>(
        offsets, cb, e, post, ld_post, G, power, addition, ld_add, addition64, ld_add64, v_lo, v_hi);
    DMX_LAUNCH_CHECK();
    mstep_heavy_kernel<LPR, SLOTS, SQUARE, FULL><<<(unsigned)heavy_blocks, MSTEP_WARPS * 32, 0, stream>>>(
        offsets, cb, e, post, ld_post, G, power, addition, ld_add, addition64, ld_add64, v_lo, v_hi);
    DMX_LAUNCH_CHECK();
    return 0;
}

template <bool SQUARE>
static int launch_mstep(cudaStream_t stream, const int64_t* offsets, const int32_t* cb, const float* e,
                        const float* post, int64_t ld_post, int G, float power, float* addition, int64_t ld_add,
                        double* addition64, int64_t ld_add64, int64_t v_lo, int64_t v_hi) {
    const int quads = (int)((ld_post + 3) / 4);
#define DMX_ARGS stream, offsets, cb, e, post, ld_post, G, power, addition, ld_add, addition64, ld_add64, v_lo, v_hi
#define DMX_SHAPE(LPR, SLOTS)                                                        \
    return quads == LPR * SLOTS ? launch_pair<LPR, SLOTS, SQUARE, true>(DMX_ARGS) \
                                : launch_pair<LPR, SLOTS, SQUARE, false>(DMX_ARGS)
    if (quads <= 8) DMX_SHAPE(8, 1);
    if (quads <= 16) DMX_SHAPE(16, 1);
    if (quads <= 32) DMX_SHAPE(32, 1);
    if (quads <= 64) DMX_SHAPE(32, 2);
    if (quads <= 128) DMX_SHAPE(32, 4);
#undef DMX_SHAPE
#undef DMX_ARGS
    set_error("M-step supports up to 512 genotypes");
    return -2;
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
    DMX_REQUIRE(ld_singlet % 4 == 0 && ld_singlet >= n_genotypes && ((uintptr_t)singlet_posteriors & 15) == 0,
                "singlet posteriors must be 16-byte aligned with a leading dimension that is a multiple of 4");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (power == 2.0)
        return launch_mstep<true>(stream, variant_offsets, csc_cb, csc_e, singlet_posteriors, ld_singlet, n_genotypes,
                                  2.f, addition, ld_addition, addition64, ld_addition64, variant_lo, variant_hi);
    return launch_mstep<false>(stream, variant_offsets, csc_cb, csc_e, singlet_posteriors, ld_singlet, n_genotypes,
                               (float)power, addition, ld_addition, addition64, ld_addition64, variant_lo, variant_hi);
}

}  // extern "C"
