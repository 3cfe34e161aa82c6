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
#include <stdlib.h>

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

// ---------------------------------------------------------------------------------------------------------------
// Planned M-step (dmx_mstep_plan): three tiers by rows per variant
//   light   (<= LIGHT_MAX rows, almost all variants): a GROUP of LPR lanes per variant, 32 / LPR variants side by side
//           in a warp; a warp owns 32 consecutive variants and its groups pull the next one as they finish (the row
//           counts of neighbouring variants differ a lot).  Rows are added strictly in order in float64 -- exactly
//           np.bincount's order -- and nothing is combined across lanes, so there is no shuffle reduction at all.
//   medium  (<= HEAVY_ROWS): one warp per variant from the plan's list (RowWalker, row groups combined in fixed order)
//   heavy   (> HEAVY_ROWS): cut into chunks of HEAVY_ROWS rows, one CTA per chunk, float64 partials to scratch, summed
//           in chunk order by a small second kernel -- no work item is longer than HEAVY_ROWS rows
// ---------------------------------------------------------------------------------------------------------------
constexpr int LIGHT_MAX = 128;
constexpr int LIGHT_VPW = 32;   // variants per warp in the light kernel
constexpr int LIGHT_WAVES = 4;  // rows in flight per group

struct MstepPlanLayout {
    int64_t cap_medium, cap_heavy_variants, cap_heavy_items;
    int64_t off_medium, off_hv, off_hv_first, off_hv_chunks, off_item_hv, off_item_chunk, total_ints;
};

static MstepPlanLayout mstep_plan_layout(int64_t n_rows) {
    MstepPlanLayout l;
    l.cap_medium = n_rows / LIGHT_MAX + 1;
    l.cap_heavy_variants = n_rows / HEAVY_ROWS + 1;
    l.cap_heavy_items = 2 * l.cap_heavy_variants + 1;
    int64_t off = 4;  // counters: n_medium, n_heavy_variants, n_heavy_items, pad
    l.off_medium = off; off += l.cap_medium;
    l.off_hv = off; off += l.cap_heavy_variants;
    l.off_hv_first = off; off += l.cap_heavy_variants;
    l.off_hv_chunks = off; off += l.cap_heavy_variants;
    l.off_item_hv = off; off += l.cap_heavy_items;
    l.off_item_chunk = off; off += l.cap_heavy_items;
    l.total_ints = off;
    return l;
}

// the order of the lists depends on the atomics, no result does (every variant is computed independently)
__global__ void mstep_plan_kernel(const int64_t* __restrict__ offsets, int64_t n_variants, int32_t* __restrict__ plan,
                                  MstepPlanLayout l) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_variants; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = offsets[v + 1] - offsets[v];
        if (n > HEAVY_ROWS) {
            const int chunks = (int)((n + HEAVY_ROWS - 1) / HEAVY_ROWS);
            const int h = atomicAdd(&plan[1], 1);
            const int first = atomicAdd(&plan[2], chunks);
            if (h < l.cap_heavy_variants && first + chunks <= l.cap_heavy_items) {
                plan[l.off_hv + h] = (int32_t)v;
                plan[l.off_hv_first + h] = first;
                plan[l.off_hv_chunks + h] = chunks;
                for (int c = 0; c < chunks; ++c) {
                    plan[l.off_item_hv + first + c] = h;
                    plan[l.off_item_chunk + first + c] = c;
                }
            }
        } else if (n > LIGHT_MAX) {
            const int m = atomicAdd(&plan[0], 1);
            if (m < l.cap_medium) plan[l.off_medium + m] = (int32_t)v;
        }
    }
}

template <int LPR, int SLOTS, bool SQUARE, bool FULL>
__device__ __forceinline__ void mstep_light_body(
    int64_t block, const int64_t* __restrict__ offsets, const int32_t* __restrict__ cb_arr,
    const float* __restrict__ e_arr, const float* __restrict__ post, int64_t ld_post, int n_genotypes, float power,
    float* __restrict__ addition, int64_t ld_add, double* __restrict__ addition64, int64_t ld_add64,
    int64_t variant_lo, int64_t variant_hi) {
    constexpr int NG = 32 / LPR;  // variants processed side by side in a warp
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR, grp = lane / LPR;
    const int64_t warp_id = (block * blockDim.x + threadIdx.x) >> 5;
    const int64_t v0 = variant_lo + warp_id * LIGHT_VPW;
    if (v0 >= variant_hi) return;
    const int count = (int)(variant_hi - v0 < LIGHT_VPW ? variant_hi - v0 : LIGHT_VPW);
    const int n_quads = (int)((ld_post + 3) / 4);
    // row ranges of the warp's variants: two coalesced loads, handed out with shuffles afterwards
    const int64_t o_lo = lane < count ? __ldg(offsets + v0 + lane) : 0;
    const int64_t o_hi = lane < count ? __ldg(offsets + v0 + lane + 1) : 0;

    double acc[SLOTS][4];
    int cur = grp, next = NG;
    int64_t row = 0, end = 0;
    bool have = false;
    auto take = [&]() {  // executed by all lanes: adopt variant `cur` (or nothing when the warp's range is exhausted)
        const int src = cur < count ? cur : 0;
        const int64_t lo = __shfl_sync(0xffffffffu, o_lo, src);
        const int64_t hi = __shfl_sync(0xffffffffu, o_hi, src);
        have = cur < count && hi - lo <= LIGHT_MAX;  // longer variants belong to the medium / heavy tiers
        row = have ? lo : 0;
        end = have ? hi : 0;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[s][c] = 0.0;
    };
    take();

    for (;;) {
        float4 x[LIGHT_WAVES][SLOTS];
        float w[LIGHT_WAVES];
        bool ok[LIGHT_WAVES];
#pragma unroll
        for (int u = 0; u < LIGHT_WAVES; ++u) {
            const int64_t r = row + u;
            ok[u] = r < end;
            const int32_t cb = ok[u] ? __ldg(cb_arr + r) : 0;
            w[u] = ok[u] ? __fsub_rn(1.f, __ldg(e_arr + r)) : 0.f;
            const float4* prow = reinterpret_cast<const float4*>(post + (int64_t)cb * ld_post);
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int q = sub + LPR * s;
                x[u][s] = (ok[u] && (FULL || q < n_quads)) ? __ldg(prow + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < LIGHT_WAVES; ++u)
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const float v[4] = {x[u][s].x, x[u][s].y, x[u][s].z, x[u][s].w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float t = __fmul_rn(v[c], w[u]);
                    if (SQUARE) acc[s][c] += (double)__fmul_rn(t, t);  // rows past the end: t = 0, adds exactly 0
                    else acc[s][c] += (double)(ok[u] ? powf(t, power) : 0.f);
                }
            }
        row += LIGHT_WAVES;
        const bool done = row >= end;
        if (done && have) {
            const int64_t v = v0 + cur;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int g = 4 * (sub + LPR * s);
                if (FULL || g < n_genotypes) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (FULL || g + c < n_genotypes) {
                            if (addition) addition[v * ld_add + g + c] = (float)acc[s][c];
                            if (addition64) addition64[v * ld_add64 + g + c] = acc[s][c];
                        }
                }
            }
        }
        // groups that finished pull the next variants of the warp's range, in group order
        const unsigned finished = __ballot_sync(0xffffffffu, done && sub == 0);
        if (done) {
            cur = next + __popc(finished & ((1u << (grp * LPR)) - 1u));
        }
        next += __popc(finished);
        const bool all_done = __all_sync(0xffffffffu, done && cur >= count);
        if (all_done) break;
        // take() shuffles: all lanes execute it, only finished groups adopt the result
        {
            const int src = cur < count ? cur : 0;
            const int64_t lo = __shfl_sync(0xffffffffu, o_lo, src);
            const int64_t hi = __shfl_sync(0xffffffffu, o_hi, src);
            if (done) {
                have = cur < count && hi - lo <= LIGHT_MAX;
                row = have ? lo : 0;
                end = have ? hi : 0;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[s][c] = 0.0;
            }
        }
    }
}

// medium tier: one warp per listed variant
template <int LPR, int SLOTS, bool SQUARE, bool FULL>
__device__ __forceinline__ void mstep_medium_body(
    int64_t block, const int32_t* __restrict__ list, int n_list, const int64_t* __restrict__ offsets,
    const int32_t* __restrict__ cb_arr, const float* __restrict__ e_arr, const float* __restrict__ post,
    int64_t ld_post, int n_genotypes, float power, float* __restrict__ addition, int64_t ld_add,
    double* __restrict__ addition64, int64_t ld_add64, int64_t variant_lo, int64_t variant_hi) {
    const int lane = threadIdx.x & 31;
    const int64_t k = (block * blockDim.x + threadIdx.x) >> 5;
    if (k >= n_list) return;
    const int64_t v = list[k];
    if (v < variant_lo || v >= variant_hi) return;
    const int64_t lo = offsets[v], hi = offsets[v + 1];
    const int n_quads = (int)((ld_post + 3) / 4);
    RowWalker<LPR, SLOTS, SQUARE, FULL> walker;
    walker.init(lane);
    for (int64_t base = lo; base < hi; base += 32)
        walker.batch(cb_arr, e_arr, base, (int)(hi - base < 32 ? hi - base : 32), post, ld_post, n_quads, power);
    walker.combine();
    walker.store(v, n_genotypes, addition, ld_add, addition64, ld_add64);
}

// heavy tier: one CTA per (variant, chunk of HEAVY_ROWS rows); float64 partial sums [n_items, G]
template <int LPR, int SLOTS, bool SQUARE, bool FULL>
__device__ __forceinline__ void mstep_heavy_chunk_body(
    int item, double (*partial)[SLOTS * 4][32], const int32_t* __restrict__ hv, const int32_t* __restrict__ item_hv,
    const int32_t* __restrict__ item_chunk, const int64_t* __restrict__ offsets, const int32_t* __restrict__ cb_arr,
    const float* __restrict__ e_arr, const float* __restrict__ post, int64_t ld_post, int n_genotypes, float power,
    double* __restrict__ scratch, int64_t variant_lo, int64_t variant_hi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t v = hv[item_hv[item]];
    if (v < variant_lo || v >= variant_hi) return;
    const int n_quads = (int)((ld_post + 3) / 4);
    const int64_t lo = offsets[v] + (int64_t)item_chunk[item] * HEAVY_ROWS;
    const int64_t hi = offsets[v + 1] < lo + HEAVY_ROWS ? offsets[v + 1] : lo + HEAVY_ROWS;
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
    if (warp == 0 && walker.rgw == 0) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int g = 4 * (walker.sub + LPR * s) + c;
                if (g < n_genotypes) {
                    double sum = partial[0][s * 4 + c][lane];
                    for (int w = 1; w < MSTEP_WARPS; ++w) sum += partial[w][s * 4 + c][lane];
                    scratch[(int64_t)item * n_genotypes + g] = sum;
                }
            }
    }
}

__global__ void mstep_heavy_combine_kernel(const int32_t* __restrict__ hv, const int32_t* __restrict__ hv_first,
                                           const int32_t* __restrict__ hv_chunks, int n_heavy, int n_genotypes,
                                           const double* __restrict__ scratch, float* __restrict__ addition,
                                           int64_t ld_add, double* __restrict__ addition64, int64_t ld_add64,
                                           int64_t variant_lo, int64_t variant_hi) {
    const int h = blockIdx.x;
    if (h >= n_heavy) return;
    const int64_t v = hv[h];
    if (v < variant_lo || v >= variant_hi) return;
    const int first = hv_first[h], chunks = hv_chunks[h];
    for (int g = threadIdx.x; g < n_genotypes; g += blockDim.x) {
        double sum = 0.0;
        for (int c = 0; c < chunks; ++c) sum += scratch[(int64_t)(first + c) * n_genotypes + g];  // chunk order
        if (addition) addition[v * ld_add + g] = (float)sum;
        if (addition64) addition64[v * ld_add64 + g] = sum;
    }
}

struct PlannedArgs {
    const int32_t* plan;
    MstepPlanLayout layout;
    int n_medium, n_heavy_variants, n_heavy_items;
    double* scratch;
};

// One launch for all tiers: the longest work items first (heavy chunks, then medium variants), the light variants
// after them, so the tiers fill each other's stalls instead of running back to back (each alone is latency bound).
template <int LPR, int SLOTS, bool SQUARE, bool FULL>
__global__ void __launch_bounds__(MSTEP_WARPS * 32) mstep_tiers_kernel(
    const int32_t* __restrict__ plan, MstepPlanLayout l, int n_medium, int n_heavy_items, int medium_blocks,
    const int64_t* __restrict__ offsets, const int32_t* __restrict__ cb_arr, const float* __restrict__ e_arr,
    const float* __restrict__ post, int64_t ld_post, int n_genotypes, float power, float* __restrict__ addition,
    int64_t ld_add, double* __restrict__ addition64, int64_t ld_add64, double* __restrict__ scratch,
    int64_t variant_lo, int64_t variant_hi) {
    __shared__ double partial[MSTEP_WARPS][SLOTS * 4][32];
    const int64_t b = blockIdx.x;
    if (b < n_heavy_items) {
        mstep_heavy_chunk_body<LPR, SLOTS, SQUARE, FULL>((int)b, partial, plan + l.off_hv, plan + l.off_item_hv,
                                                         plan + l.off_item_chunk, offsets, cb_arr, e_arr, post, ld_post,
                                                         n_genotypes, power, scratch, variant_lo, variant_hi);
    } else if (b < n_heavy_items + medium_blocks) {
        mstep_medium_body<LPR, SLOTS, SQUARE, FULL>(b - n_heavy_items, plan + l.off_medium, n_medium, offsets, cb_arr,
                                                    e_arr, post, ld_post, n_genotypes, power, addition, ld_add, addition64,
                                                    ld_add64, variant_lo, variant_hi);
    } else {
        mstep_light_body<LPR, SLOTS, SQUARE, FULL>(b - n_heavy_items - medium_blocks, offsets, cb_arr, e_arr, post, ld_post,
                                                   n_genotypes, power, addition, ld_add, addition64, ld_add64, variant_lo,
                                                   variant_hi);
    }
}

template <int LPR, int SLOTS, bool SQUARE, bool FULL>
static int launch_planned(cudaStream_t stream, const int64_t* offsets, const int32_t* cb, const float* e,
                          const float* post, int64_t ld_post, int G, float power, float* addition, int64_t ld_add,
                          double* addition64, int64_t ld_add64, int64_t v_lo, int64_t v_hi, const PlannedArgs& pa) {
    const int threads = MSTEP_WARPS * 32;
    const int64_t n = v_hi - v_lo;
    const int64_t light_blocks = ceil_div(ceil_div(n, LIGHT_VPW), MSTEP_WARPS);
    const int64_t medium_blocks = ceil_div(pa.n_medium, MSTEP_WARPS);
    const int64_t blocks = pa.n_heavy_items + medium_blocks + light_blocks;
    DMX_REQUIRE(blocks < (1ll << 31), "grid too large");
    const MstepPlanLayout& l = pa.layout;
    mstep_tiers_kernel<LPR, SLOTS, SQUARE, FULL><<<(unsigned)blocks, threads, 0, stream>>>(
        pa.plan, l, pa.n_medium, pa.n_heavy_items, (int)medium_blocks, offsets, cb, e, post, ld_post, G, power, addition,
        ld_add, addition64, ld_add64, pa.scratch, v_lo, v_hi);
    DMX_LAUNCH_CHECK();
    if (pa.n_heavy_items > 0) {
        mstep_heavy_combine_kernel<<<(unsigned)pa.n_heavy_variants, 64, 0, stream>>>(
            pa.plan + l.off_hv, pa.plan + l.off_hv_first, pa.plan + l.off_hv_chunks, pa.n_heavy_variants, G, pa.scratch,
            addition, ld_add, addition64, ld_add64, v_lo, v_hi);
        DMX_LAUNCH_CHECK();
    }
    return 0;
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
                        double* addition64, int64_t ld_add64, int64_t v_lo, int64_t v_hi,
                        const PlannedArgs* pa = nullptr) {
    const int quads = (int)((ld_post + 3) / 4);
#define DMX_ARGS stream, offsets, cb, e, post, ld_post, G, power, addition, ld_add, addition64, ld_add64, v_lo, v_hi
#define DMX_SHAPE(LPR, SLOTS)                                                                              \
    {                                                                                                      \
        if (pa)                                                                                            \
            return quads == LPR * SLOTS && G % 4 == 0 ? launch_planned<LPR, SLOTS, SQUARE, true>(DMX_ARGS, *pa)   \
                                                      : launch_planned<LPR, SLOTS, SQUARE, false>(DMX_ARGS, *pa); \
        return quads == LPR * SLOTS ? launch_pair<LPR, SLOTS, SQUARE, true>(DMX_ARGS)                      \
                                    : launch_pair<LPR, SLOTS, SQUARE, false>(DMX_ARGS);                    \
    }
    // few genotypes: fewer lanes per row, more rows (light tier: more variants) side by side in a warp -- with 8 lanes
    // per row a 4-genotype M-step kept 7 of 8 lanes idle.  DMX_MSTEP_MIN_LPR=8 restores that for comparison.
    const char* min_lpr_env = getenv("DMX_MSTEP_MIN_LPR");
    const int min_lpr = (min_lpr_env && *min_lpr_env) ? atoi(min_lpr_env) : 1;
    if (quads <= 1 && min_lpr <= 1) DMX_SHAPE(1, 1);
    if (quads <= 2 && min_lpr <= 2) DMX_SHAPE(2, 1);
    if (quads <= 4 && min_lpr <= 4) DMX_SHAPE(4, 1);
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

int64_t dmx_mstep_plan_bytes(int64_t n_rows) {
    return (int64_t)sizeof(int32_t) * dmx::mstep_plan_layout(n_rows > 0 ? n_rows : 0).total_ints;
}

int dmx_mstep_plan(const int64_t* variant_offsets, int64_t n_variants, int64_t n_rows, void* plan, int64_t plan_bytes,
                   int64_t* h_counts, void* stream_) {
    using namespace dmx;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (h_counts) h_counts[0] = h_counts[1] = h_counts[2] = 0;
    const MstepPlanLayout l = mstep_plan_layout(n_rows > 0 ? n_rows : 0);
    DMX_REQUIRE(plan && plan_bytes >= (int64_t)sizeof(int32_t) * l.total_ints, "M-step plan buffer too small");
    DMX_CUDA(cudaMemsetAsync(plan, 0, 4 * sizeof(int32_t), stream));
    if (n_variants > 0) {
        int64_t blocks = ceil_div(n_variants, 256);
        if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
        mstep_plan_kernel<<<(int)blocks, 256, 0, stream>>>(variant_offsets, n_variants, (int32_t*)plan, l);
        DMX_LAUNCH_CHECK();
    }
    int32_t counts[4] = {0, 0, 0, 0};
    DMX_CUDA(cudaMemcpyAsync(counts, plan, sizeof(counts), cudaMemcpyDeviceToHost, stream));
    DMX_CUDA(cudaStreamSynchronize(stream));
    DMX_REQUIRE(counts[0] <= l.cap_medium && counts[1] <= l.cap_heavy_variants && counts[2] <= l.cap_heavy_items,
                "internal: M-step plan capacities exceeded (%d, %d, %d)", counts[0], counts[1], counts[2]);
    if (h_counts) { h_counts[0] = counts[0]; h_counts[1] = counts[1]; h_counts[2] = counts[2]; }
    return 0;
}

int dmx_mstep_planned(const int64_t* variant_offsets, const int32_t* csc_cb, const float* csc_e,
                      const float* singlet_posteriors, int64_t ld_singlet, int32_t n_genotypes, double power,
                      float* addition, int64_t ld_addition, double* addition64, int64_t ld_addition64,
                      int64_t variant_lo, int64_t variant_hi, const void* plan, int64_t n_rows, int64_t n_medium,
                      int64_t n_heavy_variants, int64_t n_heavy_items, double* heavy_scratch, void* stream_) {
    using namespace dmx;
    if (variant_hi <= variant_lo || n_genotypes <= 0) return 0;
    DMX_REQUIRE(addition || addition64, "no output buffer");
    DMX_REQUIRE(ld_singlet % 4 == 0 && ld_singlet >= n_genotypes && ((uintptr_t)singlet_posteriors & 15) == 0,
                "singlet posteriors must be 16-byte aligned with a leading dimension that is a multiple of 4");
    DMX_REQUIRE(plan, "no M-step plan");
    DMX_REQUIRE(n_heavy_items == 0 || heavy_scratch, "heavy variants need the float64 scratch [n_heavy_items, G]");
    PlannedArgs pa;
    pa.plan = (const int32_t*)plan;
    pa.layout = mstep_plan_layout(n_rows > 0 ? n_rows : 0);
    pa.n_medium = (int)n_medium;
    pa.n_heavy_variants = (int)n_heavy_variants;
    pa.n_heavy_items = (int)n_heavy_items;
    pa.scratch = heavy_scratch;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (power == 2.0)
        return launch_mstep<true>(stream, variant_offsets, csc_cb, csc_e, singlet_posteriors, ld_singlet, n_genotypes,
                                  2.f, addition, ld_addition, addition64, ld_addition64, variant_lo, variant_hi, &pa);
    return launch_mstep<false>(stream, variant_offsets, csc_cb, csc_e, singlet_posteriors, ld_singlet, n_genotypes,
                               (float)power, addition, ld_addition, addition64, ld_addition64, variant_lo, variant_hi, &pa);
}

}  // extern "C"
