// Per-(barcode, SNP) regularised likelihood: the `aggregate_on_snps = True` branch of compute_barcode_logits
// (demux.py:204-244) on the device.
//
// The reference groups the matched molecule-level calls by (compressed_cb, snp_id) (FeatureLookup, utils.py:207-265),
// sums float32 logs per group and column in float64, divides by count ** compensation, runs a float32 log_softmax over
// the columns, mixes every column with a uniform "bad SNP" mass (np.logaddexp against a float64 scalar: float64 from
// here on), runs a float64 log_softmax and sums the groups of a barcode in ascending (barcode, SNP) order.  The
// doublet penalties only fix the number of columns there (demux.py:212) and are never added; the same holds here.
//
//   dmx_build_snp_groups   stable radix sort of (cb * n_snps + snp) keys (CUB, library code as in builder.cu) and the
//                          kernels below: calls in group order, group offsets, groups per barcode;
//   dmx_snp_logits         one CTA per barcode: its warps take the barcode's groups round-robin, lanes stride the
//                          columns, the reductions over columns are warp shuffles; every warp sums its groups in order
//                          and the CTA adds the warps' partial rows in warp order: no atomics, deterministic;
//   dmx_softmax_rows_f64   float64 row softmax (+ the optional prior logits of the first EM iteration).
//
// Arithmetic follows the reference's dtypes step by step; the float32 sum of exponentials is taken lane-strided
// instead of numpy's pairwise order and logf / expf / exp / log1p are CUDA's, so results agree to tolerance
// (tests/test_gpu_snp_aggregate.py), not bit for bit.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace dmx {

// ---------------------------------------------------------------------------------------------------------
// groups
// ---------------------------------------------------------------------------------------------------------

__global__ void snp_keys_kernel(const int32_t* __restrict__ call_variant, const int32_t* __restrict__ call_cb,
                                const int32_t* __restrict__ variant2snp, int64_t n_calls, int64_t n_snps,
                                int64_t cb_lo, int64_t cb_hi, uint64_t sentinel, uint64_t* __restrict__ keys,
                                uint32_t* __restrict__ idx) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_calls; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = call_variant[k];
        const int64_t cb = call_cb[k];
        uint64_t key = sentinel;  // unmatched calls (demux.py:345-347) and barcodes of other shards sort behind the rest
        if (v >= 0 && cb >= cb_lo && cb < cb_hi) key = (uint64_t)cb * (uint64_t)n_snps + (uint64_t)variant2snp[v];
        keys[k] = key;
        idx[k] = (uint32_t)k;
    }
}

__global__ void snp_head_flags_kernel(const uint64_t* __restrict__ keys_sorted, int64_t n, int32_t* __restrict__ flags) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        flags[k] = (k == 0 || keys_sorted[k] != keys_sorted[k - 1]) ? 1 : 0;
}

__global__ void snp_emit_kernel(const uint64_t* __restrict__ keys_sorted, const uint32_t* __restrict__ idx_sorted,
                                const int32_t* __restrict__ incl, int64_t n, uint64_t sentinel,
                                const int32_t* __restrict__ call_variant, const float* __restrict__ call_e,
                                int32_t* __restrict__ grouped_variant, float* __restrict__ grouped_e,
                                int64_t* __restrict__ group_offsets) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys_sorted[k];
        if (key != sentinel) {
            const uint32_t src = idx_sorted[k];
            grouped_variant[k] = call_variant[src];
            grouped_e[k] = call_e[src];
        }
        if (k == 0 || key != keys_sorted[k - 1]) group_offsets[incl[k] - 1] = k;  // the sentinel run closes the last group
    }
}

// thread b <= n_barcodes: number of groups of barcodes < b; thread n_barcodes also writes the totals
__global__ void snp_barcode_offsets_kernel(const uint64_t* __restrict__ keys_sorted, const int32_t* __restrict__ incl,
                                           int64_t n, int64_t n_snps, int64_t n_barcodes,
                                           int64_t* __restrict__ barcode_group_offsets,
                                           int64_t* __restrict__ group_offsets, int64_t* __restrict__ counters) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b > n_barcodes) return;
    const uint64_t target = (uint64_t)b * (uint64_t)n_snps;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys_sorted[mid] < target) lo = mid + 1; else hi = mid;
    }
    const int64_t group = lo < n ? (int64_t)incl[lo] - 1 : (int64_t)incl[n - 1];
    barcode_group_offsets[b] = group;
    if (b == n_barcodes) {  // target == sentinel: lo = number of matched calls, group = number of groups
        group_offsets[group] = lo;
        counters[0] = lo;
        counters[1] = group;
    }
}

// ---------------------------------------------------------------------------------------------------------
// logits
// ---------------------------------------------------------------------------------------------------------

// column -> (i, j) packed as i | j << 16, order of demux.py:175-191
__global__ void snp_pairs_kernel(int n_genotypes, int n_cols, int32_t* __restrict__ pairs) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cols; c += gridDim.x * blockDim.x) {
        int i = c, j = c;
        if (c >= n_genotypes) {
            int d = c - n_genotypes;
            i = 0;
            while (d >= n_genotypes - 1 - i) { d -= n_genotypes - 1 - i; ++i; }
            j = i + 1 + d;
        }
        pairs[c] = i | (j << 16);
    }
}

__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// np.logaddexp (numpy/_core/src/npymath/npy_math_internal.h.src, npy_logaddexp)
__device__ __forceinline__ double np_logaddexp(double x, double y) {
    if (x == y) return x + 0.693147180559945309417232121458176568;
    const double d = x - y;
    if (d > 0) return x + log1p(exp(-d));
    if (d <= 0) return y + log1p(exp(d));
    return d;  // NaN
}

constexpr int SNP_WARPS_PER_CTA = 8;

__global__ void __launch_bounds__(32 * SNP_WARPS_PER_CTA) snp_logits_kernel(
    const int64_t* __restrict__ barcode_group_offsets, const int64_t* __restrict__ group_offsets,
    const int32_t* __restrict__ grouped_variant, const float* __restrict__ grouped_e, int64_t n_barcodes,
    const float* __restrict__ table, int64_t ld_table, int n_cols, const int32_t* __restrict__ pairs, double log_bad,
    double compensation, double* __restrict__ logits, int64_t ld_logits, float* scratch32, double* scratch64,
    double* scratch_part, int64_t scratch_ld) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t slot = blockIdx.x * (int64_t)SNP_WARPS_PER_CTA + w;
    // a lane only ever reads back the xs / ts entries it wrote itself (columns lane, lane + 32, ...); the partial rows
    // are read by the whole CTA after a barrier
    float* xs = scratch32 + slot * scratch_ld;
    double* ts = scratch64 + slot * scratch_ld;
    double* part = scratch_part + slot * scratch_ld;
    const double* cta_parts = scratch_part + blockIdx.x * (int64_t)SNP_WARPS_PER_CTA * scratch_ld;
    for (int64_t b = blockIdx.x; b < n_barcodes; b += gridDim.x) {
        for (int c = lane; c < n_cols; c += 32) part[c] = 0.0;
        const int64_t q_lo = barcode_group_offsets[b], q_hi = barcode_group_offsets[b + 1];
        for (int64_t q = q_lo + w; q < q_hi; q += SNP_WARPS_PER_CTA) {
            const int64_t lo = group_offsets[q], hi = group_offsets[q + 1];
            // counts ** compensation (demux.py:231); 0.5 is the reference's constant and sqrt is correctly rounded
            const double denom = compensation == 0.5 ? sqrt((double)(hi - lo)) : pow((double)(hi - lo), compensation);
            float m = -INFINITY;
            for (int c = lane; c < n_cols; c += 32) {
                const int32_t pr = __ldg(pairs + c);
                const int i = pr & 0xffff, j = pr >> 16;
                double sum = 0.0;  // np.bincount: float64, calls in their original order (the sort is stable)
                for (int64_t k = lo; k < hi; ++k) {
                    const float* row = table + (int64_t)__ldg(grouped_variant + k) * ld_table;
                    const float p = i == j ? __ldg(row + i) : __fmul_rn(__fadd_rn(__ldg(row + i), __ldg(row + j)), 0.5f);
                    sum += (double)logf(__fadd_rn(p, __ldg(grouped_e + k)));  // demux.py:227-228
                }
                const float x = (float)((double)(float)sum / denom);  // float32 column, float64 quotient, float32 store
                xs[c] = x;
                m = fmaxf(m, x);
            }
            m = warp_max(m);
            if (!isfinite(m)) m = 0.f;  // scipy log_softmax
            float s = 0.f;
            for (int c = lane; c < n_cols; c += 32) s += expf(__fsub_rn(xs[c], m));
            const float l = logf(warp_sum_f32(s));
            // second log_softmax: its maximum is the image of the first one's maximum, (m - m) - l
            const double z_max = np_logaddexp((double)__fsub_rn(0.f, l), log_bad);
            double s2 = 0.0;
            for (int c = lane; c < n_cols; c += 32) {
                const float y = __fsub_rn(__fsub_rn(xs[c], m), l);
                const double t = np_logaddexp((double)y, log_bad) - z_max;
                ts[c] = t;
                s2 += exp(t);
            }
            const double l2 = log(warp_sum(s2));
            for (int c = lane; c < n_cols; c += 32) part[c] += ts[c] - l2;
        }
        __syncthreads();  // the partial rows of all warps are complete and visible to the CTA
        for (int c = threadIdx.x; c < n_cols; c += 32 * SNP_WARPS_PER_CTA) {
            double total = 0.0;
#pragma unroll
            for (int k = 0; k < SNP_WARPS_PER_CTA; ++k) total += cta_parts[k * scratch_ld + c];
            logits[b * ld_logits + c] = total;
        }
        __syncthreads();  // before the next barcode zeroes the partial rows
    }
}

// scipy.special.softmax(x, axis=-1) in float64 (demux.py:101,152 on the float64 logits of this branch); a warp per row
__global__ void __launch_bounds__(256) softmax_rows_f64_kernel(
    double* __restrict__ logits, int64_t ld_logits, const double* __restrict__ prior, int64_t ld_prior, int64_t n_rows,
    int n_cols, double* __restrict__ post, int64_t ld_post, float* __restrict__ singlets, int64_t ld_singlet,
    int n_singlets) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = gridDim.x * (int64_t)(blockDim.x >> 5);
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        double* x = logits + r * ld_logits;
        double m = -INFINITY;
        for (int c = lane; c < n_cols; c += 32) {
            double v = x[c];
            if (prior) { v += prior[r * ld_prior + c]; x[c] = v; }  // demux.py:97-99, in place like the reference
            m = fmax(m, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (!isfinite(m)) m = 0.0;
        double s = 0.0;
        for (int c = lane; c < n_cols; c += 32) s += exp(x[c] - m);
        s = warp_sum(s);
        for (int c = lane; c < n_cols; c += 32) {
            const double p = exp(x[c] - m) / s;
            if (post) post[r * ld_post + c] = p;
            if (singlets && c < n_singlets) singlets[r * ld_singlet + c] = (float)p;
        }
    }
}

static int snp_key_bits(int64_t n_snps, int64_t n_barcodes) {
    const uint64_t sentinel = (uint64_t)n_barcodes * (uint64_t)n_snps;
    int bits = 1;
    while (bits < 64 && (sentinel >> bits) != 0) ++bits;
    return bits;
}

static int64_t snp_cub_bytes(int64_t n) {
    size_t sort = 0, scan = 0;
    if (cub::DeviceRadixSort::SortPairs(nullptr, sort, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                        (const uint32_t*)nullptr, (uint32_t*)nullptr, n) != cudaSuccess) return -1;
    if (cub::DeviceScan::InclusiveSum(nullptr, scan, (const int32_t*)nullptr, (int32_t*)nullptr, n) != cudaSuccess) return -1;
    return (int64_t)(sort > scan ? sort : scan);
}

static inline int snp_grid(int64_t n, int threads) {
    int64_t blocks = ceil_div(n > 0 ? n : 1, threads);
    const int64_t cap = (int64_t)sm_count() * 32;
    return (int)(blocks < cap ? blocks : cap);
}

// persistent CTAs: one per barcode, at most as many as are resident at once (a CTA of a second wave would only start
// once a first-wave CTA has finished all of its barcodes)
static inline int64_t snp_logits_ctas(int64_t n_barcodes) {
    static int per_sm = 0;
    if (per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, snp_logits_kernel, 32 * SNP_WARPS_PER_CTA, 0) != cudaSuccess ||
            n <= 0)
            n = 2;
        per_sm = n;
    }
    const int64_t cap = (int64_t)sm_count() * per_sm;
    const int64_t want = n_barcodes > 0 ? n_barcodes : 1;
    return want < cap ? want : cap;
}

}  // namespace dmx

extern "C" {

int64_t dmx_snp_groups_workspace_bytes(int64_t n_calls) {
    using namespace dmx;
    if (n_calls < 0 || n_calls >= ((int64_t)1 << 31)) {
        set_error("n_calls %lld out of range", (long long)n_calls);
        return -1;
    }
    const int64_t n = n_calls > 0 ? n_calls : 1;
    const int64_t cub_bytes = snp_cub_bytes(n);
    if (cub_bytes < 0) {
        set_error("cub workspace query failed");
        return -1;
    }
    // keys a/b (8 B), idx a/b (4 B), flags -> inclusive sums (4 B), counters, cub scratch; every block 256-byte aligned
    return 2 * round_up(8 * n, 256) + 3 * round_up(4 * n, 256) + 256 + round_up(cub_bytes, 256);
}

int dmx_build_snp_groups(const int32_t* call_variant, const int32_t* call_cb, const float* call_e, int64_t n_calls,
                         const int32_t* variant2snp, int64_t n_snps, int64_t n_barcodes, int64_t cb_lo, int64_t cb_hi,
                         void* workspace, int64_t workspace_bytes, int32_t* grouped_variant, float* grouped_e,
                         int64_t* group_offsets, int64_t* barcode_group_offsets, int64_t* h_n_matched,
                         int64_t* h_n_groups, void* stream_) {
    using namespace dmx;
    cudaStream_t stream = (cudaStream_t)stream_;
    DMX_REQUIRE(h_n_matched && h_n_groups, "null host pointer");
    DMX_REQUIRE(n_barcodes >= 0 && n_snps >= 0, "negative sizes");
    *h_n_matched = *h_n_groups = 0;
    if (n_calls <= 0 || n_barcodes == 0) {
        DMX_CUDA(cudaMemsetAsync(barcode_group_offsets, 0, sizeof(int64_t) * (n_barcodes + 1), stream));
        DMX_CUDA(cudaMemsetAsync(group_offsets, 0, sizeof(int64_t), stream));
        return 0;
    }
    DMX_REQUIRE(n_snps > 0, "calls without SNPs");
    DMX_REQUIRE((double)n_barcodes * (double)n_snps < 9.0e18, "barcode x SNP key does not fit 63 bits");
    const int64_t need = dmx_snp_groups_workspace_bytes(n_calls);
    if (need < 0) return -1;
    DMX_REQUIRE(workspace && workspace_bytes >= need, "workspace too small: %lld < %lld", (long long)workspace_bytes,
                (long long)need);
    uint8_t* ws = (uint8_t*)workspace;
    uint64_t* keys_a = (uint64_t*)ws; ws += round_up(8 * n_calls, 256);
    uint64_t* keys_b = (uint64_t*)ws; ws += round_up(8 * n_calls, 256);
    uint32_t* idx_a = (uint32_t*)ws; ws += round_up(4 * n_calls, 256);
    uint32_t* idx_b = (uint32_t*)ws; ws += round_up(4 * n_calls, 256);
    int32_t* incl = (int32_t*)ws; ws += round_up(4 * n_calls, 256);
    int64_t* counters = (int64_t*)ws; ws += 256;
    void* cub_temp = ws;
    size_t cub_bytes = (size_t)(workspace_bytes - (ws - (uint8_t*)workspace));

    const uint64_t sentinel = (uint64_t)n_barcodes * (uint64_t)n_snps;
    const int threads = 256;
    snp_keys_kernel<<<snp_grid(n_calls, threads), threads, 0, stream>>>(call_variant, call_cb, variant2snp, n_calls,
                                                                        n_snps, cb_lo, cb_hi, sentinel, keys_a, idx_a);
    DMX_LAUNCH_CHECK();
    DMX_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, (const uint64_t*)keys_a, keys_b,
                                             (const uint32_t*)idx_a, idx_b, n_calls, 0,
                                             snp_key_bits(n_snps, n_barcodes), stream));
    snp_head_flags_kernel<<<snp_grid(n_calls, threads), threads, 0, stream>>>(keys_b, n_calls, incl);
    DMX_LAUNCH_CHECK();
    DMX_CUDA(cub::DeviceScan::InclusiveSum(cub_temp, cub_bytes, (const int32_t*)incl, incl, n_calls, stream));
    snp_emit_kernel<<<snp_grid(n_calls, threads), threads, 0, stream>>>(keys_b, idx_b, incl, n_calls, sentinel,
                                                                        call_variant, call_e, grouped_variant,
                                                                        grouped_e, group_offsets);
    DMX_LAUNCH_CHECK();
    snp_barcode_offsets_kernel<<<(int)ceil_div(n_barcodes + 1, threads), threads, 0, stream>>>(
        keys_b, incl, n_calls, n_snps, n_barcodes, barcode_group_offsets, group_offsets, counters);
    DMX_LAUNCH_CHECK();
    int64_t host_counters[2] = {0, 0};
    DMX_CUDA(cudaMemcpyAsync(host_counters, counters, sizeof(host_counters), cudaMemcpyDeviceToHost, stream));
    DMX_CUDA(cudaStreamSynchronize(stream));
    *h_n_matched = host_counters[0];
    *h_n_groups = host_counters[1];
    return 0;
}

int64_t dmx_snp_logits_workspace_bytes(int64_t n_barcodes, int32_t n_cols) {
    using namespace dmx;
    if (n_cols <= 0) return 0;
    const int64_t ld = round_up(n_cols, 32);
    // pair table + per warp: float32 staging row, float64 staging row, float64 partial sums
    return round_up(4 * (int64_t)n_cols, 256) +
           snp_logits_ctas(n_barcodes) * SNP_WARPS_PER_CTA * ld * (int64_t)(sizeof(float) + 2 * sizeof(double));
}

int dmx_snp_logits(const int64_t* barcode_group_offsets, const int64_t* group_offsets, const int32_t* grouped_variant,
                   const float* grouped_e, int64_t n_barcodes, const float* table, int64_t ld_table,
                   int32_t n_genotypes, double doublet_prior, double compensation, double* logits, int64_t ld_logits,
                   void* workspace, int64_t workspace_bytes, void* stream_) {
    using namespace dmx;
    if (n_barcodes <= 0 || n_genotypes <= 0) return 0;
    DMX_REQUIRE(doublet_prior >= 0 && doublet_prior < 1, "doublet_prior %g outside [0, 1)", doublet_prior);
    DMX_REQUIRE(n_genotypes < 32768, "too many genotypes");
    const int64_t n_cols64 = doublet_prior == 0 ? n_genotypes : (int64_t)n_genotypes * (n_genotypes + 1) / 2;
    DMX_REQUIRE(n_cols64 < ((int64_t)1 << 31), "too many columns");
    const int n_cols = (int)n_cols64;
    DMX_REQUIRE(ld_logits >= n_cols && ld_table >= n_genotypes, "leading dimensions too small");
    const int64_t need = dmx_snp_logits_workspace_bytes(n_barcodes, n_cols);
    DMX_REQUIRE(workspace && workspace_bytes >= need, "workspace too small: %lld < %lld", (long long)workspace_bytes,
                (long long)need);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t ld = round_up(n_cols, 32);
    const int64_t n_ctas = snp_logits_ctas(n_barcodes);
    const int64_t n_warps = n_ctas * SNP_WARPS_PER_CTA;
    uint8_t* ws = (uint8_t*)workspace;
    int32_t* pairs = (int32_t*)ws; ws += round_up(4 * (int64_t)n_cols, 256);
    double* scratch64 = (double*)ws; ws += n_warps * ld * (int64_t)sizeof(double);
    double* scratch_part = (double*)ws; ws += n_warps * ld * (int64_t)sizeof(double);
    float* scratch32 = (float*)ws;
    snp_pairs_kernel<<<(int)ceil_div(n_cols, 256), 256, 0, stream>>>(n_genotypes, n_cols, pairs);
    DMX_LAUNCH_CHECK();
    const double log_bad = log(0.01 / (double)n_cols);  // demux.py:234-235: np.log(p_bad_snp / len(column_names))
    snp_logits_kernel<<<(int)n_ctas, 32 * SNP_WARPS_PER_CTA, 0, stream>>>(
        barcode_group_offsets, group_offsets, grouped_variant, grouped_e, n_barcodes, table, ld_table, n_cols, pairs,
        log_bad, compensation, logits, ld_logits, scratch32, scratch64, scratch_part, ld);
    DMX_LAUNCH_CHECK();
    return 0;
}

int dmx_softmax_rows_f64(double* logits, int64_t ld_logits, const double* prior_logits, int64_t ld_prior, int64_t n_rows,
                         int32_t n_cols, double* posteriors, int64_t ld_post, float* singlet_posteriors,
                         int64_t ld_singlet, int32_t n_singlets, void* stream) {
    using namespace dmx;
    if (n_rows <= 0 || n_cols <= 0) return 0;
    DMX_REQUIRE(ld_logits >= n_cols, "ld_logits too small");
    DMX_REQUIRE(!posteriors || ld_post >= n_cols, "ld_post too small");
    DMX_REQUIRE(!singlet_posteriors || (n_singlets <= n_cols && ld_singlet >= n_singlets), "bad singlet layout");
    const int threads = 256;
    int64_t blocks = ceil_div(n_rows, threads / 32);
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    softmax_rows_f64_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(
        logits, ld_logits, prior_logits, ld_prior, n_rows, n_cols, posteriors, ld_post, singlet_posteriors, ld_singlet,
        n_singlets);
    DMX_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
