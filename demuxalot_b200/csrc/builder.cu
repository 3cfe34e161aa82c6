// Device row builder: packed count_snps records -> matched molecule calls -> (variant, barcode) rows
// in the reference's order (CSC, variant-major) and a barcode-major copy (CSR) for the E-step.
//
// Restates demux.py:334-363 (matching) and demux.py:276-300 (grouping + ordered float32 product) on the
// device.  Integer ids and the p_base_wrong bit patterns are bit-exact: the radix sort is stable, so every
// group keeps original call order, and the product is taken left to right starting from the first factor
// (1.0f * e == e), without flush-to-zero (nvcc default -ftz=false).
//
// The two stable sorts and the prefix sum use CUB device primitives (library code, like calling cuBLAS for a
// plain GEMM); matching, key construction, segmentation, the ordered product and the CSR/CSC index
// construction are the kernels below.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace dmx {

// ---------------------------------------------------------------------------------------------------------
// (a2) unpack + match
// ---------------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

__global__ void unpack_match_kernel(const uint8_t* __restrict__ calls, int64_t n_calls,
                                    const uint8_t* __restrict__ molecules, int64_t n_molecules, int molecule_stride,
                                    int64_t chrom_id, const int64_t* __restrict__ gkeys, const int32_t* __restrict__ gvids,
                                    int64_t n_variants, int32_t* __restrict__ out_variant,
                                    int32_t* __restrict__ out_cb, float* __restrict__ out_e) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_calls;
         k += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t* rec = calls + 13 * k;  // (molecule_index i4, snp_position i4, base_index u1, p_base_wrong f4)
        const int32_t mol = (int32_t)load_u32_unaligned(rec);
        const uint32_t pos = load_u32_unaligned(rec + 4);
        const uint32_t base = rec[8];
        const uint32_t e_bits = load_u32_unaligned(rec + 9);

        int32_t variant = -1;
        int32_t cb = -1;
        if (mol >= 0 && (int64_t)mol < n_molecules) {
            // compressed_cb: offset 0 of the 12-byte record, or a plain int32 array (stride 4, dmx_host_gather_cb)
            cb = *reinterpret_cast<const int32_t*>(molecules + (int64_t)molecule_stride * mol);
            const int64_t key = (chrom_id << 40) | ((int64_t)pos << 8) | (int64_t)base;
            int64_t lo = 0, hi = n_variants;  // lower_bound
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (gkeys[mid] < key) lo = mid + 1; else hi = mid;
            }
            if (lo < n_variants && gkeys[lo] == key) variant = gvids[lo];
        }
        out_variant[k] = variant;
        out_cb[k] = cb;
        out_e[k] = __uint_as_float(e_bits);
    }
}

// ---------------------------------------------------------------------------------------------------------
// (a3) grouping
// ---------------------------------------------------------------------------------------------------------

struct BuildCounters {
    unsigned long long n_matched;
    unsigned long long n_bad_barcode;
    unsigned long long n_rows;
    unsigned long long pad;
};

__global__ void make_keys_kernel(const int32_t* __restrict__ variant, const int32_t* __restrict__ cb, int64_t n,
                                 int64_t n_variants, int64_t n_barcodes, int64_t barcode_lo, int64_t barcode_hi,
                                 int cb_bits, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx,
                                 unsigned long long* __restrict__ n_mol, BuildCounters* counters) {
    const uint64_t sentinel = (uint64_t)n_variants << cb_bits;
    // the two counters are summed per thread over the grid-stride loop and reach global memory as one atomic per
    // CTA: a per-warp atomic on a single address (a million of them at 31 M calls) serialised in L2 and made this
    // streaming kernel ten times slower than its 20 bytes per call warrant
    unsigned n_matched = 0, n_bad = 0;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = variant[k];
        const int32_t b = cb[k];
        uint64_t key = sentinel;
        if (v >= 0 && (int64_t)v < n_variants) {
            if (b >= 0 && (int64_t)b < n_barcodes) {
                if (n_mol) atomicAdd(&n_mol[v], 1ull);  // the data prior counts every matched call (demux.py:381)
                if ((int64_t)b >= barcode_lo && (int64_t)b < barcode_hi) {  // this shard's barcodes
                    ++n_matched;
                    key = ((uint64_t)v << cb_bits) | (uint64_t)b;
                }
            } else {
                ++n_bad;
            }
        }
        keys[k] = key;
        idx[k] = (uint32_t)k;
    }
    __shared__ unsigned s_matched, s_bad;
    if (threadIdx.x == 0) { s_matched = 0; s_bad = 0; }
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_matched += __shfl_xor_sync(0xffffffffu, n_matched, o);
        n_bad += __shfl_xor_sync(0xffffffffu, n_bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_matched) atomicAdd(&s_matched, n_matched);
        if (n_bad) atomicAdd(&s_bad, n_bad);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_matched) atomicAdd(&counters->n_matched, (unsigned long long)s_matched);
        if (s_bad) atomicAdd(&counters->n_bad_barcode, (unsigned long long)s_bad);
    }
}

// head[i] = 1 when sorted call i starts a new (variant, barcode) group; sentinel keys never start one
__global__ void head_flags_kernel(const uint64_t* __restrict__ keys_sorted, int64_t n, uint64_t sentinel,
                                  int32_t* __restrict__ flags) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys_sorted[i];
        flags[i] = (k != sentinel && (i == 0 || keys_sorted[i - 1] != k)) ? 1 : 0;
    }
}

__global__ void finish_count_kernel(const int32_t* __restrict__ incl, BuildCounters* counters) {
    const unsigned long long m = counters->n_matched;
    counters->n_rows = m ? (unsigned long long)incl[m - 1] : 0ull;
}

__global__ void row_starts_kernel(const uint64_t* __restrict__ keys_sorted, const int32_t* __restrict__ incl,
                                  int64_t n_matched, int32_t* __restrict__ row_start) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_matched;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (i == 0 || keys_sorted[i - 1] != keys_sorted[i]) row_start[incl[i] - 1] = (int32_t)i;
    }
}

// one thread per row: ids from the key, ordered float32 product over the group's calls (original call order)
__global__ void emit_rows_kernel(const uint64_t* __restrict__ keys_sorted, const uint32_t* __restrict__ idx_sorted,
                                 const int32_t* __restrict__ row_start, const float* __restrict__ call_e,
                                 int64_t n_rows, int64_t n_matched, int cb_bits,
                                 int32_t* __restrict__ csc_variant, int32_t* __restrict__ csc_cb,
                                 float* __restrict__ csc_e, int32_t* __restrict__ csc_count,
                                 uint32_t* __restrict__ cb_keys, uint32_t* __restrict__ row_iota) {
    const uint64_t cb_mask = (1ull << cb_bits) - 1ull;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lo = row_start[r];
        const int64_t hi = (r + 1 < n_rows) ? (int64_t)row_start[r + 1] : n_matched;
        const uint64_t key = keys_sorted[lo];
        float prod = call_e[idx_sorted[lo]];
        for (int64_t i = lo + 1; i < hi; ++i) prod = __fmul_rn(prod, call_e[idx_sorted[i]]);
        const int32_t cb = (int32_t)(key & cb_mask);
        csc_variant[r] = (int32_t)(key >> cb_bits);
        csc_cb[r] = cb;
        csc_e[r] = prod;
        csc_count[r] = (int32_t)(hi - lo);
        cb_keys[r] = (uint32_t)cb;
        row_iota[r] = (uint32_t)r;
    }
}

// offsets[k] = first position in the ascending int32 array `sorted` holding a value >= k, k = 0..n_keys
template <typename T>
__global__ void lower_bound_offsets_kernel(const T* __restrict__ sorted, int64_t n, int64_t n_keys,
                                           int64_t* __restrict__ offsets) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k <= n_keys; k += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)sorted[mid] < k) lo = mid + 1; else hi = mid;
        }
        offsets[k] = lo;
    }
}

__global__ void gather_csr_kernel(const uint32_t* __restrict__ perm, const int32_t* __restrict__ csc_variant,
                                  const float* __restrict__ csc_e, int64_t n_rows, int32_t* __restrict__ csr_variant,
                                  float* __restrict__ csr_e) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t src = perm[r];
        csr_variant[r] = csc_variant[src];
        csr_e[r] = csc_e[src];
    }
}

// longest-processing-time-first schedule: barcodes by descending row count (ties: ascending id, the sort is stable)
__global__ void schedule_keys_kernel(const int64_t* __restrict__ offsets, int64_t n_barcodes,
                                     uint32_t* __restrict__ keys, uint32_t* __restrict__ ids) {
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < n_barcodes; b += (int64_t)gridDim.x * blockDim.x) {
        const int64_t rows = offsets[b + 1] - offsets[b];
        keys[b] = 0xFFFFFFFFu - (uint32_t)(rows < 0xFFFFFFFFll ? rows : 0xFFFFFFFFll);
        ids[b] = (uint32_t)b;
    }
}


// ---------------------------------------------------------------------------------------------------------
// sharded pack: every rank unpacks a slice of the calls, then the matched calls travel to the rank that owns their
// barcode (contiguous barcode ranges).  Stable: a rank's calls stay in call order inside every destination block, and
// blocks are received in rank order, so call order is kept inside every chromosome, hence inside every (variant,
// barcode) group, which is what the ordered products of demux.py:282-283 need.
// ---------------------------------------------------------------------------------------------------------

// matched calls per barcode (the weight the barcode ranges are balanced by)
__global__ void barcode_histogram_kernel(const int32_t* __restrict__ variant, const int32_t* __restrict__ cb, int64_t n,
                                         int64_t n_variants, int64_t n_barcodes, unsigned long long* __restrict__ hist) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = variant[k], b = cb[k];
        if (v >= 0 && (int64_t)v < n_variants && b >= 0 && (int64_t)b < n_barcodes) atomicAdd(&hist[b], 1ull);
    }
}

struct RouteCounters {
    unsigned long long per_rank[64];
    unsigned long long n_bad_barcode;
};

// destination of every call: rank owning its barcode, `world` for calls that are dropped (unmatched)
__global__ void route_keys_kernel(const int32_t* __restrict__ variant, const int32_t* __restrict__ cb, int64_t n,
                                  int64_t n_variants, int64_t n_barcodes, const int64_t* __restrict__ cuts, int world,
                                  uint8_t* __restrict__ keys, uint32_t* __restrict__ idx, RouteCounters* counters) {
    __shared__ unsigned s_count[65];
    for (int i = threadIdx.x; i < 65; i += blockDim.x) s_count[i] = 0;
    __syncthreads();
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = variant[k], b = cb[k];
        int dest = world;
        if (v >= 0 && (int64_t)v < n_variants) {
            if (b >= 0 && (int64_t)b < n_barcodes) {
                dest = 0;
                while (dest + 1 < world && (int64_t)b >= cuts[dest + 1]) ++dest;  // world <= 64: a short scan
                atomicAdd(&s_count[dest], 1u);
            } else {
                atomicAdd(&s_count[64], 1u);
            }
        }
        keys[k] = (uint8_t)dest;
        idx[k] = (uint32_t)k;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 65; i += blockDim.x) {
        const unsigned c = s_count[i];
        if (c) atomicAdd(i < 64 ? &counters->per_rank[i] : &counters->n_bad_barcode, (unsigned long long)c);
    }
}

__global__ void route_gather_kernel(const uint8_t* __restrict__ keys_sorted, const uint32_t* __restrict__ idx_sorted,
                                    int64_t n_kept, const int32_t* __restrict__ variant, const int32_t* __restrict__ cb,
                                    const float* __restrict__ e, const int64_t* __restrict__ cuts,
                                    int32_t* __restrict__ out_variant, int32_t* __restrict__ out_cb,
                                    float* __restrict__ out_e) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_kept; k += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t src = idx_sorted[k];
        out_variant[k] = variant[src];
        out_cb[k] = cb[src] - (int32_t)cuts[keys_sorted[k]];  // barcode id local to the owner's range
        out_e[k] = e[src];
    }
}

static int bits_for(int64_t max_value) {  // bits needed to represent values 0..max_value
    int b = 1;
    while ((max_value >> b) != 0) ++b;
    return b;
}

struct BuildLayout {
    size_t keys_a, keys_b, idx_a, idx_b, counters, cub_temp, cub_temp_bytes, total;
};

static int plan_build(int64_t n, int64_t n_variants, int64_t n_barcodes, BuildLayout* out) {
    size_t sort64 = 0, sort32 = 0, scan = 0;
    const int cb_bits = bits_for(n_barcodes > 0 ? n_barcodes - 1 : 0);
    const int v_bits = bits_for(n_variants);
    const int64_t m = n > 0 ? n : 1;
    DMX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort64, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                             (const uint32_t*)nullptr, (uint32_t*)nullptr, m, 0, cb_bits + v_bits));
    DMX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort32, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                             (const uint32_t*)nullptr, (uint32_t*)nullptr, m, 0, cb_bits));
    DMX_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan, (const int32_t*)nullptr, (int32_t*)nullptr, m));
    size_t temp = sort64 > sort32 ? sort64 : sort32;
    temp = temp > scan ? temp : scan;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (size_t)round_up((int64_t)bytes, 256); return o; };
    out->keys_a = take(8 * (size_t)m);
    out->keys_b = take(8 * (size_t)m);
    out->idx_a = take(4 * (size_t)m);
    out->idx_b = take(4 * (size_t)m);
    out->counters = take(sizeof(BuildCounters));
    out->cub_temp = take(temp);
    out->cub_temp_bytes = temp;
    out->total = off;
    return 0;
}

static inline int grid_for(int64_t n, int threads) {
    int64_t blocks = ceil_div(n > 0 ? n : 1, threads);
    const int64_t cap = (int64_t)sm_count() * 32;
    return (int)(blocks < cap ? blocks : cap);
}

}  // namespace dmx

extern "C" {

int dmx_unpack_match_calls(const uint8_t* snp_calls_packed, int64_t n_calls, const uint8_t* molecules_packed,
                           int64_t n_molecules, int32_t molecule_stride, int64_t chrom_id,
                           const int64_t* geno_keys_sorted,
                           const int32_t* geno_vids_sorted, int64_t n_variants, int32_t* out_variant,
                           int32_t* out_cb, float* out_e, void* stream) {
    if (n_calls <= 0) return 0;
    DMX_REQUIRE(chrom_id >= 0 && chrom_id < (1ll << 22), "chrom_id %lld out of range", (long long)chrom_id);
    DMX_REQUIRE(molecule_stride == 12 || molecule_stride == 4, "molecule_stride must be 12 (packed records) or 4");
    const int threads = 256;
    dmx::unpack_match_kernel<<<dmx::grid_for(n_calls, threads), threads, 0, (cudaStream_t)stream>>>(
        snp_calls_packed, n_calls, molecules_packed, n_molecules, molecule_stride, chrom_id, geno_keys_sorted,
        geno_vids_sorted,
        n_variants, out_variant, out_cb, out_e);
    DMX_LAUNCH_CHECK();
    return 0;
}


int dmx_barcode_histogram(const int32_t* call_variant, const int32_t* call_cb, int64_t n_calls, int64_t n_variants,
                          int64_t n_barcodes, int64_t* histogram, void* stream) {
    if (n_calls <= 0) return 0;
    dmx::barcode_histogram_kernel<<<dmx::grid_for(n_calls, 256), 256, 0, (cudaStream_t)stream>>>(
        call_variant, call_cb, n_calls, n_variants, n_barcodes, (unsigned long long*)histogram);
    DMX_LAUNCH_CHECK();
    return 0;
}

static int64_t route_layout(int64_t n_calls, size_t* keys_b, size_t* idx_a, size_t* idx_b, size_t* counters,
                            size_t* cub_temp, size_t* cub_bytes) {
    const int64_t m = n_calls > 0 ? n_calls : 1;
    size_t temp = 0;
    if (cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint8_t*)nullptr, (uint8_t*)nullptr,
                                        (const uint32_t*)nullptr, (uint32_t*)nullptr, m, 0, 7) != cudaSuccess)
        return -1;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (size_t)dmx::round_up((int64_t)bytes, 256); return o; };
    take((size_t)m);  // keys_a at offset 0
    *keys_b = take((size_t)m);
    *idx_a = take(4 * (size_t)m);
    *idx_b = take(4 * (size_t)m);
    *counters = take(sizeof(dmx::RouteCounters));
    *cub_temp = take(temp);
    *cub_bytes = temp;
    return (int64_t)off;
}

int64_t dmx_route_calls_workspace_bytes(int64_t n_calls) {
    size_t a, b, c, d, e, f;
    return route_layout(n_calls, &a, &b, &c, &d, &e, &f);
}

int dmx_route_calls(const int32_t* call_variant, const int32_t* call_cb, const float* call_e, int64_t n_calls,
                    int64_t n_variants, int64_t n_barcodes, const int64_t* cuts, int32_t world, void* workspace,
                    int64_t workspace_bytes, int32_t* out_variant, int32_t* out_cb, float* out_e, int64_t* h_counts,
                    void* stream_) {
    using namespace dmx;
    cudaStream_t stream = (cudaStream_t)stream_;
    DMX_REQUIRE(world >= 1 && world <= 64, "world size %d outside [1, 64]", (int)world);
    DMX_REQUIRE(n_calls >= 0 && n_calls < (1ll << 32) - 1, "n_calls out of range");
    for (int r = 0; r <= world; ++r) h_counts[r] = 0;
    if (n_calls == 0) return 0;
    size_t o_keys_b, o_idx_a, o_idx_b, o_counters, o_temp, temp_bytes;
    const int64_t need = route_layout(n_calls, &o_keys_b, &o_idx_a, &o_idx_b, &o_counters, &o_temp, &temp_bytes);
    DMX_REQUIRE(need >= 0 && workspace_bytes >= need, "route workspace too small");
    uint8_t* ws = (uint8_t*)workspace;
    uint8_t* keys_a = ws;
    uint8_t* keys_b = ws + o_keys_b;
    uint32_t* idx_a = (uint32_t*)(ws + o_idx_a);
    uint32_t* idx_b = (uint32_t*)(ws + o_idx_b);
    RouteCounters* counters = (RouteCounters*)(ws + o_counters);
    DMX_CUDA(cudaMemsetAsync(counters, 0, sizeof(RouteCounters), stream));
    route_keys_kernel<<<grid_for(n_calls, 256), 256, 0, stream>>>(call_variant, call_cb, n_calls, n_variants, n_barcodes,
                                                                  cuts, world, keys_a, idx_a, counters);
    DMX_LAUNCH_CHECK();
    DMX_CUDA(cub::DeviceRadixSort::SortPairs(ws + o_temp, temp_bytes, (const uint8_t*)keys_a, keys_b,
                                             (const uint32_t*)idx_a, idx_b, n_calls, 0, 7, stream));
    RouteCounters h;
    DMX_CUDA(cudaMemcpyAsync(&h, counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    DMX_CUDA(cudaStreamSynchronize(stream));
    h_counts[world] = (int64_t)h.n_bad_barcode;  // the caller raises on every rank (a collective follows)
    int64_t kept = 0;
    for (int r = 0; r < world; ++r) { h_counts[r] = (int64_t)h.per_rank[r]; kept += h_counts[r]; }
    if (kept > 0) {
        route_gather_kernel<<<grid_for(kept, 256), 256, 0, stream>>>(keys_b, idx_b, kept, call_variant, call_cb, call_e,
                                                                     cuts, out_variant, out_cb, out_e);
        DMX_LAUNCH_CHECK();
    }
    return 0;
}

int64_t dmx_barcode_schedule_workspace_bytes(int64_t n_barcodes) {
    size_t temp = 0;
    const int64_t m = n_barcodes > 0 ? n_barcodes : 1;
    if (cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                        (const uint32_t*)nullptr, (uint32_t*)nullptr, m) != cudaSuccess)
        return -1;
    return dmx::round_up((int64_t)temp, 256) + 3 * dmx::round_up(4 * m, 256);
}

int dmx_barcode_schedule(const int64_t* barcode_offsets, int64_t n_barcodes, int32_t* order, void* workspace,
                         int64_t workspace_bytes, void* stream_) {
    using namespace dmx;
    if (n_barcodes <= 0) return 0;
    DMX_REQUIRE(workspace_bytes >= dmx_barcode_schedule_workspace_bytes(n_barcodes), "schedule workspace too small");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t slice = round_up(4 * n_barcodes, 256);
    uint8_t* ws = (uint8_t*)workspace;
    uint32_t* keys_in = (uint32_t*)ws;
    uint32_t* keys_out = (uint32_t*)(ws + slice);
    uint32_t* ids_in = (uint32_t*)(ws + 2 * slice);
    void* temp = ws + 3 * slice;
    size_t temp_bytes = (size_t)(workspace_bytes - 3 * slice);
    schedule_keys_kernel<<<grid_for(n_barcodes, 256), 256, 0, stream>>>(barcode_offsets, n_barcodes, keys_in, ids_in);
    DMX_LAUNCH_CHECK();
    DMX_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const uint32_t*)keys_in, keys_out,
                                             (const uint32_t*)ids_in, (uint32_t*)order, n_barcodes, 0, 32, stream));
    return 0;
}

int64_t dmx_build_rows_workspace_bytes(int64_t n_calls, int64_t n_variants, int64_t n_barcodes) {
    dmx::BuildLayout lay;
    if (dmx::plan_build(n_calls, n_variants, n_barcodes, &lay)) return -1;
    return (int64_t)lay.total;
}

int dmx_build_rows(const int32_t* call_variant, const int32_t* call_cb, const float* call_e, int64_t n_calls,
                   int64_t n_variants, int64_t n_barcodes, int64_t barcode_lo, int64_t barcode_hi, void* workspace,
                   int64_t workspace_bytes,
                   int32_t* csc_variant, int32_t* csc_cb, float* csc_e, int32_t* csc_count,
                   int64_t* variant_offsets, int32_t* csr_variant, float* csr_e, int32_t* csr_row,
                   int64_t* barcode_offsets, int64_t* n_mol_per_variant, int64_t* h_n_rows, int64_t* h_n_matched,
                   void* stream_) {
    using namespace dmx;
    cudaStream_t stream = (cudaStream_t)stream_;
    DMX_REQUIRE(n_calls >= 0 && n_calls < (1ll << 31) - 1, "n_calls %lld not supported (must be < 2^31 - 1)",
                (long long)n_calls);
    DMX_REQUIRE(n_variants >= 0 && n_variants < (1ll << 31) - 1, "n_variants out of range");
    DMX_REQUIRE(n_barcodes >= 0 && n_barcodes < (1ll << 31) - 1, "n_barcodes out of range");
    BuildLayout lay;
    if (plan_build(n_calls, n_variants, n_barcodes, &lay)) return -1;
    DMX_REQUIRE(workspace_bytes >= (int64_t)lay.total, "workspace too small: %lld < %lld",
                (long long)workspace_bytes, (long long)lay.total);
    const int cb_bits = bits_for(n_barcodes > 0 ? n_barcodes - 1 : 0);
    const int v_bits = bits_for(n_variants);
    const uint64_t sentinel = (uint64_t)n_variants << cb_bits;

    uint8_t* ws = (uint8_t*)workspace;
    uint64_t* keys_a = (uint64_t*)(ws + lay.keys_a);
    uint64_t* keys_b = (uint64_t*)(ws + lay.keys_b);
    uint32_t* idx_a = (uint32_t*)(ws + lay.idx_a);
    uint32_t* idx_b = (uint32_t*)(ws + lay.idx_b);
    BuildCounters* counters = (BuildCounters*)(ws + lay.counters);
    void* cub_temp = ws + lay.cub_temp;
    size_t cub_bytes = lay.cub_temp_bytes;

    const int threads = 256;
    DMX_CUDA(cudaMemsetAsync(counters, 0, sizeof(BuildCounters), stream));
    if (n_variants > 0 && n_mol_per_variant)
        DMX_CUDA(cudaMemsetAsync(n_mol_per_variant, 0, sizeof(int64_t) * n_variants, stream));

    int64_t n_rows = 0, n_matched = 0;
    if (n_calls > 0) {
        make_keys_kernel<<<grid_for(n_calls, threads), threads, 0, stream>>>(
            call_variant, call_cb, n_calls, n_variants, n_barcodes, barcode_lo, barcode_hi, cb_bits, keys_a, idx_a,
            (unsigned long long*)n_mol_per_variant, counters);
        DMX_LAUNCH_CHECK();
        DMX_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, (const uint64_t*)keys_a, keys_b,
                                                 (const uint32_t*)idx_a, idx_b, n_calls, 0, cb_bits + v_bits, stream));
        // keys_a is free from here on: first half = flags / inclusive scan, second half = row starts
        int32_t* incl = (int32_t*)keys_a;
        int32_t* row_start = incl + n_calls;
        head_flags_kernel<<<grid_for(n_calls, threads), threads, 0, stream>>>(keys_b, n_calls, sentinel, incl);
        DMX_LAUNCH_CHECK();
        cub_bytes = lay.cub_temp_bytes;
        DMX_CUDA(cub::DeviceScan::InclusiveSum(cub_temp, cub_bytes, (const int32_t*)incl, incl, n_calls, stream));
        finish_count_kernel<<<1, 1, 0, stream>>>(incl, counters);
        DMX_LAUNCH_CHECK();
        BuildCounters h;
        DMX_CUDA(cudaMemcpyAsync(&h, counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
        DMX_CUDA(cudaStreamSynchronize(stream));
        DMX_REQUIRE(h.n_bad_barcode == 0,
                    "%llu matched calls carry a compressed_cb outside [0, n_barcodes=%lld)", h.n_bad_barcode,
                    (long long)n_barcodes);
        n_rows = (int64_t)h.n_rows;
        n_matched = (int64_t)h.n_matched;
        if (n_rows > 0) {
            row_starts_kernel<<<grid_for(n_matched, threads), threads, 0, stream>>>(keys_b, incl, n_matched, row_start);
            DMX_LAUNCH_CHECK();
            // idx_a (the sort input) is free: it receives the rows' barcode keys; the row iota that the
            // second sort permutes is staged in csr_variant, which is only filled for real by the gather below.
            emit_rows_kernel<<<grid_for(n_rows, threads), threads, 0, stream>>>(
                keys_b, idx_b, row_start, call_e, n_rows, n_matched, cb_bits, csc_variant, csc_cb, csc_e, csc_count,
                idx_a, (uint32_t*)csr_variant /* iota staged in csr_variant until the sort consumed it */);
            DMX_LAUNCH_CHECK();
            // stable sort of rows by barcode -> barcode-major permutation of the variant-major rows
            uint32_t* cb_sorted = (uint32_t*)keys_a;  // keys_a (flags/row_start) no longer needed after emit
            cub_bytes = lay.cub_temp_bytes;
            DMX_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, (const uint32_t*)idx_a, cb_sorted,
                                                     (const uint32_t*)csr_variant, (uint32_t*)csr_row, n_rows, 0,
                                                     cb_bits, stream));
            gather_csr_kernel<<<grid_for(n_rows, threads), threads, 0, stream>>>((const uint32_t*)csr_row, csc_variant,
                                                                                  csc_e, n_rows, csr_variant, csr_e);
            DMX_LAUNCH_CHECK();
            lower_bound_offsets_kernel<uint32_t><<<grid_for(n_barcodes + 1, threads), threads, 0, stream>>>(
                cb_sorted, n_rows, n_barcodes, barcode_offsets);
            DMX_LAUNCH_CHECK();
            lower_bound_offsets_kernel<int32_t><<<grid_for(n_variants + 1, threads), threads, 0, stream>>>(
                csc_variant, n_rows, n_variants, variant_offsets);
            DMX_LAUNCH_CHECK();
        }
    }
    if (n_rows == 0) {
        DMX_CUDA(cudaMemsetAsync(barcode_offsets, 0, sizeof(int64_t) * (n_barcodes + 1), stream));
        DMX_CUDA(cudaMemsetAsync(variant_offsets, 0, sizeof(int64_t) * (n_variants + 1), stream));
    }
    if (h_n_rows) *h_n_rows = n_rows;
    if (h_n_matched) *h_n_matched = n_matched;
    return 0;
}

}  // extern "C"
