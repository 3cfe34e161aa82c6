// E-step: barcode-segmented gather-reduce over the CSR rows (demux.py:246-265) with the doublet columns of
// demux.py:175-191, the doublet prior of demux.py:158-173, the optional prior logits of demux.py:97-99, and the
// row softmax of demux.py:101,152.
//
//   L[b, c] = float32( pen[c] + sum_{rows r of barcode b}^{float64} log( p_c[v_r] (1 - e_r) + max(e_r, 1e-4) ) )
//
// This file: the singlet-only kernel (doublet_prior == 0), the row softmax and the dmx_estep entry point.
// The pair kernel for doublet_prior != 0 lives in estep_pairs.cu.
//
// Arithmetic flavours (DMX_ESTEP_*) of the singlet kernel:
//   EXACT  x = fl(fl(P_g * (1-e)) + e'), t = logf(x) (float32), float64 accumulation: the reference's roundings.
//   FAST   x = fma(P_g, 1-e, e'), products of 8 consecutive row factors in float32 (every factor is in
//          [1e-4, 1.0001], so a product stays a normal number), one lg2.approx per product, float64 accumulation.
#include <math_constants.h>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace dmx {

constexpr int FLUSH_ROWS = 8;  // row factors multiplied before one lg2 (singlet kernel)
constexpr float ERROR_FLOOR = 1e-4f;

// estep_pairs.cu
int launch_estep_pairs(const int64_t* barcode_offsets, const int32_t* barcode_order, const int32_t* csr_variant,
                       const float* csr_e, int64_t n_barcodes, const float* table, int64_t ld_table, int G,
                       double doublet_prior,
                       float table_floor, const double* prior_logits, int64_t ld_prior, float* logits,
                       int64_t ld_logits, int flavour, cudaStream_t stream);

// estep_pairs_warp.cu
struct FusedSoftmax {
    float* post;
    int64_t ld_post;
    float* singlets;
    int64_t ld_singlet;
    bool logits_requested;
};
bool estep_pairs_warp_supported(int G, int flavour);
int launch_estep_pairs_warp(const int64_t* barcode_offsets, const int32_t* barcode_order, const int32_t* seg_prefix,
                            const int32_t* item_slot, int64_t n_items, int seg_rows, const int32_t* csr_variant,
                            const float* csr_e, const float* table, int64_t ld_table, int G, double doublet_prior,
                            float table_floor, const double* prior_logits, int64_t ld_prior, float* logits,
                            int64_t ld_logits, double* partial, int64_t n_cols, int flavour, const FusedSoftmax* fused,
                            int* did_fuse, cudaStream_t stream);
float pair_doublet_bonus(int n_genotypes, double dp);
int launch_plan_segments(const int64_t* offsets, const int32_t* order, int64_t n_barcodes, int seg_rows,
                         int32_t* n_seg, cudaStream_t stream);
int launch_plan_items(const int32_t* seg_prefix, int64_t n_barcodes, int64_t capacity, int32_t* item_slot,
                      cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------
// singlets only (doublet_prior == 0): one CTA of 4 warps per barcode, lanes over genotypes
// ---------------------------------------------------------------------------------------------------------------

// Each lane owns 4 consecutive genotypes (one 128-bit load of a table row), a row needs LPR lanes, a warp works on
// 32 / LPR rows at once with four such waves in flight; row records of 32 rows are read with one coalesced load per
// array and handed around with shuffles.
template <int FLAVOUR, int LPR, int SLOTS>
__global__ void __launch_bounds__(128) estep_singlets_kernel(const int64_t* __restrict__ offsets,
                                                             const int32_t* __restrict__ order,
                                                             const int32_t* __restrict__ variant,
                                                             const float* __restrict__ e_arr,
                                                             const float* __restrict__ table, int64_t ld_table,
                                                             int n_genotypes, const double* __restrict__ prior,
                                                             int64_t ld_prior, float* __restrict__ logits,
                                                             int64_t ld_logits) {
    constexpr int RGW = 32 / LPR;                  // rows a warp processes at once
    constexpr int WAVES = 4;                       // row waves in flight (measured: more costs occupancy)
    constexpr int PASSES_PER_FLUSH = FLUSH_ROWS / WAVES > 0 ? FLUSH_ROWS / WAVES : 1;
    __shared__ double partial[4][SLOTS * 4][32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane % LPR, rgw = lane / LPR;
    const int n_quads = (int)(ld_table / 4);
    const int64_t barcode = order ? (int64_t)order[blockIdx.x] : (int64_t)blockIdx.x;
    const int64_t lo = offsets[barcode], hi = offsets[barcode + 1];

    double acc[SLOTS][4];
    float prod[SLOTS][4];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[s][c] = 0.0; prod[s][c] = 1.f; }

    auto flush = [&]() {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[s][c] += (double)__log2f(prod[s][c]); prod[s][c] = 1.f; }
    };

    for (int64_t base = lo + 32 * warp; base < hi; base += 32 * 4) {
        const int n = (int)(hi - base < 32 ? hi - base : 32);
        const int32_t my_v = lane < n ? __ldg(variant + base + lane) : 0;
        const float my_e = lane < n ? __ldg(e_arr + base + lane) : 0.f;
        int pass = 0;
        for (int k0 = 0; k0 < n; k0 += RGW * WAVES, ++pass) {
            float4 pg[WAVES][SLOTS];
            float w[WAVES], ef[WAVES];
            bool valid[WAVES];
#pragma unroll
            for (int u = 0; u < WAVES; ++u) {
                const int k = k0 + u * RGW + rgw;
                const int src = k < n ? k : 0;
                valid[u] = k < n;
                const int32_t v = __shfl_sync(0xffffffffu, my_v, src);
                const float e = __shfl_sync(0xffffffffu, my_e, src);
                w[u] = __fsub_rn(1.f, e);
                ef[u] = fmaxf(e, ERROR_FLOOR);
                const float4* row = reinterpret_cast<const float4*>(table + (int64_t)v * ld_table);
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int q = sub + LPR * s;
                    pg[u][s] = (q < n_quads && valid[u]) ? __ldg(row + q) : make_float4(1.f, 1.f, 1.f, 1.f);
                }
            }
#pragma unroll
            for (int u = 0; u < WAVES; ++u)
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const float p4[4] = {pg[u][s].x, pg[u][s].y, pg[u][s].z, pg[u][s].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (FLAVOUR == DMX_ESTEP_FAST) {
                            prod[s][c] *= valid[u] ? fmaf(p4[c], w[u], ef[u]) : 1.f;
                        } else {
                            const float t = logf(__fadd_rn(__fmul_rn(p4[c], w[u]), ef[u]));
                            acc[s][c] += (double)(valid[u] ? t : 0.f);
                        }
                    }
                }
            if (FLAVOUR == DMX_ESTEP_FAST && (pass + 1) % PASSES_PER_FLUSH == 0) flush();
        }
        if (FLAVOUR == DMX_ESTEP_FAST) flush();
    }
    // combine the row groups of the warp in a fixed order, then the four warps
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double total = acc[s][c];
#pragma unroll
            for (int g = 1; g < RGW; ++g) total += __shfl_sync(0xffffffffu, acc[s][c], (sub + g * LPR) & 31);
            if (FLAVOUR == DMX_ESTEP_FAST) total *= 0.693147180559945309417232;
            partial[warp][s * 4 + c][lane] = total;
        }
    __syncthreads();
    if (warp == 0 && rgw == 0) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int g = 4 * (sub + LPR * s) + c;
                if (g < n_genotypes) {
                    const double sum = ((partial[0][s * 4 + c][lane] + partial[1][s * 4 + c][lane]) +
                                        partial[2][s * 4 + c][lane]) + partial[3][s * 4 + c][lane];
                    float logit = (float)(0.0 + sum);
                    if (prior) logit = (float)((double)logit + prior[barcode * ld_prior + g]);
                    logits[barcode * ld_logits + g] = logit;
                }
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// row softmax: one CTA per barcode
// ---------------------------------------------------------------------------------------------------------------

// Segment sums of the warp pair kernel (estep_pairs_warp.cu): a barcode cut into several work items has its float64
// log2-sums added here in segment order, then penalty, prior logits and the single rounding to float32.
struct CombineParams {
    const int32_t* order;       // schedule slot -> barcode (or nullptr); blockIdx.x is a slot when seg_prefix is set
    const int32_t* seg_prefix;  // nullptr: no combine stage, blockIdx.x is the barcode
    const double* partial;
    const double* prior;
    int64_t ld_prior;
    int n_singlets;
    float doublet_bonus;
    double scale;  // partial sums are log2-sums (ln 2) or natural-log sums (1)
    int skip_single;  // single-item barcodes were finished (softmax included) by the pair kernel itself
};

// DMX_ESTEP_AUTO: the reference's own roundings for the singlet-only E-step (doublet_prior == 0), the product
// arithmetic for every kernel with doublet columns.  Measured (profiles/r02_flavours_small.log, r01/r02 parity
// reports): every posterior that missed the 1e-6 bar under FAST was a doublet_prior == 0 case (13 of 161 in round 1),
// and EXACT costs 1.3-2.4 x there (0.37 -> 0.89 ms at G = 32, 24 M rows); with doublet columns FAST stayed inside
// 1e-6 for G <= 8 in every tested case while EXACT costs 2.2 x (G = 4) to 7.1 x (G = 8) on the lane-per-row kernel.
static inline int resolve_flavour(int flavour, int G, double doublet_prior) {
    if (flavour != DMX_ESTEP_AUTO) return flavour;
    (void)G;
    return doublet_prior == 0 ? DMX_ESTEP_EXACT : DMX_ESTEP_FAST;
}

__global__ void __launch_bounds__(128) softmax_rows_kernel(float* __restrict__ logits, int64_t ld_logits, int n_cols,
                                                           float* __restrict__ post, int64_t ld_post,
                                                           float* __restrict__ singlets, int64_t ld_singlet,
                                                           int n_singlets, const CombineParams cp) {
    __shared__ float s_max[4];
    __shared__ double s_sum[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t barcode = blockIdx.x;
    if (cp.seg_prefix) {
        const int seg_first = cp.seg_prefix[blockIdx.x];
        const int n_seg = cp.seg_prefix[blockIdx.x + 1] - seg_first;
        if (cp.order) barcode = cp.order[blockIdx.x];
        if (n_seg == 1 && cp.skip_single) return;
        if (n_seg > 1) {  // every thread later re-reads only the columns it writes here
            float* out = logits + barcode * ld_logits;
            for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
                double sum = 0.0;
                for (int k = 0; k < n_seg; ++k) sum += cp.partial[(int64_t)(seg_first + k) * n_cols + c];
                const float pen = c < cp.n_singlets ? 0.f : cp.doublet_bonus;
                float logit = (float)((double)pen + sum * cp.scale);
                if (cp.prior) logit = (float)((double)logit + cp.prior[barcode * cp.ld_prior + c]);
                out[c] = logit;
            }
        }
        if (!post && !singlets) return;
    }
    const float* row = logits + barcode * ld_logits;

    float m = -CUDART_INF_F;
    for (int c = threadIdx.x; c < n_cols; c += blockDim.x) m = fmaxf(m, row[c]);
    m = warp_max(m);
    if (lane == 0) s_max[warp] = m;
    __syncthreads();
    m = fmaxf(fmaxf(s_max[0], s_max[1]), fmaxf(s_max[2], s_max[3]));

    double sum = 0.0;
    for (int c = threadIdx.x; c < n_cols; c += blockDim.x) sum += (double)expf(__fsub_rn(row[c], m));
    sum = warp_sum(sum);
    if (lane == 0) s_sum[warp] = sum;
    __syncthreads();
    const float total = (float)(((s_sum[0] + s_sum[1]) + s_sum[2]) + s_sum[3]);

    for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
        const float pr = __fdiv_rn(expf(__fsub_rn(row[c], m)), total);
        if (post) post[barcode * ld_post + c] = pr;
        if (singlets && c < n_singlets) singlets[barcode * ld_singlet + c] = pr;
    }
}

static int launch_softmax(float* logits, int64_t ld_logits, int64_t n_rows, int n_cols, float* post,
                          int64_t ld_post, float* singlets, int64_t ld_singlet, int n_singlets, cudaStream_t stream,
                          const CombineParams* combine = nullptr) {
    if (n_rows <= 0 || n_cols <= 0) return 0;
    CombineParams cp = {};
    if (combine) cp = *combine;
    if (!post && !singlets && !cp.seg_prefix) return 0;
    softmax_rows_kernel<<<(unsigned)n_rows, 128, 0, stream>>>(logits, ld_logits, n_cols, post, ld_post, singlets,
                                                              ld_singlet, n_singlets, cp);
    DMX_LAUNCH_CHECK();
    return 0;
}

template <int FLAVOUR>
static int launch_singlets(unsigned grid, cudaStream_t stream, const int64_t* offsets, const int32_t* order,
                           const int32_t* variant, const float* e, const float* table, int64_t ld_table, int G,
                           const double* prior, int64_t ld_prior, float* logits, int64_t ld_logits) {
    const int quads = (int)(ld_table / 4);
#define DMX_SINGLETS(LPR, SLOTS)                                                                                  \
    estep_singlets_kernel<FLAVOUR, LPR, SLOTS><<<grid, 128, 0, stream>>>(offsets, order, variant, e, table, ld_table, \
                                                                         G, prior, ld_prior, logits, ld_logits)
    if (quads <= 1) DMX_SINGLETS(1, 1);  // few genotypes: fewer lanes per row, more rows side by side in a warp
    else if (quads <= 2) DMX_SINGLETS(2, 1);
    else if (quads <= 4) DMX_SINGLETS(4, 1);
    else if (quads <= 8) DMX_SINGLETS(8, 1);
    else if (quads <= 16) DMX_SINGLETS(16, 1);
    else if (quads <= 32) DMX_SINGLETS(32, 1);
    else if (quads <= 64) DMX_SINGLETS(32, 2);
    else if (quads <= 128) DMX_SINGLETS(32, 4);
    else {
        set_error("singlet E-step supports up to 512 genotypes");
        return -2;
    }
#undef DMX_SINGLETS
    DMX_LAUNCH_CHECK();
    return 0;
}

}  // namespace dmx

extern "C" {

static int64_t logits_scratch_bytes(int64_t n_barcodes, int64_t cols) {
    return dmx::round_up(n_barcodes * cols * (int64_t)sizeof(float), 256) + 256;
}

int64_t dmx_estep_workspace_bytes(int64_t n_barcodes, int32_t n_genotypes, double doublet_prior, int64_t n_items,
                                  int32_t need_logits_scratch) {
    const int64_t cols = doublet_prior == 0 ? n_genotypes : (int64_t)n_genotypes * (n_genotypes + 1) / 2;
    int64_t bytes = need_logits_scratch ? logits_scratch_bytes(n_barcodes, cols) : 0;
    if (n_items > n_barcodes) bytes += dmx::round_up(n_items * cols * (int64_t)sizeof(double), 256);
    return bytes;
}

int dmx_softmax_rows(const float* logits, int64_t ld_logits, int64_t n_rows, int32_t n_cols, float* posteriors,
                     int64_t ld_post, float* singlet_posteriors, int64_t ld_singlet, int32_t n_singlets,
                     void* stream) {
    return dmx::launch_softmax(const_cast<float*>(logits), ld_logits, n_rows, n_cols, posteriors, ld_post,
                               singlet_posteriors, ld_singlet, n_singlets, (cudaStream_t)stream);
}

int dmx_estep_plan_supported(int32_t n_genotypes, double doublet_prior, int32_t flavour) {
    flavour = dmx::resolve_flavour(flavour, n_genotypes, doublet_prior);
    if (!dmx::estep_pairs_warp_supported(n_genotypes, flavour)) return 0;
    return (doublet_prior != 0 || n_genotypes <= 8) ? 1 : 0;  // singlets only: the lane-per-row kernel (G <= 8)
}

int64_t dmx_estep_plan_workspace_bytes(int64_t n_barcodes) {
    size_t temp = 0;
    const int64_t m = n_barcodes + 1;
    if (cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr, m) != cudaSuccess)
        return -1;
    return dmx::round_up((int64_t)temp, 256) + dmx::round_up(4 * m, 256);
}

int dmx_estep_plan(const int64_t* barcode_offsets, const int32_t* barcode_order, int64_t n_barcodes, int32_t seg_rows,
                   int32_t* seg_prefix, int32_t* item_slot, int64_t item_capacity, void* workspace,
                   int64_t workspace_bytes, int64_t* h_n_items, void* stream_) {
    using namespace dmx;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (h_n_items) *h_n_items = 0;
    if (n_barcodes <= 0) return 0;
    DMX_REQUIRE(seg_rows >= 16 && seg_rows <= 4096, "seg_rows %d outside [16, 4096]", seg_rows);
    DMX_REQUIRE(workspace_bytes >= dmx_estep_plan_workspace_bytes(n_barcodes), "plan workspace too small");
    const int64_t m = n_barcodes + 1;
    int32_t* n_seg = (int32_t*)workspace;
    void* temp = (uint8_t*)workspace + round_up(4 * m, 256);
    size_t temp_bytes = (size_t)(workspace_bytes - round_up(4 * m, 256));
    if (launch_plan_segments(barcode_offsets, barcode_order, n_barcodes, seg_rows, n_seg, stream)) return -1;
    DMX_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, (const int32_t*)n_seg, seg_prefix, m, stream));
    if (launch_plan_items(seg_prefix, n_barcodes, item_capacity, item_slot, stream)) return -1;
    int32_t n_items = 0;
    DMX_CUDA(cudaMemcpyAsync(&n_items, seg_prefix + n_barcodes, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    DMX_CUDA(cudaStreamSynchronize(stream));
    DMX_REQUIRE(n_items <= item_capacity, "item_slot too small: %d items, capacity %lld", n_items,
                (long long)item_capacity);
    if (h_n_items) *h_n_items = n_items;
    return 0;
}

int dmx_estep(const int64_t* barcode_offsets, const int32_t* barcode_order, const int32_t* csr_variant,
              const float* csr_e, int64_t n_barcodes,
              const float* table, int64_t ld_table, int32_t n_genotypes, double doublet_prior,
              const double* prior_logits, int64_t ld_prior, float* logits, int64_t ld_logits, float* posteriors,
              int64_t ld_post, float* singlet_posteriors, int64_t ld_singlet, void* workspace,
              int64_t workspace_bytes, int32_t flavour, float table_floor, const int32_t* seg_prefix,
              const int32_t* item_slot, int64_t n_items, int32_t seg_rows, void* stream_) {
    using namespace dmx;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_barcodes <= 0) return 0;
    const int G = n_genotypes;
    DMX_REQUIRE(G >= 1, "n_genotypes must be positive");
    DMX_REQUIRE(doublet_prior >= 0 && doublet_prior < 1, "doublet_prior must be in [0, 1)");
    DMX_REQUIRE(flavour == DMX_ESTEP_EXACT || flavour == DMX_ESTEP_FAST || flavour == DMX_ESTEP_AUTO,
                "unknown E-step flavour %d", flavour);
    flavour = resolve_flavour(flavour, G, doublet_prior);
    DMX_REQUIRE(ld_table % 4 == 0 && ld_table >= G, "ld_table must be a multiple of 4 and >= n_genotypes");
    DMX_REQUIRE(((uintptr_t)table & 15) == 0, "table must be 16-byte aligned");
    const int64_t n_cols = doublet_prior == 0 ? G : (int64_t)G * (G + 1) / 2;
    DMX_REQUIRE(n_cols < (1ll << 31), "too many columns");

    const bool planned = seg_prefix && item_slot && n_items > 0 && (doublet_prior != 0 || G <= 8) &&
                         estep_pairs_warp_supported(G, flavour);
    uint8_t* ws = (uint8_t*)workspace;
    int64_t ws_left = workspace ? workspace_bytes : 0;
    float* out_logits = logits;
    int64_t ld_out = ld_logits;
    if (!out_logits) {
        const int64_t need = logits_scratch_bytes(n_barcodes, n_cols);
        DMX_REQUIRE(ws_left >= need, "workspace too small for the logits scratch");
        out_logits = (float*)ws;
        ld_out = n_cols;
        ws += need;
        ws_left -= need;
    }
    double* partial = nullptr;
    if (planned && n_items > n_barcodes) {
        DMX_REQUIRE(ws_left >= n_items * n_cols * (int64_t)sizeof(double), "workspace too small for the segment sums");
        partial = (double*)ws;
    }

    if (doublet_prior == 0 && !planned) {
        int rc;
        if (flavour == DMX_ESTEP_FAST)
            rc = launch_singlets<DMX_ESTEP_FAST>((unsigned)n_barcodes, stream, barcode_offsets, barcode_order, csr_variant,
                                                 csr_e, table, ld_table, G, prior_logits, ld_prior, out_logits, ld_out);
        else
            rc = launch_singlets<DMX_ESTEP_EXACT>((unsigned)n_barcodes, stream, barcode_offsets, barcode_order, csr_variant,
                                                  csr_e, table, ld_table, G, prior_logits, ld_prior, out_logits, ld_out);
        if (rc) return rc;
    } else if (planned) {
        FusedSoftmax fused = {posteriors, ld_post, singlet_posteriors, ld_singlet, logits != nullptr};
        int did_fuse = 0;
        const int rc = launch_estep_pairs_warp(barcode_offsets, barcode_order, seg_prefix, item_slot, n_items, seg_rows,
                                               csr_variant, csr_e, table, ld_table, G, doublet_prior, table_floor,
                                               prior_logits, ld_prior, out_logits, ld_out, partial, n_cols, flavour,
                                               &fused, &did_fuse, stream);
        if (rc) return rc;
        if (did_fuse && !partial) return 0;  // every barcode was a single work item: nothing left to do
        if (partial) {
            CombineParams cp;
            cp.skip_single = did_fuse;
            cp.order = barcode_order;
            cp.seg_prefix = seg_prefix;
            cp.partial = partial;
            cp.prior = prior_logits;
            cp.ld_prior = ld_prior;
            cp.n_singlets = G;
            cp.doublet_bonus = doublet_prior == 0 ? 0.f : pair_doublet_bonus(G, doublet_prior);
            cp.scale = flavour == DMX_ESTEP_EXACT ? 1.0 : 0.693147180559945309417232;
            return launch_softmax(out_logits, ld_out, n_barcodes, (int)n_cols, posteriors, ld_post, singlet_posteriors,
                                  ld_singlet, G, stream, &cp);
        }
    } else {
        const int rc = launch_estep_pairs(barcode_offsets, barcode_order, csr_variant, csr_e, n_barcodes, table, ld_table, G,
                                          doublet_prior, table_floor, prior_logits, ld_prior, out_logits, ld_out,
                                          flavour, stream);
        if (rc) return rc;
    }
    return launch_softmax(out_logits, ld_out, n_barcodes, (int)n_cols, posteriors, ld_post, singlet_posteriors,
                          ld_singlet, G, stream);
}

}  // extern "C"
