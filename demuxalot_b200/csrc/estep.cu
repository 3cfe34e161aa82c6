// E-step: barcode-segmented gather-reduce over the CSR rows (demux.py:246-265) with the doublet columns of
// demux.py:175-191, the doublet prior of demux.py:158-173, the optional prior logits of demux.py:97-99, and the
// row softmax of demux.py:101,152.
//
//   L[b, c] = float32( pen[c] + sum_{rows r of barcode b}^{float64} log( p_c[v_r] (1 - e_r) + max(e_r, 1e-4) ) )
//
// Two kernels:
//   * estep_pairs_kernel  (doublet_prior != 0): "log-semiring SYRK".  The G(G+1)/2 columns of a barcode are the
//     upper triangle (diagonal = singlets) of a G x G pair matrix; each thread owns an 8 x 4 register tile of it
//     and walks the barcode's rows, whose table rows P[v_r, :] are staged in shared memory with cp.async
//     (double buffered).  Row groups split a barcode's rows inside the CTA and are reduced in a fixed order
//     (no atomics -> deterministic).  Bound: FP32 issue / MUFU, not HBM (0.26 B per update at G = 32).
//   * estep_singlets_kernel (doublet_prior == 0): lanes over genotypes, warps over rows; HBM / L2-gather bound.
//
// Arithmetic flavours (DMX_ESTEP_*):
//   EXACT  x = fl(fl(fl(P_i + P_j) * (0.5 (1-e))) + e'), t = logf(x) (float32), float64 accumulation: the
//          reference's per-term roundings (0.5 scaling is exact, so folding it into (1-e) changes nothing).
//   FAST   a_g = fma(P_g, 1-e, e') once per row and genotype, x = a_i + a_j (= 2x the reference argument up to
//          one rounding), products of 8 consecutive row factors in float32 (every factor is in [2e-4, 2.0002], so
//          no under/overflow), one lg2.approx per product, float64 accumulation, the factor 1/2 per row removed
//          exactly at the end.  More accurate than summing rounded logs; differs from numpy by less than
//          numpy's own float32 log error (tests/test_gpu_parity.py reports the distribution).
#include <math_constants.h>

#include "common.cuh"

namespace dmx {

constexpr int TILE_I = 8;
constexpr int TILE_J = 4;
constexpr int FLUSH_ROWS = 8;  // row factors multiplied before one lg2
constexpr float ERROR_FLOOR = 1e-4f;

struct PairsParams {
    const int64_t* offsets;
    const int32_t* variant;
    const float* e;
    const float* table;
    int64_t ld_table;
    int n_genotypes;
    int gp;             // genotypes rounded up to a multiple of 8
    int n_tiles;        // 8x4 tiles covering the upper triangle
    int tiles_per_cta;  // tiles handled by one CTA
    int ctas_per_barcode;
    int row_groups;     // row groups inside a CTA
    int flushes;        // products (of FLUSH_ROWS rows) per row group and staged chunk
    int ld_smem;        // floats per staged row (gp + 4)
    float doublet_bonus;
    const float* prior;
    int64_t ld_prior;
    float* logits;
    int64_t ld_logits;
};

template <int FLAVOUR>
__global__ void __launch_bounds__(256) estep_pairs_kernel(const PairsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int n_threads = blockDim.x;
    const int64_t barcode = blockIdx.x / p.ctas_per_barcode;
    const int cta_in_barcode = (int)(blockIdx.x - barcode * p.ctas_per_barcode);

    const int tile_local = tid % p.tiles_per_cta;
    const int rg = tid / p.tiles_per_cta;
    const int tile = cta_in_barcode * p.tiles_per_cta + tile_local;
    const bool has_tile = tile < p.n_tiles;

    // tile -> (i0, j0): tiles are enumerated i-block major, for i-block pi the j-blocks 2*pi .. gp/4-1
    int i0 = 0, j0 = 0;
    {
        const int q_total = p.gp / TILE_J;
        int t = has_tile ? tile : 0, pi = 0;
        while (t >= q_total - 2 * pi) { t -= q_total - 2 * pi; ++pi; }
        i0 = pi * TILE_I;
        j0 = (2 * pi + t) * TILE_J;
    }

    const int chunk_rows = p.row_groups * p.flushes * FLUSH_ROWS;
    const int ld = p.ld_smem;
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    float* stage1 = stage0 + (size_t)chunk_rows * ld;
    double* reduce_buf = reinterpret_cast<double*>(stage1 + (size_t)chunk_rows * ld);

    const int64_t row_lo = p.offsets[barcode];
    const int64_t row_hi = p.offsets[barcode + 1];
    const int64_t n_rows = row_hi - row_lo;
    const int n_chunks = (int)((n_rows + chunk_rows - 1) / chunk_rows);

    const int quads = p.gp / 4;               // 16-byte pieces per staged row
    const int items = chunk_rows * quads;     // pieces per chunk
    constexpr int MAX_ITEMS = 8;              // per thread (host guarantees items <= MAX_ITEMS * n_threads)

    double acc[TILE_I][TILE_J];
#pragma unroll
    for (int a = 0; a < TILE_I; ++a)
#pragma unroll
        for (int b = 0; b < TILE_J; ++b) acc[a][b] = 0.0;

    // ---- staging of one chunk: issue() starts the copies, land() finishes the thread's own pieces ------------
    float item_e[MAX_ITEMS];  // p_base_wrong of the rows whose pieces this thread copies
    unsigned live = 0;        // bit s set: piece s is a real table piece that land() must finish
    auto issue = [&](int chunk, float* buf) {
        const int64_t base = row_lo + (int64_t)chunk * chunk_rows;
        live = 0;
#pragma unroll
        for (int s = 0; s < MAX_ITEMS; ++s) {
            const int it = tid + s * n_threads;
            item_e[s] = 0.f;
            if (it < items) {
                const int r = it / quads;
                const int q = it - r * quads;
                float* dst = buf + (size_t)r * ld + 4 * q;
                const int64_t row = base + r;
                if (row < row_hi && 4 * q < p.ld_table) {
                    const int32_t v = p.variant[row];
                    item_e[s] = p.e[row];
                    live |= 1u << s;
                    cp_async_16(dst, p.table + (int64_t)v * p.ld_table + 4 * q);
                } else {
                    // padding row (neutral: contributes log 1) or genotype padding beyond the table width
                    *reinterpret_cast<float4*>(dst) = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (FLAVOUR == DMX_ESTEP_EXACT && q == 0) {  // only reached for padding rows (row >= row_hi)
                        dst[p.gp] = 0.5f;       // (1 + 1) * 0.5 + 0 = 1 -> log = 0
                        dst[p.gp + 1] = 0.f;
                    }
                }
            }
        }
        cp_async_commit();
    };
    auto land = [&](float* buf) {
        cp_async_wait<0>();
#pragma unroll
        for (int s = 0; s < MAX_ITEMS; ++s) {
            if (live & (1u << s)) {
                const int it = tid + s * n_threads;
                const int r = it / quads;
                const int q = it - r * quads;
                float* dst = buf + (size_t)r * ld + 4 * q;
                const float e = item_e[s];
                const float w = __fsub_rn(1.f, e);
                const float ef = fmaxf(e, ERROR_FLOOR);
                if (FLAVOUR == DMX_ESTEP_FAST) {
                    float4 x = *reinterpret_cast<float4*>(dst);
                    x.x = fmaf(x.x, w, ef);
                    x.y = fmaf(x.y, w, ef);
                    x.z = fmaf(x.z, w, ef);
                    x.w = fmaf(x.w, w, ef);
                    *reinterpret_cast<float4*>(dst) = x;
                } else if (q == 0) {
                    dst[p.gp] = 0.5f * w;  // exact scaling
                    dst[p.gp + 1] = ef;
                }
            }
        }
    };

    if (n_chunks > 0) {
        issue(0, stage0);
        land(stage0);
        __syncthreads();
    }

    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        float* cur = (chunk & 1) ? stage1 : stage0;
        float* nxt = (chunk & 1) ? stage0 : stage1;
        const bool more = chunk + 1 < n_chunks;
        if (more) issue(chunk + 1, nxt);

        if (has_tile) {
            for (int f = 0; f < p.flushes; ++f) {
                if (FLAVOUR == DMX_ESTEP_FAST) {
                    float prod[TILE_I][TILE_J];
#pragma unroll
                    for (int a = 0; a < TILE_I; ++a)
#pragma unroll
                        for (int b = 0; b < TILE_J; ++b) prod[a][b] = 1.f;
#pragma unroll
                    for (int k = 0; k < FLUSH_ROWS; ++k) {
                        const float* s = cur + (size_t)((f * FLUSH_ROWS + k) * p.row_groups + rg) * ld;
                        const float4 a0 = *reinterpret_cast<const float4*>(s + i0);
                        const float4 a1 = *reinterpret_cast<const float4*>(s + i0 + 4);
                        const float4 bj = *reinterpret_cast<const float4*>(s + j0);
                        const float ai[TILE_I] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const float aj[TILE_J] = {bj.x, bj.y, bj.z, bj.w};
#pragma unroll
                        for (int a = 0; a < TILE_I; ++a)
#pragma unroll
                            for (int b = 0; b < TILE_J; ++b) prod[a][b] *= (ai[a] + aj[b]);
                    }
#pragma unroll
                    for (int a = 0; a < TILE_I; ++a)
#pragma unroll
                        for (int b = 0; b < TILE_J; ++b) acc[a][b] += (double)__log2f(prod[a][b]);
                } else {
#pragma unroll 2
                    for (int k = 0; k < FLUSH_ROWS; ++k) {
                        const float* s = cur + (size_t)((f * FLUSH_ROWS + k) * p.row_groups + rg) * ld;
                        const float4 a0 = *reinterpret_cast<const float4*>(s + i0);
                        const float4 a1 = *reinterpret_cast<const float4*>(s + i0 + 4);
                        const float4 bj = *reinterpret_cast<const float4*>(s + j0);
                        const float hw = s[p.gp];
                        const float ef = s[p.gp + 1];
                        const float pi[TILE_I] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const float pj[TILE_J] = {bj.x, bj.y, bj.z, bj.w};
#pragma unroll
                        for (int a = 0; a < TILE_I; ++a)
#pragma unroll
                            for (int b = 0; b < TILE_J; ++b) {
                                const float x = __fadd_rn(__fmul_rn(__fadd_rn(pi[a], pj[b]), hw), ef);
                                acc[a][b] += (double)logf(x);
                            }
                    }
                }
            }
        }

        if (more) land(nxt);
        __syncthreads();
    }

    // ---- fixed-order reduction over the row groups (deterministic) ---------------------------------------------
    if (p.row_groups > 1) {
        for (int g = 0; g < p.row_groups; ++g) {
            if (rg == g && has_tile) {
#pragma unroll
                for (int a = 0; a < TILE_I; ++a)
#pragma unroll
                    for (int b = 0; b < TILE_J; ++b) {
                        double* slot = reduce_buf + (size_t)(a * TILE_J + b) * p.tiles_per_cta + tile_local;
                        if (g == 0) *slot = acc[a][b]; else *slot += acc[a][b];
                    }
            }
            __syncthreads();
        }
        if (rg == 0 && has_tile) {
#pragma unroll
            for (int a = 0; a < TILE_I; ++a)
#pragma unroll
                for (int b = 0; b < TILE_J; ++b)
                    acc[a][b] = reduce_buf[(size_t)(a * TILE_J + b) * p.tiles_per_cta + tile_local];
        }
    }

    // ---- epilogue: penalties, prior logits, one rounding to float32 ---------------------------------------------
    if (rg == 0 && has_tile) {
        const int G = p.n_genotypes;
        const double padded_rows = (double)n_chunks * (double)chunk_rows;
#pragma unroll
        for (int a = 0; a < TILE_I; ++a) {
            const int i = i0 + a;
#pragma unroll
            for (int b = 0; b < TILE_J; ++b) {
                const int j = j0 + b;
                if (i < G && j < G && j >= i) {
                    const int64_t col = (i == j) ? i : (int64_t)G + (int64_t)i * G - (int64_t)i * (i + 1) / 2 + (j - i - 1);
                    double sum = acc[a][b];
                    if (FLAVOUR == DMX_ESTEP_FAST) sum = (sum - padded_rows) * 0.693147180559945309417232;
                    const float pen = (i == j) ? 0.f : p.doublet_bonus;
                    float logit = (float)((double)pen + sum);
                    if (p.prior) logit = (float)((double)logit + (double)p.prior[barcode * p.ld_prior + col]);
                    p.logits[barcode * p.ld_logits + col] = logit;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// singlets only (doublet_prior == 0): one CTA of 4 warps per barcode, lanes over genotypes
// ---------------------------------------------------------------------------------------------------------------

template <int FLAVOUR, int SLOTS>
__global__ void __launch_bounds__(128) estep_singlets_kernel(const int64_t* __restrict__ offsets,
                                                             const int32_t* __restrict__ variant,
                                                             const float* __restrict__ e_arr,
                                                             const float* __restrict__ table, int64_t ld_table,
                                                             int n_genotypes, const float* __restrict__ prior,
                                                             int64_t ld_prior, float* __restrict__ logits,
                                                             int64_t ld_logits) {
    __shared__ double partial[4][SLOTS * 32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t barcode = blockIdx.x;
    const int64_t lo = offsets[barcode], hi = offsets[barcode + 1];

    double acc[SLOTS];
    float prod[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) { acc[s] = 0.0; prod[s] = 1.f; }

    int in_prod = 0;
    for (int64_t r = lo + warp; r < hi; r += 4) {
        const int64_t v = variant[r];
        const float e = e_arr[r];
        const float w = __fsub_rn(1.f, e);
        const float ef = fmaxf(e, ERROR_FLOOR);
        const float* row = table + v * ld_table;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int g = lane + 32 * s;
            const float pg = (g < n_genotypes) ? __ldg(row + g) : 1.f;
            if (FLAVOUR == DMX_ESTEP_FAST) {
                prod[s] *= fmaf(pg, w, ef);
            } else {
                acc[s] += (double)logf(__fadd_rn(__fmul_rn(pg, w), ef));
            }
        }
        if (FLAVOUR == DMX_ESTEP_FAST && ++in_prod == FLUSH_ROWS) {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) { acc[s] += (double)__log2f(prod[s]); prod[s] = 1.f; }
            in_prod = 0;
        }
    }
    if (FLAVOUR == DMX_ESTEP_FAST) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) acc[s] = (acc[s] + (double)__log2f(prod[s])) * 0.693147180559945309417232;
    }
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) partial[warp][s * 32 + lane] = acc[s];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int g = lane + 32 * s;
            if (g < n_genotypes) {
                const double sum = ((partial[0][s * 32 + lane] + partial[1][s * 32 + lane]) +
                                    partial[2][s * 32 + lane]) + partial[3][s * 32 + lane];
                float logit = (float)(0.0 + sum);
                if (prior) logit = (float)((double)logit + (double)prior[barcode * ld_prior + g]);
                logits[barcode * ld_logits + g] = logit;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// row softmax: one CTA per barcode
// ---------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) softmax_rows_kernel(const float* __restrict__ logits, int64_t ld_logits,
                                                           int n_cols, float* __restrict__ post, int64_t ld_post,
                                                           float* __restrict__ singlets, int64_t ld_singlet,
                                                           int n_singlets) {
    __shared__ float s_max[4];
    __shared__ double s_sum[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* row = logits + (int64_t)blockIdx.x * ld_logits;

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
        if (post) post[(int64_t)blockIdx.x * ld_post + c] = pr;
        if (singlets && c < n_singlets) singlets[(int64_t)blockIdx.x * ld_singlet + c] = pr;
    }
}

static int launch_softmax(const float* logits, int64_t ld_logits, int64_t n_rows, int n_cols, float* post,
                          int64_t ld_post, float* singlets, int64_t ld_singlet, int n_singlets, cudaStream_t stream) {
    if (n_rows <= 0 || n_cols <= 0) return 0;
    if (!post && !singlets) return 0;
    softmax_rows_kernel<<<(unsigned)n_rows, 128, 0, stream>>>(logits, ld_logits, n_cols, post, ld_post, singlets,
                                                              ld_singlet, n_singlets);
    DMX_LAUNCH_CHECK();
    return 0;
}

static float doublet_bonus(int n_genotypes, double dp) {
    // demux.py:168-172 (float64, rounded to float32 on assignment)
    const double g = (double)n_genotypes;
    double bonus = log(g * dp);
    bonus -= log(g * (double)(n_genotypes - 1 > 1 ? n_genotypes - 1 : 1) / 2 * (1 - dp));
    return (float)bonus;
}

template <int FLAVOUR>
static int launch_singlets(int slots, unsigned grid, cudaStream_t stream, const int64_t* offsets,
                           const int32_t* variant, const float* e, const float* table, int64_t ld_table, int G,
                           const float* prior, int64_t ld_prior, float* logits, int64_t ld_logits) {
#define DMX_SINGLETS_CASE(S)                                                                                       \
    case S:                                                                                                        \
        estep_singlets_kernel<FLAVOUR, S><<<grid, 128, 0, stream>>>(offsets, variant, e, table, ld_table, G, prior, \
                                                                    ld_prior, logits, ld_logits);                 \
        break;
    switch (slots) {
        DMX_SINGLETS_CASE(1)
        DMX_SINGLETS_CASE(2)
        DMX_SINGLETS_CASE(4)
        DMX_SINGLETS_CASE(8)
        DMX_SINGLETS_CASE(16)
        default:
            set_error("singlet E-step supports up to 512 genotypes");
            return -2;
    }
#undef DMX_SINGLETS_CASE
    DMX_LAUNCH_CHECK();
    return 0;
}

}  // namespace dmx

extern "C" {

int64_t dmx_estep_workspace_bytes(int64_t n_barcodes, int32_t n_genotypes, double doublet_prior) {
    const int64_t cols = doublet_prior == 0 ? n_genotypes : (int64_t)n_genotypes * (n_genotypes + 1) / 2;
    return dmx::round_up(n_barcodes * cols * (int64_t)sizeof(float), 256) + 256;
}

int dmx_softmax_rows(const float* logits, int64_t ld_logits, int64_t n_rows, int32_t n_cols, float* posteriors,
                     int64_t ld_post, float* singlet_posteriors, int64_t ld_singlet, int32_t n_singlets,
                     void* stream) {
    return dmx::launch_softmax(logits, ld_logits, n_rows, n_cols, posteriors, ld_post, singlet_posteriors, ld_singlet,
                               n_singlets, (cudaStream_t)stream);
}

int dmx_estep(const int64_t* barcode_offsets, const int32_t* csr_variant, const float* csr_e, int64_t n_barcodes,
              const float* table, int64_t ld_table, int32_t n_genotypes, double doublet_prior,
              const float* prior_logits, int64_t ld_prior, float* logits, int64_t ld_logits, float* posteriors,
              int64_t ld_post, float* singlet_posteriors, int64_t ld_singlet, void* workspace,
              int64_t workspace_bytes, int32_t flavour, void* stream_) {
    using namespace dmx;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_barcodes <= 0) return 0;
    const int G = n_genotypes;
    DMX_REQUIRE(G >= 1, "n_genotypes must be positive");
    DMX_REQUIRE(doublet_prior >= 0 && doublet_prior < 1, "doublet_prior must be in [0, 1)");
    DMX_REQUIRE(flavour == DMX_ESTEP_EXACT || flavour == DMX_ESTEP_FAST, "unknown E-step flavour %d", flavour);
    DMX_REQUIRE(ld_table % 4 == 0 && ld_table >= G, "ld_table must be a multiple of 4 and >= n_genotypes");
    DMX_REQUIRE(((uintptr_t)table & 15) == 0, "table must be 16-byte aligned");
    const int64_t n_cols = doublet_prior == 0 ? G : (int64_t)G * (G + 1) / 2;
    DMX_REQUIRE(n_cols < (1ll << 31), "too many columns");

    float* out_logits = logits;
    int64_t ld_out = ld_logits;
    if (!out_logits) {
        DMX_REQUIRE(workspace && workspace_bytes >= n_barcodes * n_cols * (int64_t)sizeof(float),
                    "workspace too small for the logits scratch");
        out_logits = (float*)workspace;
        ld_out = n_cols;
    }

    if (doublet_prior == 0) {
        const int slots_needed = (int)ceil_div(G, 32);
        int slots = 1;
        while (slots < slots_needed) slots *= 2;
        int rc;
        if (flavour == DMX_ESTEP_FAST)
            rc = launch_singlets<DMX_ESTEP_FAST>(slots, (unsigned)n_barcodes, stream, barcode_offsets, csr_variant,
                                                 csr_e, table, ld_table, G, prior_logits, ld_prior, out_logits, ld_out);
        else
            rc = launch_singlets<DMX_ESTEP_EXACT>(slots, (unsigned)n_barcodes, stream, barcode_offsets, csr_variant,
                                                  csr_e, table, ld_table, G, prior_logits, ld_prior, out_logits, ld_out);
        if (rc) return rc;
    } else {
        PairsParams p;
        p.offsets = barcode_offsets;
        p.variant = csr_variant;
        p.e = csr_e;
        p.table = table;
        p.ld_table = ld_table;
        p.n_genotypes = G;
        p.gp = (int)round_up(G, TILE_I);
        const int q_total = p.gp / TILE_J;
        int n_tiles = 0;
        for (int pi = 0; pi < p.gp / TILE_I; ++pi) n_tiles += q_total - 2 * pi;
        p.n_tiles = n_tiles;
        if (n_tiles <= 128) {
            p.ctas_per_barcode = 1;
            p.tiles_per_cta = n_tiles;
            int rgs = 256 / n_tiles;
            p.row_groups = rgs < 1 ? 1 : (rgs > 16 ? 16 : rgs);
        } else {
            p.ctas_per_barcode = (int)ceil_div(n_tiles, 256);
            p.tiles_per_cta = (int)round_up(ceil_div(n_tiles, p.ctas_per_barcode), 32);
            p.row_groups = 1;
        }
        const int threads = p.tiles_per_cta * p.row_groups;
        DMX_REQUIRE(threads <= 256, "internal: CTA too large");
        p.ld_smem = p.gp + 4;
        // rows per staged chunk: at least FLUSH_ROWS per row group, more while the thread's copy list stays short
        p.flushes = 1;
        while (p.flushes < 4 &&
               (int64_t)p.row_groups * (p.flushes * 2) * FLUSH_ROWS * (p.gp / 4) <= (int64_t)8 * threads &&
               (int64_t)p.row_groups * (p.flushes * 2) * FLUSH_ROWS <= 128)
            p.flushes *= 2;
        const int64_t chunk_rows = (int64_t)p.row_groups * p.flushes * FLUSH_ROWS;
        DMX_REQUIRE(chunk_rows * (p.gp / 4) <= (int64_t)8 * threads,
                    "n_genotypes %d too large for the pair kernel's staging (max ~1000)", G);
        size_t smem = 2 * (size_t)chunk_rows * p.ld_smem * sizeof(float);
        if (p.row_groups > 1) smem += (size_t)p.tiles_per_cta * TILE_I * TILE_J * sizeof(double);
        DMX_REQUIRE(smem <= 200 * 1024, "shared memory request too large");
        p.doublet_bonus = doublet_bonus(G, doublet_prior);
        p.prior = prior_logits;
        p.ld_prior = ld_prior;
        p.logits = out_logits;
        p.ld_logits = ld_out;
        const int64_t grid = n_barcodes * p.ctas_per_barcode;
        DMX_REQUIRE(grid < (1ll << 31), "grid too large");
        if (flavour == DMX_ESTEP_FAST) {
            DMX_CUDA(cudaFuncSetAttribute(estep_pairs_kernel<DMX_ESTEP_FAST>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            estep_pairs_kernel<DMX_ESTEP_FAST><<<(unsigned)grid, threads, smem, stream>>>(p);
        } else {
            DMX_CUDA(cudaFuncSetAttribute(estep_pairs_kernel<DMX_ESTEP_EXACT>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            estep_pairs_kernel<DMX_ESTEP_EXACT><<<(unsigned)grid, threads, smem, stream>>>(p);
        }
        DMX_LAUNCH_CHECK();
    }
    return launch_softmax(out_logits, ld_out, n_barcodes, (int)n_cols, posteriors, ld_post, singlet_posteriors,
                          ld_singlet, G, stream);
}

}  // extern "C"
