// Pair E-step, warp-autonomous flavour (FAST arithmetic): one warp = one work item.  Two kernels: the warp kernel
// (9 <= G <= 56, a warp holds a barcode's whole pair triangle), the patch kernel (73 <= G <= 256) and the lane-per-row
// kernel (G <= 8), further down.
//
//   S_b[i, j] = sum_{rows r of barcode b} log( 0.5 (P[v_r, i] + P[v_r, j]) (1 - e_r) + max(e_r, 1e-4) ),  i <= j
//
// Same arithmetic as the FAST flavour of estep_pairs.cu (running float32 products of x = a_i + a_j, a =
// fma(P, 1 - e, e'); the binary exponents are moved into integer sums every FLUSH_ROWS rows; one lg2 per pair at
// the end), different decomposition:
//   * a work item is a barcode, or a segment of at most ~seg_rows rows of a deep barcode (dmx_estep_plan); items
//     are launched deepest barcode first, one 32-thread CTA each, so there is no CTA-wide barrier anywhere: a warp
//     stages its own table rows (cp.async, double buffered) and only ever executes __syncwarp;
//   * the upper triangle of the pair matrix is cut into 8 x 8 register tiles (64 running products per lane, 32
//     independent FADD2 -> FMUL2 chains): four LDS.128 feed 64 updates, half the shared-memory wavefronts per
//     update of the 4 x 4 tiles, which were co-limiting with the packed-FP32 pipe;
//   * lanes = RG row groups x T tiles (G = 32: 3 x 10); the row groups are reduced with shuffles in a fixed order;
//   * the exponent sums of two packed products share one register (2 x 16 bits): a lane flushes at most
//     seg_rows / FLUSH_ROWS + 1 times and a flushed product is < 2^17, so 16 bits cannot overflow for
//     seg_rows <= 4096 (checked by the launcher);
//   * segments of a multi-segment barcode write float64 log2-sums to a scratch matrix; the softmax kernel adds
//     them in segment order (deterministic) and produces the float32 logits.
#include "common.cuh"

namespace dmx {

constexpr float WARP_ERROR_FLOOR = 1e-4f;

static int warp_env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}

// outputs of the row softmax when the pair kernel takes it over for single-item barcodes (estep_pairs_strip.cu)
struct FusedSoftmax {
    float* post;
    int64_t ld_post;
    float* singlets;
    int64_t ld_singlet;
    bool logits_requested;  // false: the logits only live in the workspace, nobody reads the rows of fused barcodes
};

struct WarpPairsParams {
    const int64_t* offsets;     // barcode_offsets [B + 1]
    const int32_t* order;       // schedule slot -> barcode, or nullptr (identity)
    const int32_t* seg_prefix;  // [B + 1] by schedule slot: first item of the slot
    const int32_t* item_slot;   // [n_items]
    const int32_t* variant;
    const float* e;
    const float* table;
    int64_t ld_table;
    int n_genotypes;
    float doublet_bonus;
    const double* prior;
    int64_t ld_prior;
    float* logits;
    int64_t ld_logits;
    double* partial;  // [n_items, n_cols] log2-sums, written by the segments of multi-segment barcodes only
    int64_t n_cols;
    unsigned mant_mask, one_bits;  // 0x007fffff, 0x3f800000: kernel parameters so that they stay in registers
};

__device__ __forceinline__ uint64_t wpack2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void wunpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t wadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t wmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// (x & 0x007fffff) | 0x3f800000 as ONE LOP3 (the constants live in registers; with immediates ptxas emits two)
__device__ __forceinline__ unsigned wreset_mantissa(unsigned bits, unsigned mant_mask, unsigned one_bits) {
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(bits), "r"(mant_mask), "r"(one_bits));
    return d;
}
__device__ __forceinline__ float wlg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int NB>
struct WarpShape {
    static constexpr int GP = 8 * NB;               // genotypes padded to the tile width
    static constexpr int T = NB * (NB + 1) / 2;     // 8 x 8 tiles of the upper triangle (diagonal included)
    static constexpr int RG = 32 / T;               // row groups in the warp
    static constexpr int LD = GP + 4;               // floats per staged row (bank-conflict-free row stride)
    static constexpr int QUADS = GP / 4;            // 16-byte pieces per row
};

// NB          blocks of 8 genotypes (G <= 8 NB)
// FLUSH_ROWS  rows a lane multiplies into its products before the exponents are moved out (16, or 8 for tiny clips)
// SR          rows per row group in one staged chunk (FLUSH_ROWS % SR == 0); the row loop over SR is fully unrolled
// ESUM_SMEM   exponent sums live in shared memory (frees 32 registers -> more resident warps) instead of registers
// PREFETCH    explicit register double buffer for the operands of the next row
template <int NB, int FLUSH_ROWS, int SR, bool ESUM_SMEM, int MAX_REGS, bool PREFETCH>
__global__ void __maxnreg__(MAX_REGS) estep_pairs_warp_kernel(const WarpPairsParams p) {
    using S = WarpShape<NB>;
    constexpr int T = S::T, RG = S::RG, LD = S::LD, QUADS = S::QUADS;
    constexpr int CHUNK = RG * SR;                  // rows per staged chunk
    constexpr int CPF = FLUSH_ROWS / SR;            // chunks per flush
    static_assert(FLUSH_ROWS % SR == 0, "FLUSH_ROWS must be a multiple of SR");
    // staging slot = 1 / PIECES of a row; slot s of the chunk belongs to lane s % 32
    constexpr int PIECES = (NB % 2 == 0 && (CHUNK * 2) % 32 != 0) ? 4 : 2;
    constexpr int QPS = QUADS / PIECES;             // quads per slot
    static_assert(QUADS % PIECES == 0, "row does not split into equal slots");
    constexpr int RPP = 32 / PIECES;                // chunk rows covered by one pass of the warp
    constexpr int SPL = (CHUNK * PIECES + 31) / 32; // staging slots per lane and chunk
    constexpr int DUMP_LD = 33;                     // words per lane in the dump / exponent arrays (+1: no conflicts)
    constexpr int STAGE_FLOATS = 2 * CHUNK * LD;
    constexpr int DUMP_FLOATS = (ESUM_SMEM ? 1 : 2) * 32 * DUMP_LD;
    constexpr int SMEM_FLOATS = STAGE_FLOATS > DUMP_FLOATS ? STAGE_FLOATS : DUMP_FLOATS;
    __shared__ __align__(16) float smem[SMEM_FLOATS];
    __shared__ unsigned esum_s[ESUM_SMEM ? 32 * DUMP_LD : 1];
    float* const stage0 = smem;
    float* const stage1 = smem + CHUNK * LD;

    const int lane = threadIdx.x;
    const unsigned mant_mask = p.mant_mask, one_bits = p.one_bits;
    const int item = blockIdx.x;
    const int slot = __ldg(p.item_slot + item);
    const int seg_first = __ldg(p.seg_prefix + slot);
    const int n_seg = __ldg(p.seg_prefix + slot + 1) - seg_first;
    const int seg = item - seg_first;
    const int64_t barcode = p.order ? (int64_t)__ldg(p.order + slot) : (int64_t)slot;
    const int64_t b_lo = __ldg(p.offsets + barcode), b_hi = __ldg(p.offsets + barcode + 1);
    // segments: equal length, a multiple of the flush period, so only the last one carries padding rows
    constexpr int PERIOD = RG * FLUSH_ROWS;
    const int64_t per = ((b_hi - b_lo + n_seg - 1) / n_seg + PERIOD - 1) / PERIOD * PERIOD;
    int64_t row_lo = b_lo + (int64_t)seg * per;
    if (row_lo > b_hi) row_lo = b_hi;
    const int64_t row_hi = row_lo + per < b_hi ? row_lo + per : b_hi;
    const int n_chunks = (int)((row_hi - row_lo + CHUNK - 1) / CHUNK);

    // lane -> (row group, tile); lanes beyond RG * T shadow the last row group (they only stage)
    int rg = lane / T;
    const int tile = lane - rg * T;
    if (rg >= RG) rg = RG - 1;
    int ti = 0, tj = tile;  // tile -> (I, J), I <= J, I-major
    while (tj >= NB - ti) { tj -= NB - ti; ++ti; }
    tj += ti;

    uint64_t prod[4][8];                  // running products: [i pair][j], mantissas kept in [1, 2) by the flushes
    unsigned esum_r[ESUM_SMEM ? 1 : 4][8];  // 2 x 16-bit biased exponent sums per packed product (register flavour)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            prod[a][b] = wpack2(1.f, 1.f);
            if constexpr (ESUM_SMEM) esum_s[(a * 8 + b) * DUMP_LD + lane] = 0u;  // only ever touched by this lane
            else esum_r[a][b] = 0u;
        }

    // ---- staging ---------------------------------------------------------------------------------------------------
    const int piece = lane % PIECES;
    const int row0 = lane / PIECES;
    const int n_table_quads = (int)(p.ld_table / 4);
    int v_pre[SPL];
    float e_pre[SPL], e_cur[SPL];
    unsigned live = 0;

    auto prefetch = [&](int chunk) {
        const int64_t base = row_lo + (int64_t)chunk * CHUNK + row0;
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const int64_t row = base + RPP * k;
            v_pre[k] = -1;
            e_pre[k] = 0.f;
            if (row0 + RPP * k < CHUNK && row < row_hi) {
                v_pre[k] = __ldg(p.variant + row);
                e_pre[k] = __ldg(p.e + row);
            }
        }
    };
    auto issue = [&](float* buf) {
        live = 0;
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const int r = row0 + RPP * k;
            if (r < CHUNK) {
                float* dst = buf + r * LD + 4 * QPS * piece;
                e_cur[k] = e_pre[k];
                if (v_pre[k] >= 0) {
                    const float* src = p.table + (int64_t)v_pre[k] * p.ld_table + 4 * QPS * piece;
                    live |= 1u << k;
#pragma unroll
                    for (int u = 0; u < QPS; ++u) {
                        if (QPS * piece + u < n_table_quads) cp_async_16(dst + 4 * u, src + 4 * u);
                        else *reinterpret_cast<float4*>(dst + 4 * u) = make_float4(1.f, 1.f, 1.f, 1.f);
                    }
                } else {  // padding row: a = 1 -> factor 2 -> log2 = 1, removed in the epilogue
#pragma unroll
                    for (int u = 0; u < QPS; ++u) *reinterpret_cast<float4*>(dst + 4 * u) = make_float4(1.f, 1.f, 1.f, 1.f);
                }
            }
        }
        cp_async_commit();
    };
    auto land = [&](float* buf) {  // every lane finishes the pieces it copied itself
        cp_async_wait<0>();
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            if (live & (1u << k)) {
                float* dst = buf + (row0 + RPP * k) * LD + 4 * QPS * piece;
                const float e = e_cur[k];
                const float w = __fsub_rn(1.f, e);
                const float ef = fmaxf(e, WARP_ERROR_FLOOR);
                float4 x[QPS];  // all loads of the slot first: the fma chains then overlap
#pragma unroll
                for (int u = 0; u < QPS; ++u) x[u] = *reinterpret_cast<float4*>(dst + 4 * u);
#pragma unroll
                for (int u = 0; u < QPS; ++u) {
                    if (QPS * piece + u < n_table_quads) {
                        x[u].x = fmaf(x[u].x, w, ef);
                        x[u].y = fmaf(x[u].y, w, ef);
                        x[u].z = fmaf(x[u].z, w, ef);
                        x[u].w = fmaf(x[u].w, w, ef);
                        *reinterpret_cast<float4*>(dst + 4 * u) = x[u];
                    }
                }
            }
        }
    };

    if (n_chunks > 0) {
        prefetch(0);
        issue(stage0);
        if (n_chunks > 1) prefetch(1);
        land(stage0);
        __syncwarp();
    }

    int n_flushes = 0;
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        float* cur = (chunk & 1) ? stage1 : stage0;
        float* nxt = (chunk & 1) ? stage0 : stage1;
        const bool more = chunk + 1 < n_chunks;
        if (more) {
            issue(nxt);
            if (chunk + 2 < n_chunks) prefetch(chunk + 2);
        }

        // The row loop covers SR (8) rows per chunk, not the whole flush period: 16 rows x 64 packed instructions were
        // 25 KB of straight-line code with every warp of the SM at a different place in it, and instruction fetch
        // became the top stall (ncu: no_instruction, 81 % i-cache hit rate).
        const float* rows = cur + rg * LD;
        auto row_updates = [&](const float4& i_lo, const float4& i_hi, const float4& j_lo, const float4& j_hi) {
            const uint64_t ai[4] = {wpack2(i_lo.x, i_lo.y), wpack2(i_lo.z, i_lo.w), wpack2(i_hi.x, i_hi.y),
                                    wpack2(i_hi.z, i_hi.w)};
            const float aj[8] = {j_lo.x, j_lo.y, j_lo.z, j_lo.w, j_hi.x, j_hi.y, j_hi.z, j_hi.w};
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const uint64_t bj = wpack2(aj[b], aj[b]);
#pragma unroll
                for (int a = 0; a < 4; ++a) prod[a][b] = wmul2(prod[a][b], wadd2(ai[a], bj));
            }
        };
        if constexpr (PREFETCH) {  // operands of the next row are loaded before the current one is consumed
            float4 i_lo = *reinterpret_cast<const float4*>(rows + 8 * ti);
            float4 i_hi = *reinterpret_cast<const float4*>(rows + 8 * ti + 4);
            float4 j_lo = *reinterpret_cast<const float4*>(rows + 8 * tj);
            float4 j_hi = *reinterpret_cast<const float4*>(rows + 8 * tj + 4);
#pragma unroll
            for (int k = 0; k < SR; ++k) {
                const float* s = rows + (k + 1 < SR ? k + 1 : k) * (RG * LD);
                const float4 n_i_lo = *reinterpret_cast<const float4*>(s + 8 * ti);
                const float4 n_i_hi = *reinterpret_cast<const float4*>(s + 8 * ti + 4);
                const float4 n_j_lo = *reinterpret_cast<const float4*>(s + 8 * tj);
                const float4 n_j_hi = *reinterpret_cast<const float4*>(s + 8 * tj + 4);
                row_updates(i_lo, i_hi, j_lo, j_hi);
                i_lo = n_i_lo; i_hi = n_i_hi; j_lo = n_j_lo; j_hi = n_j_hi;
            }
        } else {
#pragma unroll
            for (int k = 0; k < SR; ++k) {
                const float* s = rows + k * (RG * LD);
                row_updates(*reinterpret_cast<const float4*>(s + 8 * ti), *reinterpret_cast<const float4*>(s + 8 * ti + 4),
                            *reinterpret_cast<const float4*>(s + 8 * tj), *reinterpret_cast<const float4*>(s + 8 * tj + 4));
            }
        }
        // renormalise every CPF chunks: exponents into the integer sums, mantissas back to [1, 2) (exact)
        if (CPF == 1 || (chunk + 1) % CPF == 0) {
            ++n_flushes;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    float lo, hi;
                    wunpack2(prod[a][b], lo, hi);
                    const unsigned blo = __float_as_uint(lo), bhi = __float_as_uint(hi);
                    if constexpr (ESUM_SMEM) {
                        esum_s[(a * 8 + b) * DUMP_LD + lane] += (blo >> 23) + ((bhi >> 23) << 16);
                    } else {
                        esum_r[a][b] += blo >> 23;          // LEA.HI
                        esum_r[a][b] += (bhi >> 23) << 16;  // SHF + LEA (the sign bit is clear)
                    }
                    prod[a][b] = wpack2(__uint_as_float(wreset_mantissa(blo, mant_mask, one_bits)),
                                        __uint_as_float(wreset_mantissa(bhi, mant_mask, one_bits)));
                }
        }

        if (more) land(nxt);
        __syncwarp();
    }

    // ---- epilogue -------------------------------------------------------------------------------------------------
    // Every lane dumps its (exponent sums, log2 of the products) into shared memory (the staging area is idle now);
    // then the warp walks the barcode's pairs as a rolled, lane-parallel loop: fixed-order sum over the row groups
    // (deterministic), penalty, prior, one rounding to float32.  Runs of 8 consecutive columns are written together.
    // The products may carry up to CPF - 1 unflushed chunks: lg2 of the whole float covers that.
    const int G = p.n_genotypes;
    const int bias = 127 * n_flushes;
    const double padded_rows = (double)n_chunks * (double)CHUNK;
    unsigned* const dump_e = ESUM_SMEM ? esum_s : reinterpret_cast<unsigned*>(smem + 32 * DUMP_LD);
    float* const dump_l = smem;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                float lo, hi;
                wunpack2(prod[a][b], lo, hi);
                dump_l[(a * 8 + b) * DUMP_LD + lane] = wlg2(h ? hi : lo);
                if constexpr (!ESUM_SMEM)
                    if (h == 0) dump_e[(a * 8 + b) * DUMP_LD + lane] = esum_r[a][b];
            }
        __syncwarp();
#pragma unroll 1
        for (int idx = lane; idx < T * 32; idx += 32) {
            const int t = idx >> 5, q = idx & 31;
            const int a = q >> 3, b = q & 7;
            int oi = 0, oj = t;
            while (oj >= NB - oi) { oj -= NB - oi; ++oi; }
            oj += oi;
            const int i = 8 * oi + 2 * a + h, j = 8 * oj + b;
            if (i < G && j < G && j >= i) {
                double sum = 0.0;
#pragma unroll
                for (int g = 0; g < RG; ++g) {
                    const int src = (a * 8 + b) * DUMP_LD + g * T + t;
                    const unsigned e2 = dump_e[src];
                    const int ev = (int)(h ? e2 >> 16 : e2 & 0xffffu) - bias;
                    sum += (double)ev + (double)dump_l[src];
                }
                sum -= padded_rows;
                const int64_t col = (i == j) ? i : (int64_t)G + (int64_t)i * G - (int64_t)i * (i + 1) / 2 + (j - i - 1);
                if (n_seg == 1) {
                    const float pen = (i == j) ? 0.f : p.doublet_bonus;
                    float logit = (float)((double)pen + sum * 0.693147180559945309417232);
                    if (p.prior) logit = (float)((double)logit + p.prior[barcode * p.ld_prior + col]);
                    p.logits[barcode * p.ld_logits + col] = logit;
                } else {
                    p.partial[(int64_t)item * p.n_cols + col] = sum;
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Many genotypes (G = 200: 325 tiles of 8 x 8): a warp cannot hold a barcode's whole triangle, so a work item is a
// (barcode segment, PATCH) pair and a warp owns the 32 tiles of one patch.  Tiles are enumerated in bands of 4 tile
// rows, column by column inside a band, and cut into consecutive runs of 32: a patch then touches at most 14 of the
// 8-genotype blocks (4-8 row blocks + a contiguous run of column blocks), and only those operand blocks of each table
// row are staged (compacted: block k of the patch's ascending block list sits at floats [8k, 8k + 8) of the staged
// row).  Patches of one item are adjacent in the grid, so the warps that gather the same table rows run at the same
// time and share them through L2.  Everything else -- arithmetic, flush, staging protocol, epilogue -- is the
// warp kernel above with one row group.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PATCH_BAND = 4;        // tile rows per band
constexpr int PATCH_MAX_BLOCKS = 16; // operand blocks a patch may touch (checked by the launcher)
constexpr int PATCH_LD = PATCH_MAX_BLOCKS * 8 + 4;

// tile number g (band-major, column-major inside a band) -> (I, J)
__host__ __device__ inline void patch_tile(int nb, int g, int* ti, int* tj) {
    for (int i0 = 0; i0 < nb; i0 += PATCH_BAND) {
        const int rows = nb - i0 < PATCH_BAND ? nb - i0 : PATCH_BAND;
        const int cols = nb - i0;
        // column c of the band (J = i0 + c) holds min(rows, c + 1) tiles
        const int head = rows * (rows - 1) / 2;                           // tiles in the first rows - 1 columns
        const int in_band = cols >= rows ? head + rows * (cols - rows + 1) : cols * (cols + 1) / 2;
        if (g < in_band) {
            int c = 0;
            if (g < head) {
                while (g >= c + 1) { g -= c + 1; ++c; }
            } else {
                g -= head;
                c = rows - 1 + g / rows;
                g -= (c - (rows - 1)) * rows;
            }
            *ti = i0 + g;
            *tj = i0 + c;
            return;
        }
        g -= in_band;
    }
    *ti = *tj = nb - 1;
}

// CPF: staged chunks per flush.  2: the operands are staged times 4 (exact) so that 32 factors in [0.077, 8.001] keep a
// product normal (needs table entries >= 0.0095; the default clip is 0.01) -- half the flush instructions; the flush is
// exact, so its period does not change any result bit.
template <int FLUSH_ROWS, int CPF, int MAX_REGS = 168>
__global__ void __maxnreg__(MAX_REGS) estep_pairs_patch_kernel(const WarpPairsParams p, int nb, int n_tiles, int n_patches) {
    constexpr int CHUNK = FLUSH_ROWS;  // one row group: rows per staged chunk (<= 16)
    constexpr float SCALE = CPF == 2 ? 4.f : 1.f;
    constexpr int LD = PATCH_LD;
    constexpr int DUMP_LD = 33;
    constexpr int STAGE_FLOATS = 2 * CHUNK * LD;
    constexpr int SMEM_FLOATS = STAGE_FLOATS > 2 * 32 * DUMP_LD ? STAGE_FLOATS : 2 * 32 * DUMP_LD;
    __shared__ __align__(16) float smem[SMEM_FLOATS];
    __shared__ int tile_ij[32];
    float* const stage0 = smem;
    float* const stage1 = smem + CHUNK * LD;

    const int lane = threadIdx.x;
    const unsigned mant_mask = p.mant_mask, one_bits = p.one_bits;
    const int item = blockIdx.x / n_patches;
    const int patch = blockIdx.x - item * n_patches;
    const int slot = __ldg(p.item_slot + item);
    const int seg_first = __ldg(p.seg_prefix + slot);
    const int n_seg = __ldg(p.seg_prefix + slot + 1) - seg_first;
    const int seg = item - seg_first;
    const int64_t barcode = p.order ? (int64_t)__ldg(p.order + slot) : (int64_t)slot;
    const int64_t b_lo = __ldg(p.offsets + barcode), b_hi = __ldg(p.offsets + barcode + 1);
    const int64_t per = ((b_hi - b_lo + n_seg - 1) / n_seg + CHUNK - 1) / CHUNK * CHUNK;
    int64_t row_lo = b_lo + (int64_t)seg * per;
    if (row_lo > b_hi) row_lo = b_hi;
    const int64_t row_hi = row_lo + per < b_hi ? row_lo + per : b_hi;
    const int n_chunks = (int)((row_hi - row_lo + CHUNK - 1) / CHUNK);

    // lane -> tile of the patch (lanes past the last tile shadow it: they stage and compute, but write nothing)
    const int first_tile = patch * 32;
    const int g = first_tile + lane < n_tiles ? first_tile + lane : n_tiles - 1;
    int ti, tj;
    patch_tile(nb, g, &ti, &tj);
    tile_ij[lane] = first_tile + lane < n_tiles ? (ti | (tj << 8)) : -1;
    const unsigned mask = __reduce_or_sync(0xffffffffu, (1u << ti) | (1u << tj));  // operand blocks of the patch
    const int n_blocks = __popc(mask);
    const int pos_i = __popc(mask & ((1u << ti) - 1u));  // where the lane's blocks sit in the staged rows
    const int pos_j = __popc(mask & ((1u << tj) - 1u));

    uint64_t prod[4][8];
    unsigned esum[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            prod[a][b] = wpack2(1.f, 1.f);
            esum[a][b] = 0u;
        }

    // ---- staging: lane = (chunk row lane >> 1, half lane & 1); half 0 stages the first ceil(n / 2) blocks of the
    // patch's ascending block list, half 1 the rest.  At most PATCH_MAX_BLOCKS / 2 blocks per lane: the loops below
    // are straight-line predicated code with constant shared-memory offsets (the first version walked the whole list
    // in a dynamic loop per chunk and spent more instructions on staging than on anything but the products).
    constexpr int MAXB = PATCH_MAX_BLOCKS / 2;
    const int row_in_chunk = lane >> 1;
    const int half = lane & 1;
    const int n_first = (n_blocks + 1) >> 1;
    unsigned my_mask = mask;
    {
        unsigned first = 0, m = mask;
        for (int k = 0; k < n_first; ++k) { first |= m & (0u - m); m &= m - 1; }
        my_mask = half ? (mask & ~first) : first;
    }
    const int my_count = half ? n_blocks - n_first : n_first;
    const int k0 = half ? n_first : 0;
    const int full_blocks = (int)(p.ld_table / 8);  // blocks whose two quads both lie inside the table row
    const int n_table_quads = (int)(p.ld_table / 4);
    int v_pre = -1;
    float e_pre = 0.f, e_cur = 0.f;
    bool live = false;

    auto prefetch = [&](int chunk) {
        const int64_t row = row_lo + (int64_t)chunk * CHUNK + row_in_chunk;
        v_pre = -1;
        e_pre = 0.f;
        if (row_in_chunk < CHUNK && row < row_hi) {
            v_pre = __ldg(p.variant + row);
            e_pre = __ldg(p.e + row);
        }
    };
    auto issue = [&](float* buf) {
        live = false;
        if (row_in_chunk < CHUNK) {
            float* dst = buf + row_in_chunk * LD + 8 * k0;
            e_cur = e_pre;
            live = v_pre >= 0;
            const float* src_row = p.table + (int64_t)(live ? v_pre : 0) * p.ld_table;
            unsigned m = my_mask;
#pragma unroll
            for (int t = 0; t < MAXB; ++t) {
                if (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const float* src = src_row + 8 * b;
                    if (live && b < full_blocks) {
                        cp_async_16(dst + 8 * t, src);
                        cp_async_16(dst + 8 * t + 4, src + 4);
                    } else {
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            if (live && 2 * b + u < n_table_quads) cp_async_16(dst + 8 * t + 4 * u, src + 4 * u);
                            else *reinterpret_cast<float4*>(dst + 8 * t + 4 * u) = make_float4(SCALE, SCALE, SCALE, SCALE);
                        }
                    }
                }
            }
        }
        cp_async_commit();
    };
    auto land = [&](float* buf) {
        cp_async_wait<0>();
        if (live) {  // columns past the table width hold 1 and get transformed too: no pair that uses them is written
            float* dst = buf + row_in_chunk * LD + 8 * k0;
            const float w = __fmul_rn(__fsub_rn(1.f, e_cur), SCALE);
            const float ef = __fmul_rn(fmaxf(e_cur, WARP_ERROR_FLOOR), SCALE);
            // two blocks per batch: all loads of a batch first (the operand registers of the row loop are dead
            // here), so the fma chains do not each wait for their own shared-memory round trip
#pragma unroll
            for (int t0 = 0; t0 < MAXB; t0 += 2) {
                float4 x[4];
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (t0 + t < my_count) {
                        x[2 * t] = *reinterpret_cast<float4*>(dst + 8 * (t0 + t));
                        x[2 * t + 1] = *reinterpret_cast<float4*>(dst + 8 * (t0 + t) + 4);
                    }
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (t0 + t < my_count) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            float4& v = x[2 * t + u];
                            v.x = fmaf(v.x, w, ef); v.y = fmaf(v.y, w, ef); v.z = fmaf(v.z, w, ef); v.w = fmaf(v.w, w, ef);
                            *reinterpret_cast<float4*>(dst + 8 * (t0 + t) + 4 * u) = v;
                        }
                    }
            }
        }
    };

    if (n_chunks > 0) {
        prefetch(0);
        issue(stage0);
        if (n_chunks > 1) prefetch(1);
        land(stage0);
    }
    __syncwarp();

    int n_flushes = 0;
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        float* cur = (chunk & 1) ? stage1 : stage0;
        float* nxt = (chunk & 1) ? stage0 : stage1;
        const bool more = chunk + 1 < n_chunks;
        if (more) {
            issue(nxt);
            if (chunk + 2 < n_chunks) prefetch(chunk + 2);
        }
        const float* oi = cur + 8 * pos_i;
        const float* oj = cur + 8 * pos_j;
        float4 i_lo = *reinterpret_cast<const float4*>(oi);
        float4 i_hi = *reinterpret_cast<const float4*>(oi + 4);
        float4 j_lo = *reinterpret_cast<const float4*>(oj);
        float4 j_hi = *reinterpret_cast<const float4*>(oj + 4);
#pragma unroll
        for (int k = 0; k < CHUNK; ++k) {
            const int kn = k + 1 < CHUNK ? k + 1 : k;
            const float4 n_i_lo = *reinterpret_cast<const float4*>(oi + kn * LD);
            const float4 n_i_hi = *reinterpret_cast<const float4*>(oi + kn * LD + 4);
            const float4 n_j_lo = *reinterpret_cast<const float4*>(oj + kn * LD);
            const float4 n_j_hi = *reinterpret_cast<const float4*>(oj + kn * LD + 4);
            const uint64_t ai[4] = {wpack2(i_lo.x, i_lo.y), wpack2(i_lo.z, i_lo.w), wpack2(i_hi.x, i_hi.y),
                                    wpack2(i_hi.z, i_hi.w)};
            const float aj[8] = {j_lo.x, j_lo.y, j_lo.z, j_lo.w, j_hi.x, j_hi.y, j_hi.z, j_hi.w};
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const uint64_t bj = wpack2(aj[b], aj[b]);
#pragma unroll
                for (int a = 0; a < 4; ++a) prod[a][b] = wmul2(prod[a][b], wadd2(ai[a], bj));
            }
            i_lo = n_i_lo; i_hi = n_i_hi; j_lo = n_j_lo; j_hi = n_j_hi;
        }
        if (CPF == 1 || (chunk & 1)) {
            ++n_flushes;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    float lo, hi;
                    wunpack2(prod[a][b], lo, hi);
                    const unsigned blo = __float_as_uint(lo), bhi = __float_as_uint(hi);
                    esum[a][b] += blo >> 23;
                    esum[a][b] += (bhi >> 23) << 16;
                    prod[a][b] = wpack2(__uint_as_float(wreset_mantissa(blo, mant_mask, one_bits)),
                                        __uint_as_float(wreset_mantissa(bhi, mant_mask, one_bits)));
                }
        }
        if (more) land(nxt);
        __syncwarp();
    }

    // ---- epilogue (see the warp kernel): dump, then a rolled lane-parallel walk over the patch's pairs -------------
    const int G = p.n_genotypes;
    const int bias = 127 * n_flushes;
    // every staged row carries the 2 of the pair sum and SCALE: log2(2 SCALE) = 3 per row when SCALE = 4
    const double padded_rows = (double)n_chunks * (double)CHUNK * (CPF == 2 ? 3.0 : 1.0);
    unsigned* const dump_e = reinterpret_cast<unsigned*>(smem + 32 * DUMP_LD);
    float* const dump_l = smem;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                float lo, hi;
                wunpack2(prod[a][b], lo, hi);
                dump_l[(a * 8 + b) * DUMP_LD + lane] = wlg2(h ? hi : lo);
                if (h == 0) dump_e[(a * 8 + b) * DUMP_LD + lane] = esum[a][b];
            }
        __syncwarp();
#pragma unroll 1
        for (int idx = lane; idx < 32 * 32; idx += 32) {
            const int t = idx >> 5, q = idx & 31;
            const int a = q >> 3, b = q & 7;
            const int ij = tile_ij[t];
            if (ij < 0) continue;
            const int i = 8 * (ij & 0xff) + 2 * a + h, j = 8 * (ij >> 8) + b;
            if (i < G && j < G && j >= i) {
                const int src = (a * 8 + b) * DUMP_LD + t;
                const unsigned e2 = dump_e[src];
                const int ev = (int)(h ? e2 >> 16 : e2 & 0xffffu) - bias;
                const double sum = (double)ev + (double)dump_l[src] - padded_rows;
                const int64_t col = (i == j) ? i : (int64_t)G + (int64_t)i * G - (int64_t)i * (i + 1) / 2 + (j - i - 1);
                if (n_seg == 1) {
                    const float pen = (i == j) ? 0.f : p.doublet_bonus;
                    float logit = (float)((double)pen + sum * 0.693147180559945309417232);
                    if (p.prior) logit = (float)((double)logit + p.prior[barcode * p.ld_prior + col]);
                    p.logits[barcode * p.ld_logits + col] = logit;
                } else {
                    p.partial[(int64_t)item * p.n_cols + col] = sum;
                }
            }
        }
        __syncwarp();
    }
}

// patches of the band-major tile order: how many, and the largest operand-block list any of them needs
static void patch_shape(int nb, int* n_tiles, int* n_patches, int* max_blocks) {
    *n_tiles = nb * (nb + 1) / 2;
    *n_patches = (*n_tiles + 31) / 32;
    *max_blocks = 0;
    for (int p = 0; p < *n_patches; ++p) {
        unsigned mask = 0;
        for (int l = 0; l < 32 && p * 32 + l < *n_tiles; ++l) {
            int ti, tj;
            patch_tile(nb, p * 32 + l, &ti, &tj);
            mask |= (1u << ti) | (1u << tj);
        }
        int bits = 0;
        for (unsigned m = mask; m; m &= m - 1) ++bits;
        if (bits > *max_blocks) *max_blocks = bits;
    }
}

static bool patch_kernel_supported(int G) {
    const int nb = (G + 7) / 8;
    // measured against the CTA kernel (scripts/sweep_estep_patch.py, profiles/r01_sweep_patch_*.log):
    // G = 104: 40.1 vs 30.2 updates/clk/SM, G = 128: 37.2 vs 27.2, G = 200: 39.9 vs 24.7
    const int min_nb = warp_env_int("DMX_PAIRS_PATCH_MIN_NB", 9);
    if (nb < min_nb || nb < 9 || nb > 32) return false;
    if (warp_env_int("DMX_PAIRS_PATCH", 1) == 0) return false;
    int n_tiles, n_patches, max_blocks;
    patch_shape(nb, &n_tiles, &n_patches, &max_blocks);
    // per cent of the lanes that must carry tiles; at 69-70 % (G = 72, 88) the patch kernel still leads the CTA kernel,
    // 31.0 against 27.7 / 29.2 updates/clk/SM
    const int min_eff = warp_env_int("DMX_PAIRS_PATCH_MIN_EFF", 65);
    return max_blocks <= PATCH_MAX_BLOCKS && 100 * n_tiles >= min_eff * 32 * n_patches;
}

// ---------------------------------------------------------------------------------------------------------------
// Few genotypes (G <= 8, at most 36 columns; the reference's own example has 4): a row feeds too few pairs for
// register tiles shared across lanes, and the kernel is bound by the row stream, not by arithmetic.  Here a LANE owns
// whole rows: lane l of the item's warp takes rows l, l + 32, ... (coalesced record loads, one or two 128-bit gathers
// of the table row), keeps all GT (GT + 1) / 2 running products of its rows in registers with the same exponent
// bookkeeping, and the 32 lanes are combined at the end with an xor-shuffle tree in float64 (fixed order:
// deterministic).  Rows past the end multiply by exactly 1.  Same work items as the other warp kernels.
// ---------------------------------------------------------------------------------------------------------------
// SINGLETS: doublet_prior == 0 -- only the G singlet columns, factor a_g itself (no pair sum, no factor 2).
// EXACT: the reference's own roundings (demux.py:188-190, 261): p = fl(fl(P_i + P_j) * 0.5), x = fl(fl(p (1 - e)) + e'),
// t = logf(x), float64 accumulation, so that the float32 logits carry the reference's rounding noise instead of an
// independent one (posterior bar 1e-6).  Not free: 24 M rows, G = 4: 0.107 -> 0.142 ms (singlets), 0.122 -> 0.267 ms
// (pairs); G = 8: 0.159 -> 0.247 ms (singlets), 0.214 -> 1.515 ms (pairs: 36 logf and 36 float64 sums per row and
// lane).  DMX_ESTEP_AUTO picks it for the singlet columns only.
template <int GT, int FLUSH_ROWS, bool SINGLETS, bool EXACT>
__global__ void __launch_bounds__(32) estep_pairs_small_kernel(const WarpPairsParams p) {
    constexpr int CT = SINGLETS ? GT : GT * (GT + 1) / 2;
    constexpr int WAVES = 4;  // rows in flight per lane
    const int lane = threadIdx.x;
    const int item = blockIdx.x;
    const int slot = __ldg(p.item_slot + item);
    const int seg_first = __ldg(p.seg_prefix + slot);
    const int n_seg = __ldg(p.seg_prefix + slot + 1) - seg_first;
    const int seg = item - seg_first;
    const int64_t barcode = p.order ? (int64_t)__ldg(p.order + slot) : (int64_t)slot;
    const int64_t b_lo = __ldg(p.offsets + barcode), b_hi = __ldg(p.offsets + barcode + 1);
    const int64_t per = ((b_hi - b_lo + n_seg - 1) / n_seg + 31) / 32 * 32;
    int64_t row_lo = b_lo + (int64_t)seg * per;
    if (row_lo > b_hi) row_lo = b_hi;
    const int64_t row_hi = row_lo + per < b_hi ? row_lo + per : b_hi;

    float prod[EXACT ? 1 : CT];
    int esum[EXACT ? 1 : CT];
    double acc[EXACT ? CT : 1];
#pragma unroll
    for (int c = 0; c < (EXACT ? 1 : CT); ++c) { prod[c] = 1.f; esum[c] = 0; }
#pragma unroll
    for (int c = 0; c < (EXACT ? CT : 1); ++c) acc[c] = 0.0;
    int n_flushes = 0, since_flush = 0;

    for (int64_t base = row_lo + lane; base < row_hi; base += 32 * WAVES) {  // lanes run out of rows at different times
        float a[WAVES][GT];
        float w_row[WAVES], ef_row[WAVES];
        bool ok_row[WAVES];
#pragma unroll
        for (int u = 0; u < WAVES; ++u) {
            const int64_t row = base + 32 * u;
            const bool ok = row < row_hi;
            const int32_t v = ok ? __ldg(p.variant + row) : 0;
            const float e = ok ? __ldg(p.e + row) : 0.f;
            const float w = __fsub_rn(1.f, e);
            const float ef = fmaxf(e, WARP_ERROR_FLOOR);
            w_row[u] = w; ef_row[u] = ef; ok_row[u] = ok;
            const float4* src = reinterpret_cast<const float4*>(p.table + (int64_t)v * p.ld_table);
#pragma unroll
            for (int q = 0; q < (GT + 3) / 4; ++q) {
                float4 t = make_float4(1.f, 1.f, 1.f, 1.f);
                if (ok && 4 * q < p.ld_table) t = __ldg(src + q);
                const float x4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int k = 0; k < 4; ++k)  // rows past the end are neutral: 0.5 + 0.5 = 1 (pairs), 1 (singlets)
                    if (4 * q + k < GT) {
                        if constexpr (EXACT) a[u][4 * q + k] = x4[k];  // the table entry itself
                        else a[u][4 * q + k] = ok ? fmaf(x4[k], w, ef) : (SINGLETS ? 1.f : 0.5f);
                    }
            }
        }
        if constexpr (EXACT) {
#pragma unroll
            for (int u = 0; u < WAVES; ++u) {
                if (!ok_row[u]) continue;
                int c = 0;
#pragma unroll
                for (int i = 0; i < GT; ++i)
#pragma unroll
                    for (int j = i; j < (SINGLETS ? i + 1 : GT); ++j, ++c) {
                        const float pc = (i == j) ? a[u][i] : __fmul_rn(__fadd_rn(a[u][i], a[u][j]), 0.5f);
                        acc[c] += (double)logf(__fadd_rn(__fmul_rn(pc, w_row[u]), ef_row[u]));
                    }
            }
        } else {
#pragma unroll
        for (int u = 0; u < WAVES; ++u) {
            int c = 0;
#pragma unroll
            for (int i = 0; i < GT; ++i) {
                if constexpr (SINGLETS) {
                    prod[i] = __fmul_rn(prod[i], a[u][i]);
                } else {
#pragma unroll
                    for (int j = i; j < GT; ++j, ++c) prod[c] = __fmul_rn(prod[c], __fadd_rn(a[u][i], a[u][j]));
                }
            }
        }
        }
        since_flush += WAVES;
        if (!EXACT && since_flush >= FLUSH_ROWS) {  // lane-local row count: uniform across the warp's active lanes
            since_flush = 0;
            ++n_flushes;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const unsigned bits = __float_as_uint(prod[c]);
                esum[c] += (int)(bits >> 23);
                prod[c] = __uint_as_float((bits & 0x007fffffu) | 0x3f800000u);
            }
        }
    }

    // ---- epilogue: per-lane log2, float64 xor-tree over the lanes, penalties, rounding ------------------------------
    const int G = p.n_genotypes;
    const double real_rows = SINGLETS ? 0.0 : (double)(row_hi - row_lo);  // a real row's pair factor carries a 2
    int c = 0;
#pragma unroll
    for (int i = 0; i < GT; ++i)
#pragma unroll
        for (int j = i; j < (SINGLETS ? i + 1 : GT); ++j, ++c) {
            double v;
            if constexpr (EXACT) v = acc[c];  // natural-log sum
            else v = (double)(esum[c] - 127 * n_flushes) + (double)wlg2(prod[c]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == (c & 31) && i < G && j < G) {
                const double sum = EXACT ? v : v - real_rows;
                const int64_t col = (i == j) ? i : (int64_t)G + (int64_t)i * G - (int64_t)i * (i + 1) / 2 + (j - i - 1);
                if (n_seg == 1) {
                    const float pen = (i == j) ? 0.f : p.doublet_bonus;
                    float logit = (float)((double)pen + (EXACT ? sum : sum * 0.693147180559945309417232));
                    if (p.prior) logit = (float)((double)logit + p.prior[barcode * p.ld_prior + col]);
                    p.logits[barcode * p.ld_logits + col] = logit;
                } else {
                    p.partial[(int64_t)item * p.n_cols + col] = sum;
                }
            }
        }
}

// ---- work items ---------------------------------------------------------------------------------------------------

__global__ void plan_segments_kernel(const int64_t* __restrict__ offsets, const int32_t* __restrict__ order,
                                     int64_t n_barcodes, int seg_rows, int32_t* __restrict__ n_seg) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s <= n_barcodes;
         s += (int64_t)gridDim.x * blockDim.x) {
        int32_t n = 0;
        if (s < n_barcodes) {
            const int64_t b = order ? (int64_t)order[s] : s;
            const int64_t rows = offsets[b + 1] - offsets[b];
            n = (int32_t)(rows <= seg_rows ? 1 : (rows + seg_rows - 1) / seg_rows);
        }
        n_seg[s] = n;
    }
}

__global__ void plan_items_kernel(const int32_t* __restrict__ seg_prefix, int64_t n_barcodes, int64_t capacity,
                                  int32_t* __restrict__ item_slot) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n_barcodes;
         s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lo = seg_prefix[s], hi = seg_prefix[s + 1];
        for (int64_t k = lo; k < hi && k < capacity; ++k) item_slot[k] = (int32_t)s;
    }
}

static inline int plan_grid(int64_t n, int threads) {
    int64_t blocks = ceil_div(n > 0 ? n : 1, threads);
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(blocks < cap ? blocks : cap);
}

int launch_plan_segments(const int64_t* offsets, const int32_t* order, int64_t n_barcodes, int seg_rows,
                         int32_t* n_seg, cudaStream_t stream) {
    plan_segments_kernel<<<plan_grid(n_barcodes + 1, 256), 256, 0, stream>>>(offsets, order, n_barcodes, seg_rows, n_seg);
    DMX_LAUNCH_CHECK();
    return 0;
}

int launch_plan_items(const int32_t* seg_prefix, int64_t n_barcodes, int64_t capacity, int32_t* item_slot,
                      cudaStream_t stream) {
    plan_items_kernel<<<plan_grid(n_barcodes, 256), 256, 0, stream>>>(seg_prefix, n_barcodes, capacity, item_slot);
    DMX_LAUNCH_CHECK();
    return 0;
}

float pair_doublet_bonus(int n_genotypes, double dp) {  // demux.py:168-172
    const double g = (double)n_genotypes;
    double bonus = log(g * dp);
    bonus -= log(g * (double)(n_genotypes - 1 > 1 ? n_genotypes - 1 : 1) / 2 * (1 - dp));
    return (float)bonus;
}

// estep_pairs_strip.cu
bool estep_pairs_strip_supported(int G);
int launch_estep_pairs_strip(const int64_t* barcode_offsets, const int32_t* barcode_order, const int32_t* seg_prefix,
                             const int32_t* item_slot, int64_t n_items, int seg_rows, const int32_t* csr_variant,
                             const float* csr_e, const float* table, int64_t ld_table, int G, double doublet_prior,
                             float table_floor, const double* prior_logits, int64_t ld_prior, float* logits,
                             int64_t ld_logits, double* partial, int64_t n_cols, float* post, int64_t ld_post,
                             float* singlets, int64_t ld_singlet, cudaStream_t stream);

bool estep_pairs_warp_supported(int G, int flavour) {
    if (warp_env_int("DMX_PAIRS_WARP", 1) == 0) return false;
    const int nb = (G + 7) / 8;
    if (G <= 8) return warp_env_int("DMX_PAIRS_SMALL", 1) != 0;  // lane-per-row kernel, both flavours
    if (flavour != DMX_ESTEP_FAST) return false;
    if (G > 16 && estep_pairs_strip_supported(G)) return true;  // strip kernel: 25..32 and 57..64 genotypes
    if (nb == 2) return true;                                    // 3 tiles x 10 row groups
    if (nb == 3 || nb == 4 || nb == 5 || nb == 7) return true;  // lane utilisation >= 28 / 32
    return patch_kernel_supported(G);                            // other widths: estep_pairs.cu
}

template <int NB, int FLUSH_ROWS, int SR, bool ESUM_SMEM, int MAX_REGS, bool PREFETCH>
static int launch_warp_variant(const WarpPairsParams& p, int64_t n_items, cudaStream_t stream) {
    auto kernel = estep_pairs_warp_kernel<NB, FLUSH_ROWS, SR, ESUM_SMEM, MAX_REGS, PREFETCH>;
    DMX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
    kernel<<<(unsigned)n_items, 32, 0, stream>>>(p);
    DMX_LAUNCH_CHECK();
    return 0;
}

int launch_estep_pairs_warp(const int64_t* barcode_offsets, const int32_t* barcode_order, const int32_t* seg_prefix,
                            const int32_t* item_slot, int64_t n_items, int seg_rows, const int32_t* csr_variant,
                            const float* csr_e, const float* table, int64_t ld_table, int G, double doublet_prior,
                            float table_floor, const double* prior_logits, int64_t ld_prior, float* logits,
                            int64_t ld_logits, double* partial, int64_t n_cols, int flavour, const FusedSoftmax* fused,
                            int* did_fuse, cudaStream_t stream) {
    if (did_fuse) *did_fuse = 0;
    DMX_REQUIRE(seg_rows >= 16 && seg_rows <= 4096, "seg_rows %d outside [16, 4096]", seg_rows);
    DMX_REQUIRE(n_items > 0 && n_items < (1ll << 31), "bad item count %lld", (long long)n_items);
    WarpPairsParams p;
    p.offsets = barcode_offsets;
    p.order = barcode_order;
    p.seg_prefix = seg_prefix;
    p.item_slot = item_slot;
    p.variant = csr_variant;
    p.e = csr_e;
    p.table = table;
    p.ld_table = ld_table;
    p.n_genotypes = G;
    p.doublet_bonus = doublet_prior == 0 ? 0.f : pair_doublet_bonus(G, doublet_prior);
    p.prior = prior_logits;
    p.ld_prior = ld_prior;
    p.logits = logits;
    p.ld_logits = ld_logits;
    p.partial = partial;
    p.n_cols = n_cols;
    p.mant_mask = 0x007fffffu;
    p.one_bits = 0x3f800000u;
    // 16 factors per product need 16 * -log2(2 (floor + 1e-4)) <= 120 binades; 8 factors are always safe
    const bool long_products = table_floor >= 0.0027f && warp_env_int("DMX_FLUSH_ROWS", 16) == 16;
    const int nb = (G + 7) / 8;
    if (G <= 8) {
        // 16 (8) factors in [2 (floor + 1e-4), 2.0002] per product between flushes, as in the tiled kernels; singlet
        // factors are only >= 1e-4: always 8
#define DMX_SMALL(GT_)                                                                                                \
        if (flavour == DMX_ESTEP_EXACT) {                                                                             \
            if (doublet_prior == 0) estep_pairs_small_kernel<GT_, 8, true, true><<<(unsigned)n_items, 32, 0, stream>>>(p); \
            else estep_pairs_small_kernel<GT_, 8, false, true><<<(unsigned)n_items, 32, 0, stream>>>(p);              \
        }                                                                                                             \
        else if (doublet_prior == 0) estep_pairs_small_kernel<GT_, 8, true, false><<<(unsigned)n_items, 32, 0, stream>>>(p);  \
        else if (long_products) estep_pairs_small_kernel<GT_, 16, false, false><<<(unsigned)n_items, 32, 0, stream>>>(p); \
        else estep_pairs_small_kernel<GT_, 8, false, false><<<(unsigned)n_items, 32, 0, stream>>>(p)
        if (G <= 2) { DMX_SMALL(2); }
        else if (G <= 4) { DMX_SMALL(4); }
        else { DMX_SMALL(8); }
#undef DMX_SMALL
        DMX_LAUNCH_CHECK();
        return 0;
    }
    // 9 <= G <= 16: 3 tiles x 10 row groups, 8-row chunks with a flush per chunk (not FP32 bound at this width).  Measured
    // at G = 16, 20 M rows: 0.47 ms against 0.81 ms for the CTA kernel and 0.73 ms for a lane-pair-per-row kernel (dropped)
    if (G <= 16) return launch_warp_variant<2, 8, 8, false, 168, true>(p, n_items, stream);
    if (estep_pairs_strip_supported(G)) {
        // the warp that owns a single-item barcode also takes its softmax (fused), and then only writes the logits when
        // the caller asked for them
        const bool fuse = fused && (fused->post || fused->singlets) && warp_env_int("DMX_FUSE_SOFTMAX", 1) != 0;
        if (fuse && did_fuse) *did_fuse = 1;
        return launch_estep_pairs_strip(barcode_offsets, barcode_order, seg_prefix, item_slot, n_items, seg_rows,
                                        csr_variant, csr_e, table, ld_table, G, doublet_prior, table_floor, prior_logits,
                                        ld_prior, fuse && !fused->logits_requested ? nullptr : logits, ld_logits, partial,
                                        n_cols, fuse ? fused->post : nullptr, fuse ? fused->ld_post : 0,
                                        fuse ? fused->singlets : nullptr, fuse ? fused->ld_singlet : 0, stream);
    }
    const int variant = warp_env_int("DMX_WARP_VARIANT", 0);  // experiments: exponent packing / occupancy target
#define DMX_WARP(NB_, SR_, ESM_, REGS_, PF_)                                                                   \
    return long_products ? launch_warp_variant<NB_, 16, SR_, ESM_, REGS_, PF_>(p, n_items, stream)            \
                         : launch_warp_variant<NB_, 8, 8, ESM_, REGS_, PF_>(p, n_items, stream)
    // measured on B200 at G = 32 (profiles/r01_sweep_warp_*.log): one 16-row chunk per flush and 12 resident warps per
    // SM (168 registers, exponent sums in registers) 1.08 ms; 8-row chunks 1.19 ms (twice the chunk hand-overs);
    // 16 warps per SM with the exponent sums in shared memory (128 registers) 1.17 ms
    switch (nb) {
        case 3: DMX_WARP(3, 16, false, 168, true);
        case 4:
            if (variant == 1) { DMX_WARP(4, 8, false, 168, true); }
            if (variant == 2) { DMX_WARP(4, 8, true, 128, true); }
            DMX_WARP(4, 16, false, 168, true);
        case 5: DMX_WARP(5, 16, false, 168, true);
        case 7: DMX_WARP(7, 16, false, 168, true);
        default: break;
    }
#undef DMX_WARP
    if (patch_kernel_supported(G)) {
        int n_tiles, n_patches, max_blocks;
        patch_shape(nb, &n_tiles, &n_patches, &max_blocks);
        const int64_t grid = n_items * n_patches;
        DMX_REQUIRE(grid < (1ll << 31), "grid too large");
        const bool very_long = long_products && table_floor >= 0.0095f && warp_env_int("DMX_PATCH_PERIOD", 16) >= 32;  // opt-in: ptxas spills
        if (very_long && warp_env_int("DMX_PATCH_REGS", 168) == 184) {
            auto kernel = estep_pairs_patch_kernel<16, 2, 184>;
            DMX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int)cudaSharedmemCarveoutMaxShared));
            kernel<<<(unsigned)grid, 32, 0, stream>>>(p, nb, n_tiles, n_patches);
        } else if (very_long) {
            auto kernel = estep_pairs_patch_kernel<16, 2>;
            DMX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int)cudaSharedmemCarveoutMaxShared));
            kernel<<<(unsigned)grid, 32, 0, stream>>>(p, nb, n_tiles, n_patches);
        } else if (long_products) {
            auto kernel = estep_pairs_patch_kernel<16, 1>;
            DMX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int)cudaSharedmemCarveoutMaxShared));
            kernel<<<(unsigned)grid, 32, 0, stream>>>(p, nb, n_tiles, n_patches);
        } else {
            auto kernel = estep_pairs_patch_kernel<8, 1>;
            DMX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int)cudaSharedmemCarveoutMaxShared));
            kernel<<<(unsigned)grid, 32, 0, stream>>>(p, nb, n_tiles, n_patches);
        }
        DMX_LAUNCH_CHECK();
        return 0;
    }
    set_error("warp pair kernel does not support %d genotypes", G);
    return -2;
}

}  // namespace dmx
