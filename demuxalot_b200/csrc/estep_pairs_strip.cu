// Pair E-step, "strip" flavour of the warp-autonomous kernel (FAST arithmetic) for 25..32 and 57..64 genotypes.
//
//   S_b[i, j] = sum_{rows r of barcode b} log( 0.5 (P[v_r, i] + P[v_r, j]) (1 - e_r) + max(e_r, 1e-4) ),  i <= j
//
// Same arithmetic, staging protocol, flush and work items as estep_pairs_warp.cu; what changes is how the pair triangle
// is cut up.  The 8 x 8 register tiles of the warp kernel waste the lower halves of the diagonal tiles (528 of 640
// product slots useful at G = 32, and 30 of 32 lanes carry tiles), and -- measured with scripts/microbench_packed_tile.cu
// -- a packed FADD2 / FMUL2 holds the issue port for two cycles while every other instruction takes one more, so the
// cost of a row is 2 x (packed instructions) + (everything else): wasted product slots cost exactly as much as useful
// ones.  Here the unit of work is a STRIP: one pair of genotypes (a packed register pair) against one block of 8
// genotypes (broadcast operands), 8 packed products.  An off-diagonal tile is 4 strips, in either orientation (pair
// from the row block, broadcast over the column block, or transposed); a diagonal block is two HALVES of 12 packed
// products (pair 2h against d[0..8), pair 3 - h ... against d[4..8): 20 of 24 slots useful).  Every lane executes the
// same instruction stream:
//       strips 0, 1 :  pair(p0), pair(p1)  x  block Q1         16 packed products
//       strip  2    :  pair(p2)            x  block D           8
//       unit   3    :  pair(P1) x D[0..8),  pair(P2) x D[4..8)  12   (a diagonal half of block D, or a fourth strip)
// with lane-specific shared-memory offsets, 36 packed products per lane and row:
//   G <= 32: 8 slots x 4 row groups = 32 lanes, 528 of 576 slots useful, 72 instead of 85.3 packed instructions per
//            row and warp, and 5.5 instead of 5.3 operand wavefronts per row (the third strip of a slot is chosen so that
//            its block is the slot's diagonal block, and the diagonal pairs are selected from the block registers);
//   G <= 64: 32 slots x 1 row group, 2080 of 2304 slots useful -- the width that used to fall back to the CTA kernel.
// The assignment of strips to slots is computed on the host (strip_layout) and travels as a kernel parameter.
#include "common.cuh"

namespace dmx {

constexpr float STRIP_ERROR_FLOOR = 1e-4f;

struct StripSlot {
    int16_t p0, p1, p2;  // float offset of the pair of strips 0..2 inside a staged row
    int16_t q1;          // block of strips 0 and 1
    int16_t d;           // block of strip 2 and of unit 3
    int16_t P1, P2;      // pairs of unit 3 (P2 < 0: unit 3 is a plain strip, its second part is idle)
    int16_t pad;
};

struct StripLayout {
    StripSlot slot[32];
};

struct StripParams {
    const int64_t* offsets;
    const int32_t* order;
    const int32_t* seg_prefix;
    const int32_t* item_slot;
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
    double* partial;
    int64_t n_cols;
    unsigned mant_mask, one_bits;
    // fused row softmax (demux.py:101,152) for barcodes that are a single work item: the warp that owns the barcode
    // keeps its float32 logits in shared memory and writes the posteriors itself
    float* post;       // [B, ld_post] or nullptr
    int64_t ld_post;
    float* singlets;   // [B, ld_singlet] or nullptr
    int64_t ld_singlet;
    int fuse_softmax;
};

__device__ __forceinline__ uint64_t spack2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void sunpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t smul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned sreset_mantissa(unsigned bits, unsigned mant_mask, unsigned one_bits) {
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(bits), "r"(mant_mask), "r"(one_bits));
    return d;
}
__device__ __forceinline__ float slg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int STRIP_NP = 36;  // packed products per lane

// (i, j), i <= j, of element `el` (0 / 1) of packed product q of a slot; false for idle product slots
__device__ __forceinline__ bool strip_pair_of(const StripSlot& s, int q, int el, int* i, int* j) {
    int pair, other;
    if (q < 16) { pair = (q < 8 ? s.p0 : s.p1); other = s.q1 + (q & 7); }
    else if (q < 24) { pair = s.p2; other = s.d + (q & 7); }
    else if (q < 32) { pair = s.P1; other = s.d + (q & 7); }
    else { pair = s.P2; other = s.d + 4 + (q & 3); }
    if (pair < 0) return false;
    const int x = pair + el;
    const bool diagonal_unit = q >= 24 && (pair >> 3) == (s.d >> 3);  // a pair taken from its own block
    if (diagonal_unit && other < x) return false;                      // lower triangle of a diagonal block
    *i = x < other ? x : other;
    *j = x < other ? other : x;
    return true;
}

// NB     blocks of 8 genotypes (4 or 8);  RG row groups, 32 / RG slots
// FLUSH  rows a lane multiplies into its products between exponent flushes (= rows per row group and staged chunk)
// SELECT_DIAG_PAIRS: unit 3 of every slot is a diagonal half whose pairs sit in the block registers at positions
//        (2h, 2h + 1) and (6 - 2h, 7 - 2h): selected with predicated moves instead of two more shared-memory loads
// FULL_COLS: the table row is exactly 8 NB floats wide (no padding columns to fill in while staging)
// CPF    staged chunks per flush (2: the operands are staged times 4 so that 2 FLUSH factors keep a product normal)
// UNROLL rows of the row loop that are unrolled (the loop body is 72 packed instructions per row: code size vs loop overhead)
template <int NB, int RG, int FLUSH, bool SELECT_DIAG_PAIRS, int MAX_REGS, bool FULL_COLS, int CPF, int UNROLL>
__global__ void __maxnreg__(MAX_REGS) estep_pairs_strip_kernel(const StripParams p, const StripLayout layout) {
    constexpr int GP = 8 * NB;
    constexpr int LD = GP + 4;
    constexpr int QUADS = GP / 4;
    constexpr int SLOTS = 32 / RG;
    constexpr int CHUNK = RG * FLUSH;
    constexpr int PIECES = 2;                       // a lane stages half a row at a time
    constexpr int QPS = QUADS / PIECES;
    constexpr int RPP = 32 / PIECES;
    constexpr int SPL = (CHUNK * PIECES + 31) / 32;
    constexpr int DUMP_LD = 33;
    constexpr float SCALE = CPF == 2 ? 4.f : 1.f;   // operands are staged times SCALE (exact), removed in the epilogue
    constexpr int STAGE_FLOATS = 2 * CHUNK * LD;
    constexpr int C_MAX = GP * (GP + 1) / 2;        // logits of one barcode (fused softmax)
    constexpr int DUMP_FLOATS = 2 * STRIP_NP * DUMP_LD + C_MAX;
    constexpr int SMEM_FLOATS = STAGE_FLOATS > DUMP_FLOATS ? STAGE_FLOATS : DUMP_FLOATS;
    __shared__ __align__(16) float smem[SMEM_FLOATS];
    float* const stage0 = smem;
    float* const stage1 = smem + CHUNK * LD;

    const int lane = threadIdx.x;
    const unsigned mant_mask = p.mant_mask, one_bits = p.one_bits;
    const int item = blockIdx.x;
    const int slot_of_item = __ldg(p.item_slot + item);
    const int seg_first = __ldg(p.seg_prefix + slot_of_item);
    const int n_seg = __ldg(p.seg_prefix + slot_of_item + 1) - seg_first;
    const int seg = item - seg_first;
    const int64_t barcode = p.order ? (int64_t)__ldg(p.order + slot_of_item) : (int64_t)slot_of_item;
    const int64_t b_lo = __ldg(p.offsets + barcode), b_hi = __ldg(p.offsets + barcode + 1);
    const int64_t per = ((b_hi - b_lo + n_seg - 1) / n_seg + CHUNK - 1) / CHUNK * CHUNK;
    int64_t row_lo = b_lo + (int64_t)seg * per;
    if (row_lo > b_hi) row_lo = b_hi;
    const int64_t row_hi = row_lo + per < b_hi ? row_lo + per : b_hi;
    const int n_chunks = (int)((row_hi - row_lo + CHUNK - 1) / CHUNK);

    const int rg = lane / SLOTS;
    const StripSlot my = layout.slot[lane % SLOTS];
    const int o_p0 = my.p0, o_p1 = my.p1, o_p2 = my.p2, o_q1 = my.q1, o_d = my.d;
    const int o_P1 = my.P1, o_P2 = my.P2 < 0 ? my.P1 : my.P2;  // an idle second part multiplies harmless values
    const bool upper_half = SELECT_DIAG_PAIRS && ((my.P1 & 7) != 0);  // h = 1: pairs (2, 3) and (4, 5) of the block

    uint64_t prod[STRIP_NP];
    unsigned esum[STRIP_NP];  // 2 x 16-bit biased exponent sums per packed product
#pragma unroll
    for (int q = 0; q < STRIP_NP; ++q) {
        prod[q] = spack2(1.f, 1.f);
        esum[q] = 0u;
    }

    // ---- staging (as in the warp kernel): lane = (row of the pass, half row) --------------------------------------
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
                        if (FULL_COLS || QPS * piece + u < n_table_quads) cp_async_16(dst + 4 * u, src + 4 * u);
                        else *reinterpret_cast<float4*>(dst + 4 * u) = make_float4(SCALE, SCALE, SCALE, SCALE);
                    }
                } else {  // padding row: a = SCALE -> factor 2 SCALE, removed in the epilogue
#pragma unroll
                    for (int u = 0; u < QPS; ++u) *reinterpret_cast<float4*>(dst + 4 * u) = make_float4(SCALE, SCALE, SCALE, SCALE);
                }
            }
        }
        cp_async_commit();
    };
    // chunks that lie entirely inside the item (all but the last one) with full-width table rows: no predicates at all
    auto issue_full = [&](float* buf) {
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            float* dst = buf + (row0 + RPP * k) * LD + 4 * QPS * piece;
            const float* src = p.table + (int64_t)v_pre[k] * p.ld_table + 4 * QPS * piece;
            e_cur[k] = e_pre[k];
#pragma unroll
            for (int u = 0; u < QPS; ++u) cp_async_16(dst + 4 * u, src + 4 * u);
        }
        cp_async_commit();
    };
    auto land_full = [&](float* buf) {
        cp_async_wait<0>();
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            float* dst = buf + (row0 + RPP * k) * LD + 4 * QPS * piece;
            const float e = e_cur[k];
            const float w = __fmul_rn(__fsub_rn(1.f, e), SCALE);
            const float ef = __fmul_rn(fmaxf(e, STRIP_ERROR_FLOOR), SCALE);
            float4 x[QPS];
#pragma unroll
            for (int u = 0; u < QPS; ++u) x[u] = *reinterpret_cast<float4*>(dst + 4 * u);
#pragma unroll
            for (int u = 0; u < QPS; ++u) {
                x[u].x = fmaf(x[u].x, w, ef);
                x[u].y = fmaf(x[u].y, w, ef);
                x[u].z = fmaf(x[u].z, w, ef);
                x[u].w = fmaf(x[u].w, w, ef);
                *reinterpret_cast<float4*>(dst + 4 * u) = x[u];
            }
        }
    };
    constexpr bool FAST_STAGING = FULL_COLS && (CHUNK * PIECES) % 32 == 0;
    auto land = [&](float* buf) {  // every lane finishes the pieces it copied itself
        cp_async_wait<0>();
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            if (live & (1u << k)) {
                float* dst = buf + (row0 + RPP * k) * LD + 4 * QPS * piece;
                const float e = e_cur[k];
                const float w = __fmul_rn(__fsub_rn(1.f, e), SCALE);
                const float ef = __fmul_rn(fmaxf(e, STRIP_ERROR_FLOOR), SCALE);
                float4 x[QPS];
#pragma unroll
                for (int u = 0; u < QPS; ++u) x[u] = *reinterpret_cast<float4*>(dst + 4 * u);
#pragma unroll
                for (int u = 0; u < QPS; ++u) {
                    if (FULL_COLS || QPS * piece + u < n_table_quads) {
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
        // the next chunk is full when it is not the item's last one (the last one may carry padding rows)
        const bool next_full = FAST_STAGING && chunk + 2 < n_chunks;
        if (more) {
            if (next_full) issue_full(nxt); else issue(nxt);
            if (chunk + 2 < n_chunks) prefetch(chunk + 2);
        }

        // Row loop, software-pipelined by hand: a row's products come in two phases -- strips 0, 1 against block Q1, then
        // strip 2 and unit 3 against block D -- and the operands of each phase are fetched from shared memory while the
        // other phase computes (the registers they land in have just been consumed), so no load sits in front of its
        // first use and no second register set is needed (ncu on the first version: short_scoreboard was the top stall).
        const float* rows = cur + rg * LD;
        float2 a0 = *reinterpret_cast<const float2*>(rows + o_p0);
        float2 a1 = *reinterpret_cast<const float2*>(rows + o_p1);
        float4 q_lo = *reinterpret_cast<const float4*>(rows + o_q1);
        float4 q_hi = *reinterpret_cast<const float4*>(rows + o_q1 + 4);
#pragma unroll UNROLL
        for (int k = 0; k < FLUSH; ++k) {
            const float* s = rows + k * (RG * LD);
            const float2 a2 = *reinterpret_cast<const float2*>(s + o_p2);
            const float4 d_lo = *reinterpret_cast<const float4*>(s + o_d);
            const float4 d_hi = *reinterpret_cast<const float4*>(s + o_d + 4);
            float2 b1, b2;
            if constexpr (!SELECT_DIAG_PAIRS) {
                b1 = *reinterpret_cast<const float2*>(s + o_P1);
                b2 = *reinterpret_cast<const float2*>(s + o_P2);
            }
            {   // phase A
                const uint64_t pr0 = spack2(a0.x, a0.y), pr1 = spack2(a1.x, a1.y);
                const float qv[8] = {q_lo.x, q_lo.y, q_lo.z, q_lo.w, q_hi.x, q_hi.y, q_hi.z, q_hi.w};
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const uint64_t qb = spack2(qv[b], qv[b]);
                    prod[b] = smul2(prod[b], sadd2(pr0, qb));
                    prod[8 + b] = smul2(prod[8 + b], sadd2(pr1, qb));
                }
            }
            if (k + 1 < FLUSH) {  // operands of the next row's phase A
                const float* n = s + RG * LD;
                a0 = *reinterpret_cast<const float2*>(n + o_p0);
                a1 = *reinterpret_cast<const float2*>(n + o_p1);
                q_lo = *reinterpret_cast<const float4*>(n + o_q1);
                q_hi = *reinterpret_cast<const float4*>(n + o_q1 + 4);
            }
            {   // phase B
                if constexpr (SELECT_DIAG_PAIRS) {
                    b1 = upper_half ? make_float2(d_lo.z, d_lo.w) : make_float2(d_lo.x, d_lo.y);
                    b2 = upper_half ? make_float2(d_hi.x, d_hi.y) : make_float2(d_hi.z, d_hi.w);
                }
                const uint64_t pr2 = spack2(a2.x, a2.y), pu1 = spack2(b1.x, b1.y), pu2 = spack2(b2.x, b2.y);
                const float dv[8] = {d_lo.x, d_lo.y, d_lo.z, d_lo.w, d_hi.x, d_hi.y, d_hi.z, d_hi.w};
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const uint64_t db = spack2(dv[b], dv[b]);
                    prod[16 + b] = smul2(prod[16 + b], sadd2(pr2, db));
                    prod[24 + b] = smul2(prod[24 + b], sadd2(pu1, db));
                    if (b >= 4) prod[32 + b - 4] = smul2(prod[32 + b - 4], sadd2(pu2, db));
                }
            }
        }
        // renormalise every CPF chunks: exponents into the integer sums, mantissas back to [1, 2) (exact, so the
        // flush period does not change any result bit)
        if (CPF == 1 || (chunk & 1)) {
            ++n_flushes;
#pragma unroll
        for (int q = 0; q < STRIP_NP; ++q) {
            float lo, hi;
            sunpack2(prod[q], lo, hi);
            const unsigned blo = __float_as_uint(lo), bhi = __float_as_uint(hi);
            esum[q] += blo >> 23;
            esum[q] += (bhi >> 23) << 16;
            prod[q] = spack2(__uint_as_float(sreset_mantissa(blo, mant_mask, one_bits)),
                             __uint_as_float(sreset_mantissa(bhi, mant_mask, one_bits)));
        }
        }
        if (more) {
            if (next_full) land_full(nxt); else land(nxt);
        }
        __syncwarp();
    }

    // ---- epilogue: dump (exponent sums, log2 of the products), then a rolled, lane-parallel walk over the slots' pairs --
    const int G = p.n_genotypes;
    const int bias = 127 * n_flushes;
    // every staged row (real or padding) carries the factor 2 of the pair sum and SCALE twice... once: x = SCALE (a_i + a_j)
    const double padded_rows = (double)n_chunks * (double)CHUNK * (CPF == 2 ? 3.0 : 1.0);
    float* const dump_l = smem;
    unsigned* const dump_e = reinterpret_cast<unsigned*>(smem + STRIP_NP * DUMP_LD);
    float* const logits_s = smem + 2 * STRIP_NP * DUMP_LD;
    const bool fuse = p.fuse_softmax && n_seg == 1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int q = 0; q < STRIP_NP; ++q) {
            float lo, hi;
            sunpack2(prod[q], lo, hi);
            dump_l[q * DUMP_LD + lane] = slg2(h ? hi : lo);
            if (h == 0) dump_e[q * DUMP_LD + lane] = esum[q];
        }
        __syncwarp();
#pragma unroll 1
        for (int idx = lane; idx < SLOTS * STRIP_NP; idx += 32) {
            const int sl = idx / STRIP_NP, q = idx - sl * STRIP_NP;
            int i, j;
            if (!strip_pair_of(layout.slot[sl], q, h, &i, &j)) continue;
            if (i >= G || j >= G) continue;
            double sum = 0.0;
#pragma unroll
            for (int g = 0; g < RG; ++g) {
                const int src = q * DUMP_LD + g * SLOTS + sl;
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
                if (p.logits) p.logits[barcode * p.ld_logits + col] = logit;
                if (fuse) logits_s[col] = logit;
            } else {
                p.partial[(int64_t)item * p.n_cols + col] = sum;
            }
        }
        __syncwarp();
    }
    if (fuse) {  // same arithmetic as softmax_rows_kernel (estep.cu): float32 max and exp, float64 sum, float32 divide
        const int n_cols = (int)p.n_cols;
        float m = -3.402823466e+38f;
        for (int c = lane; c < n_cols; c += 32) m = fmaxf(m, logits_s[c]);
        m = warp_max(m);
        double sum = 0.0;
        for (int c = lane; c < n_cols; c += 32) sum += (double)expf(__fsub_rn(logits_s[c], m));
        sum = warp_sum(sum);
        const float total = (float)sum;
        for (int c = lane; c < n_cols; c += 32) {
            const float pr = __fdiv_rn(expf(__fsub_rn(logits_s[c], m)), total);
            if (p.post) p.post[barcode * p.ld_post + c] = pr;
            if (p.singlets && c < G) p.singlets[barcode * p.ld_singlet + c] = pr;
        }
    }
}

// ---- assignment of strips to slots (host) ---------------------------------------------------------------------------
// Tiles (I, J), I < J, are cut into 4 strips each, all in one orientation: normal = pair 8 I + 2 a against block J,
// transposed = pair 8 J + 2 a against block I.  A slot takes one PAIR UNIT (two strips of a tile: strips 0, 1 and their
// common block Q1), a third strip whose block is the slot's block D, and as unit 3 either a half of the diagonal block
// D (G <= 32: every slot; G <= 64: 16 of the 32 slots) or a fourth strip against D.
static bool strip_layout(int nb, StripLayout* out) {
    struct Strip { int pair, block; };
    auto strip_of = [](int I, int J, bool transposed, int a) {
        return transposed ? Strip{8 * J + 2 * a, 8 * I} : Strip{8 * I + 2 * a, 8 * J};
    };
    const int n_slots = nb == 4 ? 8 : 32;
    int filled01[32] = {0}, filled2[32] = {0}, filled3[32] = {0};
    // slots with a diagonal half: block D = slot / 2, half h = slot & 1 (both widths: the first 2 nb slots)
    for (int s = 0; s < n_slots; ++s) {
        StripSlot& sl = out->slot[s];
        sl = StripSlot{-1, -1, -1, -1, -1, -1, -1, 0};
        if (s < 2 * nb) {
            const int D = s / 2, h = s & 1;
            sl.d = (int16_t)(8 * D);
            sl.P1 = (int16_t)(8 * D + 2 * h);          // pair h against d[0..8)
            sl.P2 = (int16_t)(8 * D + (h ? 4 : 6));    // pair 2 (h = 1) or 3 (h = 0) against d[4..8)
            filled3[s] = 1;
        }
    }
    // the tile that lends block D to the third strips of the two slots of diagonal block D: (D, D + 1) transposed, and
    // (0, nb - 1) normal for the last block
    bool lender[32][32] = {};
    int n_pair_units = 0;
    Strip pair_units[64][2];
    for (int D = 0; D < nb; ++D) {
        const int I = D + 1 < nb ? D : 0, J = D + 1 < nb ? D + 1 : nb - 1;
        const bool transposed = D + 1 < nb;
        lender[I][J] = true;
        for (int h = 0; h < 2; ++h) {
            const Strip st = strip_of(I, J, transposed, 2 + h);
            StripSlot& sl = out->slot[2 * D + h];
            if (st.block != sl.d) return false;
            sl.p2 = (int16_t)st.pair;
            filled2[2 * D + h] = 1;
        }
        pair_units[n_pair_units][0] = strip_of(I, J, transposed, 0);
        pair_units[n_pair_units][1] = strip_of(I, J, transposed, 1);
        ++n_pair_units;
    }
    for (int I = 0; I < nb; ++I)
        for (int J = I + 1; J < nb; ++J) {
            if (lender[I][J]) continue;
            for (int u = 0; u < 2; ++u) {
                pair_units[n_pair_units][0] = strip_of(I, J, false, 2 * u);
                pair_units[n_pair_units][1] = strip_of(I, J, false, 2 * u + 1);
                ++n_pair_units;
            }
        }
    // every slot takes one pair unit as strips 0, 1; slots without a diagonal half take a second one as strip 2 + unit 3
    int next = 0;
    for (int s = 0; s < n_slots; ++s) {
        if (next >= n_pair_units) return false;
        StripSlot& sl = out->slot[s];
        sl.p0 = (int16_t)pair_units[next][0].pair;
        sl.p1 = (int16_t)pair_units[next][1].pair;
        sl.q1 = (int16_t)pair_units[next][0].block;
        filled01[s] = 1;
        ++next;
    }
    for (int s = 0; s < n_slots; ++s) {
        if (filled3[s]) continue;
        if (next >= n_pair_units) return false;
        StripSlot& sl = out->slot[s];
        sl.p2 = (int16_t)pair_units[next][0].pair;
        sl.P1 = (int16_t)pair_units[next][1].pair;
        sl.P2 = -1;
        sl.d = (int16_t)pair_units[next][0].block;
        filled2[s] = filled3[s] = 1;
        ++next;
    }
    if (next != n_pair_units) return false;
    for (int s = 0; s < n_slots; ++s)
        if (!filled01[s] || !filled2[s] || !filled3[s]) return false;
    return true;
}

}  // namespace dmx

extern "C" int dmx_estep_strip_layout(int32_t n_blocks, int16_t* h_slots, int32_t* h_n_slots) {
    dmx::StripLayout layout;
    DMX_REQUIRE(n_blocks == 4 || n_blocks == 8, "the strip kernel serves 4 or 8 blocks of 8 genotypes, not %d", (int)n_blocks);
    DMX_REQUIRE(dmx::strip_layout(n_blocks, &layout), "no strip layout for %d blocks", (int)n_blocks);
    const int n_slots = n_blocks == 4 ? 8 : 32;
    for (int s = 0; s < n_slots; ++s) {
        const dmx::StripSlot& sl = layout.slot[s];
        const int16_t fields[8] = {sl.p0, sl.p1, sl.p2, sl.q1, sl.d, sl.P1, sl.P2, 0};
        for (int k = 0; k < 8; ++k) h_slots[8 * s + k] = fields[k];
    }
    *h_n_slots = n_slots;
    return 0;
}

namespace dmx {

bool estep_pairs_strip_supported(int G) {
    const char* v = getenv("DMX_PAIRS_STRIP");
    if (v && *v && atoi(v) == 0) return false;
    const int nb = (G + 7) / 8;
    return nb == 4 || nb == 8;
}

float pair_doublet_bonus(int n_genotypes, double dp);  // estep_pairs_warp.cu

template <int NB, int RG, int FLUSH, bool SEL, int MAX_REGS, bool FULL_COLS, int CPF, int UNROLL = FLUSH>
static int launch_strip_variant(const StripParams& p, const StripLayout& layout, int64_t n_items, cudaStream_t stream) {
    auto kernel = estep_pairs_strip_kernel<NB, RG, FLUSH, SEL, MAX_REGS, FULL_COLS, CPF, UNROLL>;
    DMX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
    kernel<<<(unsigned)n_items, 32, 0, stream>>>(p, layout);
    DMX_LAUNCH_CHECK();
    return 0;
}

int launch_estep_pairs_strip(const int64_t* barcode_offsets, const int32_t* barcode_order, const int32_t* seg_prefix,
                             const int32_t* item_slot, int64_t n_items, int seg_rows, const int32_t* csr_variant,
                             const float* csr_e, const float* table, int64_t ld_table, int G, double doublet_prior,
                             float table_floor, const double* prior_logits, int64_t ld_prior, float* logits,
                             int64_t ld_logits, double* partial, int64_t n_cols, float* post, int64_t ld_post,
                             float* singlets, int64_t ld_singlet, cudaStream_t stream) {
    DMX_REQUIRE(seg_rows >= 16 && seg_rows <= 4096, "seg_rows %d outside [16, 4096]", seg_rows);
    DMX_REQUIRE(n_items > 0 && n_items < (1ll << 31), "bad item count %lld", (long long)n_items);
    const int nb = (G + 7) / 8;
    static StripLayout layouts[2];
    static bool ready[2] = {false, false};
    const int which = nb == 4 ? 0 : 1;
    if (!ready[which]) {
        DMX_REQUIRE(strip_layout(nb, &layouts[which]), "no strip layout for %d blocks", nb);
        ready[which] = true;
    }
    StripParams p;
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
    p.post = post;
    p.ld_post = ld_post;
    p.singlets = singlets;
    p.ld_singlet = ld_singlet;
    p.fuse_softmax = (post || singlets) ? 1 : 0;
    // factors lie in [2 (floor + 1e-4), 2.0002]: 8 of them always keep a product normal, 16 need floor >= 0.0027, and 32
    // (operands staged times 4: factors in [0.077, 8.001]) need floor >= 0.0095 -- the default clip is 0.01
    const int period_env = getenv("DMX_STRIP_PERIOD") ? atoi(getenv("DMX_STRIP_PERIOD")) : 32;
    const int period = table_floor >= 0.0095f && period_env >= 32 ? 32 : (table_floor >= 0.0027f && period_env >= 16 ? 16 : 8);
    const bool full = ld_table == 8 * nb;
    // measured on B200 (profiles/r02_sweep_strip.log): G = 32: 168 registers (12 warps / SM), 8 rows unrolled 36.6
    // updates/clk/SM (16 rows unrolled 34.7: instruction fetch; 184 / 200 registers 33.6 / 33.8: occupancy);
    // G = 64: 200 registers, 8 rows unrolled 37.0 (168 registers 33.8)
    const int unroll = getenv("DMX_STRIP_UNROLL") ? atoi(getenv("DMX_STRIP_UNROLL")) : 8;
    const int regs = getenv("DMX_STRIP_REGS") ? atoi(getenv("DMX_STRIP_REGS")) : (nb == 4 ? 168 : 200);
#define DMX_STRIP_V(NB_, RG_, SEL_, REGS_, UNROLL_)                                                                   \
    if (regs == REGS_ && unroll == UNROLL_)                                                                           \
        return launch_strip_variant<NB_, RG_, 16, SEL_, REGS_, true, 2, UNROLL_>(p, layouts[which], n_items, stream)
#define DMX_STRIP(NB_, RG_, SEL_)                                                                                     \
    do {                                                                                                              \
        if (period == 32) {                                                                                           \
            if (full) {                                                                                               \
                DMX_STRIP_V(NB_, RG_, SEL_, 168, 8);                                                                  \
                DMX_STRIP_V(NB_, RG_, SEL_, 184, 8);                                                                  \
                DMX_STRIP_V(NB_, RG_, SEL_, 200, 8);                                                                  \
                DMX_STRIP_V(NB_, RG_, SEL_, 184, 16);                                                                 \
                DMX_STRIP_V(NB_, RG_, SEL_, 200, 16);                                                                 \
                return launch_strip_variant<NB_, RG_, 16, SEL_, 168, true, 2, 16>(p, layouts[which], n_items, stream); \
            }                                                                                                         \
            return launch_strip_variant<NB_, RG_, 16, SEL_, 168, false, 2>(p, layouts[which], n_items, stream);       \
        }                                                                                                             \
        if (period == 16) {                                                                                           \
            if (full) return launch_strip_variant<NB_, RG_, 16, SEL_, 168, true, 1>(p, layouts[which], n_items, stream);  \
            return launch_strip_variant<NB_, RG_, 16, SEL_, 168, false, 1>(p, layouts[which], n_items, stream);       \
        }                                                                                                             \
        return launch_strip_variant<NB_, RG_, 8, SEL_, 168, false, 1>(p, layouts[which], n_items, stream);            \
    } while (0)
    if (nb == 4) DMX_STRIP(4, 4, true);
    DMX_STRIP(8, 1, false);
#undef DMX_STRIP_V
#undef DMX_STRIP
}

}  // namespace dmx
