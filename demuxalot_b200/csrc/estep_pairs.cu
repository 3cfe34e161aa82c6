// E-step with doublet columns (doublet_prior != 0): a barcode-batched "log-semiring SYRK".
//
//   S_b[i, j] = sum_{rows r of barcode b} log( 0.5 (P[v_r, i] + P[v_r, j]) (1 - e_r) + max(e_r, 1e-4) ),  i <= j
//
// (diagonal = singlet columns, demux.py:179-191, 261).  There is no tensor-core formulation (the log of a sum
// of two gathered values is not a contraction); the kernel is bound by FP32 issue and, for the EXACT flavour,
// by the logf expansion -- HBM traffic is only 8 + 4G bytes per row against G(G+1)/2 updates.
//
// Decomposition
//   * the upper triangle of the G x G pair matrix is cut into 4 x 4 register tiles, one per thread
//     (16 float64 accumulators + 8 packed products: ~90 registers, two 256-thread CTAs per SM);
//   * a CTA owns one barcode -- or a slice of its tiles when there are many -- and `row_groups` copies of the
//     tile set, each walking a different subset of the barcode's rows; the copies are reduced through shared
//     memory in a fixed order at the end (no atomics -> deterministic);
//   * table rows P[v_r, :] are gathered with cp.async into a double-buffered shared-memory stage; row records
//     (variant, p_base_wrong) are prefetched into registers one chunk ahead; each thread finishes the pieces it
//     copied itself (a = fma(P, 1-e, e') for FAST), so one __syncthreads per stage is enough;
//   * FAST inner loop: packed f32x2 adds/multiplies (FADD2 / FMUL2 on sm_100, the scalar operand is broadcast
//     by the instruction) keep a running float32 product per pair; every FLUSH_ROWS rows the binary exponent
//     of each product is moved into an integer sum and the mantissa reset to [1, 2) (exact; 3 ALU instructions),
//     so the loop contains no MUFU and no FP64 at all and one lg2 per pair is taken at the very end.  Factors
//     are 2x the reference argument (a_i + a_j); the 1/2 per row is removed exactly in the epilogue.  Padding
//     rows are staged as a = 1 (factor 2, log2 = 1) and cancel there too.
#include "common.cuh"

namespace dmx {

constexpr int TILE = 4;        // thread tile: 4 (i) x TJ (j) pairs, TJ = 4 or 8 (template parameter)
constexpr int MAX_PASSES = 2;  // staging slots per thread and chunk
constexpr int QPT = 8;         // 16-byte quads per staging slot
constexpr int MAX_THREADS = 256;
constexpr float ERROR_FLOOR = 1e-4f;

struct PairsParams {
    const int64_t* offsets;
    const int32_t* order;  // launch schedule (barcodes by descending row count) or nullptr
    const int32_t* variant;
    const float* e;
    const float* table;
    int64_t ld_table;
    int n_genotypes;
    int gp;             // genotypes rounded up to a multiple of 4
    int n_tiles;        // 4x4 tiles covering the upper triangle (diagonal included)
    int tiles_per_cta;  // tiles handled by one CTA
    int ctas_per_barcode;
    int row_groups;     // row groups inside a CTA
    int flushes;        // products per row group and staged chunk
    int ld_smem;        // floats per staged row (gp + 4)
    float doublet_bonus;
    const double* prior;  // float64, as numpy adds it: float32(float64(logit) + prior), demux.py:99
    int64_t ld_prior;
    float* logits;
    int64_t ld_logits;
};

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// products are always normal numbers here, so the flush-to-zero form (a bare MUFU.LG2) loses nothing
__device__ __forceinline__ float lg2_raw(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Variants measured and dropped (scripts/sweep_estep.py, profiles/): 8x4 tiles (168 registers, same speed),
// __launch_bounds__(256, 3) (80 registers: no faster even without spills -- the kernel is not occupancy bound),
// one lg2 + float64 add per 16-row product instead of the integer exponent bookkeeping (3-6 % slower), scalar
// FADD/FMUL instead of the packed forms (8-10 % slower).
template <int FLAVOUR, int FLUSH_ROWS, int TJ>
__global__ void __launch_bounds__(MAX_THREADS, 2) estep_pairs_kernel(const PairsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int n_threads = blockDim.x;
    const int64_t slot_in_grid = blockIdx.x / p.ctas_per_barcode;
    const int cta_in_barcode = (int)(blockIdx.x - slot_in_grid * p.ctas_per_barcode);
    const int64_t barcode = p.order ? (int64_t)p.order[slot_in_grid] : slot_in_grid;

    const int tile_local = tid % p.tiles_per_cta;
    const int rg = tid / p.tiles_per_cta;
    const int tile = cta_in_barcode * p.tiles_per_cta + tile_local;
    const bool has_tile = tile < p.n_tiles && rg < p.row_groups;  // CTAs may carry staging-only threads

    // tile -> (i0, j0): tiles are enumerated i-block major; i-block pi (4 wide) owns the j-blocks (TJ wide) that
    // reach the diagonal or lie above it: 4 pi / TJ .. gp / TJ - 1
    int i0 = 0, j0 = 0;
    {
        const int q_total = p.gp / TJ;
        int t = has_tile ? tile : 0, pi = 0;
        while (t >= q_total - (pi * TILE) / TJ) { t -= q_total - (pi * TILE) / TJ; ++pi; }
        i0 = pi * TILE;
        j0 = ((pi * TILE) / TJ + t) * TJ;
    }

    const int chunk_rows = p.row_groups * p.flushes * FLUSH_ROWS;
    const int ld = p.ld_smem;
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    float* stage1 = stage0 + chunk_rows * ld;
    double* reduce_buf = reinterpret_cast<double*>(stage1 + chunk_rows * ld);

    const int64_t row_lo = p.offsets[barcode];
    const int64_t row_hi = p.offsets[barcode + 1];
    const int n_chunks = (int)((row_hi - row_lo + chunk_rows - 1) / chunk_rows);

    double acc[TILE][TJ];          // EXACT: running float64 sums; FAST: filled once after the row loop
    uint64_t prod[TILE / 2][TJ];   // FAST: running products (packed float32 pairs), mantissas kept in [1, 2)
    int esum[TILE][TJ];            // FAST: biased binary exponents moved out of the products
#pragma unroll
    for (int a = 0; a < TILE; ++a)
#pragma unroll
        for (int b = 0; b < TJ; ++b) { acc[a][b] = 0.0; esum[a][b] = 0; }
#pragma unroll
    for (int a = 0; a < TILE / 2; ++a)
#pragma unroll
        for (int b = 0; b < TJ; ++b) prod[a][b] = pack2(1.f, 1.f);

    // ---- staging -----------------------------------------------------------------------------------------------
    // A slot = up to QPT consecutive 16-byte quads of one staged row (a whole row for G <= 32); thread t owns the
    // slots t, t + n_threads, ...  Row records are prefetched into registers one chunk ahead, so the cp.async
    // addresses never wait on a global load inside the steady state.
    const int quads = p.gp / 4;
    const int pieces_per_row = (quads + QPT - 1) / QPT;
    const int n_slots = chunk_rows * pieces_per_row;
    int slot_row[MAX_PASSES], slot_q0[MAX_PASSES];
#pragma unroll
    for (int s = 0; s < MAX_PASSES; ++s) {
        const int slot = tid + s * n_threads;
        slot_row[s] = -1;
        slot_q0[s] = 0;
        if (slot < n_slots) {
            slot_row[s] = slot / pieces_per_row;
            slot_q0[s] = (slot - slot_row[s] * pieces_per_row) * QPT;
        }
    }
    int v_pre[MAX_PASSES];
    float e_pre[MAX_PASSES], e_cur[MAX_PASSES];
    unsigned live = 0;  // bit s: slot s holds real table data that land() must finish

    auto prefetch = [&](int chunk) {
        const int64_t base = row_lo + (int64_t)chunk * chunk_rows;
#pragma unroll
        for (int s = 0; s < MAX_PASSES; ++s) {
            v_pre[s] = -1;  // padding row
            e_pre[s] = 0.f;
            if (slot_row[s] >= 0) {
                const int64_t row = base + slot_row[s];
                if (row < row_hi) {
                    v_pre[s] = __ldg(p.variant + row);
                    e_pre[s] = __ldg(p.e + row);
                }
            }
        }
    };
    auto issue = [&](float* buf) {  // consumes the registers filled by prefetch()
        live = 0;
#pragma unroll
        for (int s = 0; s < MAX_PASSES; ++s) {
            if (slot_row[s] >= 0) {
                float* dst = buf + slot_row[s] * ld + 4 * slot_q0[s];
                e_cur[s] = e_pre[s];
                const bool whole = slot_q0[s] + QPT <= quads && 4 * (slot_q0[s] + QPT) <= p.ld_table;
                if (v_pre[s] >= 0) {
                    const float* src = p.table + (int64_t)v_pre[s] * p.ld_table + 4 * slot_q0[s];
                    live |= 1u << s;
                    if (whole) {
#pragma unroll
                        for (int u = 0; u < QPT; ++u) cp_async_16(dst + 4 * u, src + 4 * u);
                    } else {
#pragma unroll
                        for (int u = 0; u < QPT; ++u) {
                            const int q = slot_q0[s] + u;
                            if (q < quads) {
                                if (4 * q < p.ld_table) cp_async_16(dst + 4 * u, src + 4 * u);
                                else *reinterpret_cast<float4*>(dst + 4 * u) = make_float4(1.f, 1.f, 1.f, 1.f);
                            }
                        }
                    }
                } else {  // padding row: neutral element (factor 2 -> log2 = 1, removed in the epilogue)
#pragma unroll
                    for (int u = 0; u < QPT; ++u)
                        if (slot_q0[s] + u < quads)
                            *reinterpret_cast<float4*>(dst + 4 * u) = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (FLAVOUR == DMX_ESTEP_EXACT && slot_q0[s] == 0) {
                        float* row_consts = buf + slot_row[s] * ld + p.gp;
                        row_consts[0] = 0.5f;  // (1 + 1) * 0.5 + 0 = 1 -> log = 0
                        row_consts[1] = 0.f;
                    }
                }
            }
        }
        cp_async_commit();
    };
    auto land = [&](float* buf) {  // each thread finishes the pieces it copied itself: no extra barrier needed
        cp_async_wait<0>();
#pragma unroll
        for (int s = 0; s < MAX_PASSES; ++s) {
            if (live & (1u << s)) {
                float* dst = buf + slot_row[s] * ld + 4 * slot_q0[s];
                const float e = e_cur[s];
                const float w = __fsub_rn(1.f, e);
                const float ef = fmaxf(e, ERROR_FLOOR);
                if (FLAVOUR == DMX_ESTEP_FAST) {
#pragma unroll
                    for (int u = 0; u < QPT; ++u) {
                        const int q = slot_q0[s] + u;
                        if (q < quads && 4 * q < p.ld_table) {
                            float4 x = *reinterpret_cast<float4*>(dst + 4 * u);
                            x.x = fmaf(x.x, w, ef);
                            x.y = fmaf(x.y, w, ef);
                            x.z = fmaf(x.z, w, ef);
                            x.w = fmaf(x.w, w, ef);
                            *reinterpret_cast<float4*>(dst + 4 * u) = x;
                        }
                    }
                } else if (slot_q0[s] == 0) {
                    float* row_consts = buf + slot_row[s] * ld + p.gp;
                    row_consts[0] = 0.5f * w;  // exact scaling
                    row_consts[1] = ef;
                }
            }
        }
    };

    if (n_chunks > 0) {
        prefetch(0);
        issue(stage0);
        if (n_chunks > 1) prefetch(1);
        land(stage0);
        __syncthreads();
    }

    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        float* cur = (chunk & 1) ? stage1 : stage0;
        float* nxt = (chunk & 1) ? stage0 : stage1;
        const bool more = chunk + 1 < n_chunks;
        if (more) {
            issue(nxt);
            if (chunk + 2 < n_chunks) prefetch(chunk + 2);
        }

        if (has_tile) {
            for (int f = 0; f < p.flushes; ++f) {
                const float* rows = cur + (f * FLUSH_ROWS * p.row_groups + rg) * ld;
                const int row_stride = p.row_groups * ld;
                if (FLAVOUR == DMX_ESTEP_FAST) {
#pragma unroll
                    for (int k = 0; k < FLUSH_ROWS; ++k) {
                        const float* s = rows + k * row_stride;
                        const float4 ai = *reinterpret_cast<const float4*>(s + i0);
                        const uint64_t a2[TILE / 2] = {pack2(ai.x, ai.y), pack2(ai.z, ai.w)};
                        float aj[TJ];
#pragma unroll
                        for (int q = 0; q < TJ / 4; ++q) {
                            const float4 bj = *reinterpret_cast<const float4*>(s + j0 + 4 * q);
                            aj[4 * q] = bj.x; aj[4 * q + 1] = bj.y; aj[4 * q + 2] = bj.z; aj[4 * q + 3] = bj.w;
                        }
#pragma unroll
                        for (int a = 0; a < TILE / 2; ++a)
#pragma unroll
                            for (int b = 0; b < TJ; ++b)
                                prod[a][b] = mul2(prod[a][b], add2(a2[a], pack2(aj[b], aj[b])));
                    }
                    // renormalise: move the binary exponent of every running product into an integer sum and keep
                    // the mantissa in [1, 2) -- exact, three ALU instructions per product, no MUFU / FP64 in the loop
#pragma unroll
                    for (int a = 0; a < TILE / 2; ++a)
#pragma unroll
                        for (int b = 0; b < TJ; ++b) {
                            float lo, hi;
                            unpack2(prod[a][b], lo, hi);
                            const unsigned blo = __float_as_uint(lo), bhi = __float_as_uint(hi);
                            esum[2 * a][b] += (int)(blo >> 23);
                            esum[2 * a + 1][b] += (int)(bhi >> 23);
                            prod[a][b] = pack2(__uint_as_float((blo & 0x007fffffu) | 0x3f800000u),
                                               __uint_as_float((bhi & 0x007fffffu) | 0x3f800000u));
                        }
                } else {
#pragma unroll 4
                    for (int k = 0; k < FLUSH_ROWS; ++k) {
                        const float* s = rows + k * row_stride;
                        const float4 ai = *reinterpret_cast<const float4*>(s + i0);
                        const float hw = s[p.gp];
                        const float ef = s[p.gp + 1];
                        const float pi[TILE] = {ai.x, ai.y, ai.z, ai.w};
                        float pj[TJ];
#pragma unroll
                        for (int q = 0; q < TJ / 4; ++q) {
                            const float4 bj = *reinterpret_cast<const float4*>(s + j0 + 4 * q);
                            pj[4 * q] = bj.x; pj[4 * q + 1] = bj.y; pj[4 * q + 2] = bj.z; pj[4 * q + 3] = bj.w;
                        }
#pragma unroll
                        for (int a = 0; a < TILE; ++a)
#pragma unroll
                            for (int b = 0; b < TJ; ++b) {
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

    if (FLAVOUR == DMX_ESTEP_FAST) {
        // log2 of the row group's product = (sum of unbiased exponents) + log2(mantissa in [1, 2))
        const int bias = 127 * n_chunks * p.flushes;
#pragma unroll
        for (int a = 0; a < TILE / 2; ++a)
#pragma unroll
            for (int b = 0; b < TJ; ++b) {
                float lo, hi;
                unpack2(prod[a][b], lo, hi);
                acc[2 * a][b] = (double)(esum[2 * a][b] - bias) + (double)lg2_raw(lo);
                acc[2 * a + 1][b] = (double)(esum[2 * a + 1][b] - bias) + (double)lg2_raw(hi);
            }
    }

    // ---- fixed-order reduction over the row groups (deterministic) ---------------------------------------------
    if (p.row_groups > 1) {
        for (int g = 0; g < p.row_groups; ++g) {
            if (rg == g && has_tile) {
#pragma unroll
                for (int a = 0; a < TILE; ++a)
#pragma unroll
                    for (int b = 0; b < TJ; ++b) {
                        double* slot = reduce_buf + (a * TJ + b) * p.tiles_per_cta + tile_local;
                        if (g == 0) *slot = acc[a][b]; else *slot += acc[a][b];
                    }
            }
            __syncthreads();
        }
        if (rg == 0 && has_tile) {
#pragma unroll
            for (int a = 0; a < TILE; ++a)
#pragma unroll
                for (int b = 0; b < TJ; ++b) acc[a][b] = reduce_buf[(a * TJ + b) * p.tiles_per_cta + tile_local];
        }
    }

    // ---- epilogue: penalties, prior logits, one rounding to float32 ---------------------------------------------
    if (rg == 0 && has_tile) {
        const int G = p.n_genotypes;
        const double padded_rows = (double)n_chunks * (double)chunk_rows;
#pragma unroll
        for (int a = 0; a < TILE; ++a) {
            const int i = i0 + a;
#pragma unroll
            for (int b = 0; b < TJ; ++b) {
                const int j = j0 + b;
                if (i < G && j < G && j >= i) {
                    const int64_t col = (i == j) ? i : (int64_t)G + (int64_t)i * G - (int64_t)i * (i + 1) / 2 + (j - i - 1);
                    double sum = acc[a][b];
                    if (FLAVOUR == DMX_ESTEP_FAST) sum = (sum - padded_rows) * 0.693147180559945309417232;
                    const float pen = (i == j) ? 0.f : p.doublet_bonus;
                    float logit = (float)((double)pen + sum);
                    if (p.prior) logit = (float)((double)logit + p.prior[barcode * p.ld_prior + col]);
                    p.logits[barcode * p.ld_logits + col] = logit;
                }
            }
        }
    }
}

static float doublet_bonus(int n_genotypes, double dp) {
    // demux.py:168-172 (float64, rounded to float32 on assignment)
    const double g = (double)n_genotypes;
    double bonus = log(g * dp);
    bonus -= log(g * (double)(n_genotypes - 1 > 1 ? n_genotypes - 1 : 1) / 2 * (1 - dp));
    return (float)bonus;
}

static int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}

template <int FLAVOUR, int FLUSH_ROWS, int TJ>
static int launch_variant(const PairsParams& p, unsigned grid, int threads, size_t smem, cudaStream_t stream) {
    auto kernel = estep_pairs_kernel<FLAVOUR, FLUSH_ROWS, TJ>;
    DMX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DMX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
    kernel<<<grid, threads, smem, stream>>>(p);
    DMX_LAUNCH_CHECK();
    return 0;
}

// table_floor: lower bound of the table entries (the clip of demux.py:274), 0 if unknown.  It decides how many
// row factors (>= 2 (table_floor + 1e-4) each) can be multiplied in float32 without leaving the normal range.
// Tuning overrides (experiments only): DMX_RG, DMX_FLUSHES, DMX_FLUSH_ROWS, DMX_MAX_THREADS, DMX_VERBOSE.
int launch_estep_pairs(const int64_t* barcode_offsets, const int32_t* barcode_order, const int32_t* csr_variant,
                       const float* csr_e, int64_t n_barcodes, const float* table, int64_t ld_table, int G,
                       double doublet_prior,
                       float table_floor, const double* prior_logits, int64_t ld_prior, float* logits,
                       int64_t ld_logits, int flavour, cudaStream_t stream) {
    PairsParams p;
    p.offsets = barcode_offsets;
    p.order = barcode_order;
    p.variant = csr_variant;
    p.e = csr_e;
    p.table = table;
    p.ld_table = ld_table;
    p.n_genotypes = G;
    // thread tile 4 x TJ: the wider tile reads 3 instead of 4 shared-memory vectors per 32 updates but wastes more
    // lanes on the diagonal, so it pays off for many genotypes (measured: G = 200 yes, G = 32 no)
    int tj = env_int("DMX_TJ", G >= 48 ? 8 : 4);
    if (flavour != DMX_ESTEP_FAST || (tj != 4 && tj != 8)) tj = 4;
    p.gp = (int)round_up(G, tj);
    int n_tiles = 0;
    for (int pi = 0; pi < p.gp / TILE; ++pi) n_tiles += p.gp / tj - (pi * TILE) / tj;
    p.n_tiles = n_tiles;
    const int quads = p.gp / 4;
    const int pieces_per_row = (quads + QPT - 1) / QPT;
    p.ld_smem = p.gp + 4;

    // 16 factors per product need 16 * -log2(2 (floor + 1e-4)) <= 120 binades
    const bool long_products_safe = flavour == DMX_ESTEP_FAST && table_floor >= 0.0027f;
    int flush_rows = (long_products_safe && env_int("DMX_FLUSH_ROWS", 16) == 16) ? 16 : 8;

    // Shape of a CTA: `ctas_per_barcode` CTAs split the tiles of a barcode, each with `row_groups` copies of its
    // tile slice, at most 256 threads (two CTAs resident per SM).
    // measured on B200 (scripts/sweep_estep.py): 128-thread CTAs win while one CTA covers a barcode's tiles
    // (G = 32: 1.58 vs 1.71 ms), 256-thread CTAs win once the tiles are split over CTAs (G = 200: 7.9 vs 11.2 ms)
    int max_threads = env_int("DMX_MAX_THREADS", n_tiles <= 64 ? 128 : MAX_THREADS);
    if (max_threads > MAX_THREADS || max_threads < 32) max_threads = MAX_THREADS;
    {
        // as few CTAs per barcode as the thread budget allows (every CTA stages whole table rows, so splitting
        // a barcode's tiles repeats the staging), then as many row groups as fit
        p.ctas_per_barcode = (int)ceil_div(n_tiles, max_threads);
        p.tiles_per_cta = (int)ceil_div(n_tiles, p.ctas_per_barcode);
        int rg = max_threads / p.tiles_per_cta;
        const int forced_rg = env_int("DMX_RG", 0);
        if (forced_rg > 0 && forced_rg * p.tiles_per_cta <= MAX_THREADS) rg = forced_rg;
        p.row_groups = rg < 1 ? 1 : (rg > 16 ? 16 : rg);
    }
    const int compute_threads = p.tiles_per_cta * p.row_groups;
    DMX_REQUIRE(compute_threads <= MAX_THREADS, "internal: CTA too large");
    // Rows per staged chunk.  Every thread has MAX_PASSES staging slots; when a barcode has only a few tiles the
    // CTA is padded with threads that only stage (has_tile == false).  Stages are capped at 96 KB.
    const int max_chunk_rows = MAX_PASSES * MAX_THREADS / pieces_per_row;
    while (p.row_groups > 1 && p.row_groups * 8 > max_chunk_rows) --p.row_groups;
    if (p.row_groups * flush_rows > max_chunk_rows) flush_rows = 8;
    DMX_REQUIRE(p.row_groups * flush_rows <= max_chunk_rows, "n_genotypes %d too large for the pair kernel's staging", G);
    p.flushes = 1;
    const int want_flushes = env_int("DMX_FLUSHES", 2);
    while (p.flushes < want_flushes && p.row_groups * flush_rows * (p.flushes + 1) <= max_chunk_rows &&
           2 * (size_t)p.row_groups * flush_rows * (p.flushes + 1) * p.ld_smem * sizeof(float) <= 96 * 1024)
        ++p.flushes;
    const int64_t chunk_rows = (int64_t)p.row_groups * p.flushes * flush_rows;
    const int stagers = (int)ceil_div(chunk_rows * pieces_per_row, MAX_PASSES);
    const int threads = (int)round_up(compute_threads > stagers ? compute_threads : stagers, 32);
    DMX_REQUIRE(threads <= MAX_THREADS, "internal: CTA too large after adding staging threads");
    size_t smem = 2 * (size_t)chunk_rows * p.ld_smem * sizeof(float);
    if (p.row_groups > 1) smem += (size_t)p.tiles_per_cta * TILE * tj * sizeof(double);
    DMX_REQUIRE(smem <= 200 * 1024, "shared memory request too large");
    p.doublet_bonus = doublet_bonus(G, doublet_prior);
    p.prior = prior_logits;
    p.ld_prior = ld_prior;
    p.logits = logits;
    p.ld_logits = ld_logits;
    const int64_t grid = n_barcodes * p.ctas_per_barcode;
    DMX_REQUIRE(grid < (1ll << 31), "grid too large");
    if (env_int("DMX_VERBOSE", 0))
        fprintf(stderr,
                "[dmx] pair E-step: G=%d tiles=%d ctas/barcode=%d tiles/cta=%d row_groups=%d threads=%d flush_rows=%d "
                "flushes=%d chunk_rows=%lld smem=%zu flavour=%d tj=%d\n",
                G, n_tiles, p.ctas_per_barcode, p.tiles_per_cta, p.row_groups, threads, flush_rows, p.flushes,
                (long long)chunk_rows, smem, flavour, tj);

    if (flavour == DMX_ESTEP_EXACT) return launch_variant<DMX_ESTEP_EXACT, 8, 4>(p, (unsigned)grid, threads, smem, stream);
    if (tj == 8) {
        if (flush_rows == 16) return launch_variant<DMX_ESTEP_FAST, 16, 8>(p, (unsigned)grid, threads, smem, stream);
        return launch_variant<DMX_ESTEP_FAST, 8, 8>(p, (unsigned)grid, threads, smem, stream);
    }
    if (flush_rows == 16) return launch_variant<DMX_ESTEP_FAST, 16, 4>(p, (unsigned)grid, threads, smem, stream);
    return launch_variant<DMX_ESTEP_FAST, 8, 4>(p, (unsigned)grid, threads, smem, stream);
}

}  // namespace dmx
