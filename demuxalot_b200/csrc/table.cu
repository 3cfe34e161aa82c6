// Regularised betas (demux.py:367-390) and the per-SNP normalised probability table (demux.py:267-274).
//
// Both are small streaming kernels over the [V, G] float32 betas: HBM-bound, 12 bytes per element for the
// table (read betas, read addition, write P).  All SNP sums are float64 and are taken over the SNP's
// variants in ascending variant id, which is the order np.bincount adds them in, so results are bit-exact.
#include "common.cuh"

#include <stdlib.h>

namespace dmx {

// numpy's float32 pairwise summation (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum) as used by
// `betas.sum(axis=1)` at demux.py:383: 8 interleaved accumulators for n <= 128, recursive halving above.
__device__ float np_pairwise_sum_f32(const float* __restrict__ a, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], a[i + k]);
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum_f32(a, n2), np_pairwise_sum_f32(a + n2, n - n2));
}

__global__ void rowsum_kernel(const float* __restrict__ betas, int64_t ld, int64_t n_variants, int n_genotypes,
                              float* __restrict__ rowsum) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_variants;
         v += (int64_t)gridDim.x * blockDim.x) {
        rowsum[v] = __fadd_rn(0.f, np_pairwise_sum_f32(betas + v * ld, n_genotypes));
    }
}

// one thread per SNP: float64 sums over its variants, then the per-variant prior addition (float32)
__global__ void prior_addition_kernel(const float* rowsum, const int32_t* __restrict__ snp_offsets,
                                      const int32_t* __restrict__ snp_variants, int64_t n_snps,
                                      const int64_t* __restrict__ n_mol, double default_prior,
                                      float* addition /* may alias rowsum: a thread writes only its own SNP's
                                                         variants, each after its last read */) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n_snps; s += (int64_t)gridDim.x * blockDim.x) {
        const int lo = snp_offsets[s], hi = snp_offsets[s + 1];
        double sum_betas = 0.0, sum_mol = 0.0;
        for (int k = lo; k < hi; ++k) {
            const int v = snp_variants[k];
            sum_betas += (double)rowsum[v];
            if (n_mol) sum_mol += (double)n_mol[v];
        }
        for (int k = lo; k < hi; ++k) {
            const int v = snp_variants[k];
            double prior = 1.0;
            if (n_mol) prior = prior + (double)n_mol[v] / (sum_mol + 100.0);
            prior = prior + (double)rowsum[v] / (sum_betas + 100.0);
            addition[v] = (float)(prior * default_prior);
        }
    }
}

__global__ void add_prior_kernel(const float* __restrict__ raw, int64_t ld_raw, const float* __restrict__ addition,
                                 int64_t n_variants, int n_genotypes, float* __restrict__ out, int64_t ld_out) {
    const int64_t total = n_variants * n_genotypes;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = k / n_genotypes;
        const int g = (int)(k - v * n_genotypes);
        out[v * ld_out + g] = __fadd_rn(raw[v * ld_raw + g], addition[v]);
    }
}

// Probability table.  Thread (s, g): g runs fastest so a warp reads/writes contiguous pieces of table rows.
// Work item = (SNP, column) over the padded width ld_table; padded columns are filled with 1.
__global__ void probs_table_kernel(const float* __restrict__ betas, int64_t ld_betas,
                                   const float* __restrict__ addition, int64_t ld_add, int n_genotypes,
                                   const int32_t* __restrict__ snp_offsets, const int32_t* __restrict__ snp_variants,
                                   int64_t n_snps, float clip_lo, float clip_hi, float* __restrict__ table,
                                   int64_t ld_table) {
    const int64_t total = n_snps * ld_table;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = k / ld_table;
        const int g = (int)(k - s * ld_table);
        const int lo = snp_offsets[s], hi = snp_offsets[s + 1];
        if (g >= n_genotypes) {
            for (int q = lo; q < hi; ++q) table[(int64_t)snp_variants[q] * ld_table + g] = 1.0f;
            continue;
        }
        double den = 0.0;
        for (int q = lo; q < hi; ++q) {
            const int64_t v = snp_variants[q];
            float b = betas[v * ld_betas + g];
            if (addition) b = __fadd_rn(b, addition[v * ld_add + g]);
            den += (double)b;
        }
        den = fmax(den, 1e-7);
        for (int q = lo; q < hi; ++q) {
            const int64_t v = snp_variants[q];
            float b = betas[v * ld_betas + g];
            if (addition) b = __fadd_rn(b, addition[v * ld_add + g]);
            float p = (float)((double)b / den);
            p = fminf(fmaxf(p, clip_lo), clip_hi);
            table[v * ld_table + g] = p;
        }
    }
}

// Vector flavour for G % 4 == 0 (no padding columns): a thread owns 4 consecutive columns of one SNP and keeps the
// rows of up to TABLE_MAXV variants in registers, so each element is read once with 128-bit loads and several times the
// bytes are in flight per thread (the scalar kernel reached only 29 % of the HBM peak: one 4-byte load per thread at
// the end of a dependent offsets -> variant -> betas chain).  Same arithmetic, same summation order: bit-exact.
// TABLE_MAXV / MINB (resident CTAs asked of the compiler) are template knobs so that DMX_TABLE_* can sweep them.
template <int TABLE_MAXV, int MINB>
__global__ void __launch_bounds__(256, MINB) probs_table_vec4_kernel(
    const float* __restrict__ betas, int64_t ld_betas, const float* __restrict__ addition, int64_t ld_add, int quads,
    const int32_t* __restrict__ snp_offsets, const int32_t* __restrict__ snp_variants, int64_t n_snps, float clip_lo,
    float clip_hi, float* __restrict__ table, int64_t ld_table) {
    const int64_t total = n_snps * quads;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = k / quads;
        const int g = 4 * (int)(k - s * quads);
        const int lo = __ldg(snp_offsets + s), hi = __ldg(snp_offsets + s + 1);
        const int n = hi - lo;
        auto load_row = [&](int64_t v) {
            float4 b = ldg_stream_f4(reinterpret_cast<const float4*>(betas + v * ld_betas + g));
            if (addition) {
                const float4 a = ldg_stream_f4(reinterpret_cast<const float4*>(addition + v * ld_add + g));
                b.x = __fadd_rn(b.x, a.x); b.y = __fadd_rn(b.y, a.y); b.z = __fadd_rn(b.z, a.z); b.w = __fadd_rn(b.w, a.w);
            }
            return b;
        };
        auto finish = [&](float b, double den) {
            const float p = (float)((double)b / den);
            return fminf(fmaxf(p, clip_lo), clip_hi);
        };
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
        if (n <= TABLE_MAXV) {
            int64_t v[TABLE_MAXV];
            float4 b[TABLE_MAXV];
#pragma unroll
            for (int q = 0; q < TABLE_MAXV; ++q) v[q] = q < n ? (int64_t)__ldg(snp_variants + lo + q) : -1;
#pragma unroll
            for (int q = 0; q < TABLE_MAXV; ++q) b[q] = q < n ? load_row(v[q]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < TABLE_MAXV; ++q)
                if (q < n) { d0 += (double)b[q].x; d1 += (double)b[q].y; d2 += (double)b[q].z; d3 += (double)b[q].w; }
            d0 = fmax(d0, 1e-7); d1 = fmax(d1, 1e-7); d2 = fmax(d2, 1e-7); d3 = fmax(d3, 1e-7);
#pragma unroll
            for (int q = 0; q < TABLE_MAXV; ++q)
                if (q < n) {
                    const float4 p = make_float4(finish(b[q].x, d0), finish(b[q].y, d1), finish(b[q].z, d2), finish(b[q].w, d3));
                    *reinterpret_cast<float4*>(table + v[q] * ld_table + g) = p;
                }
        } else {  // SNPs with many alleles: two passes over the rows
            for (int q = lo; q < hi; ++q) {
                const float4 b = load_row(snp_variants[q]);
                d0 += (double)b.x; d1 += (double)b.y; d2 += (double)b.z; d3 += (double)b.w;
            }
            d0 = fmax(d0, 1e-7); d1 = fmax(d1, 1e-7); d2 = fmax(d2, 1e-7); d3 = fmax(d3, 1e-7);
            for (int q = lo; q < hi; ++q) {
                const int64_t v = snp_variants[q];
                const float4 b = load_row(v);
                *reinterpret_cast<float4*>(table + v * ld_table + g) =
                    make_float4(finish(b.x, d0), finish(b.y, d1), finish(b.z, d2), finish(b.w, d3));
            }
        }
    }
}

static inline int grid_1d(int64_t n, int threads, int ctas_per_sm = 32) {
    int64_t blocks = ceil_div(n > 0 ? n : 1, threads);
    const int64_t cap = ctas_per_sm > 0 ? (int64_t)sm_count() * ctas_per_sm : (int64_t)0x7fffffff;
    return (int)(blocks < cap ? blocks : cap);
}

static inline int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : fallback;
}

}  // namespace dmx

extern "C" {

int dmx_prior_betas(const float* raw_betas, int64_t ld_raw, int64_t n_variants, int32_t n_genotypes,
                    const int32_t* snp_offsets, const int32_t* snp_variants, int64_t n_snps,
                    const int64_t* n_mol_per_variant, double default_prior, float* scratch_rowsum,
                    float* out_betas, int64_t ld_out, void* stream_) {
    using namespace dmx;
    if (n_variants <= 0 || n_genotypes <= 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int threads = 256;
    rowsum_kernel<<<grid_1d(n_variants, threads), threads, 0, stream>>>(raw_betas, ld_raw, n_variants, n_genotypes,
                                                                       scratch_rowsum);
    DMX_LAUNCH_CHECK();
    prior_addition_kernel<<<grid_1d(n_snps, threads), threads, 0, stream>>>(
        scratch_rowsum, snp_offsets, snp_variants, n_snps, n_mol_per_variant, default_prior, scratch_rowsum);
    DMX_LAUNCH_CHECK();
    add_prior_kernel<<<grid_1d(n_variants * n_genotypes, threads), threads, 0, stream>>>(
        raw_betas, ld_raw, scratch_rowsum, n_variants, n_genotypes, out_betas, ld_out);
    DMX_LAUNCH_CHECK();
    return 0;
}

int dmx_probs_from_betas(const float* betas, int64_t ld_betas, const float* addition, int64_t ld_addition,
                         int64_t n_variants, int32_t n_genotypes, const int32_t* snp_offsets,
                         const int32_t* snp_variants, int64_t n_snps, float clip_lo, float clip_hi, float* table,
                         int64_t ld_table, void* stream_) {
    using namespace dmx;
    if (n_variants <= 0 || n_genotypes <= 0) return 0;
    DMX_REQUIRE(ld_table >= n_genotypes, "ld_table %lld < n_genotypes %d", (long long)ld_table, n_genotypes);
    const int threads = 256;
    const bool vec4 = n_genotypes % 4 == 0 && ld_table % 4 == 0 && ld_betas % 4 == 0 && ((uintptr_t)betas & 15) == 0 &&
                      ((uintptr_t)table & 15) == 0 && (!addition || (ld_addition % 4 == 0 && ((uintptr_t)addition & 15) == 0));
    if (vec4) {
        const int quads = n_genotypes / 4;
        // Defaults from scripts/sweep_table.py (profiles/r01_sweep_table.log): 2 rows in registers (bi-allelic SNPs,
        // anything wider takes the two-pass loop) and 5 resident CTAs (48 registers, 32 bytes of spills) run 25-28 % faster
        // than 4 rows at 68 registers; 6 CTAs spill 156 bytes and lose.  Knobs: rows in registers, resident CTAs, grid cap in CTAs per SM.
        const int maxv = env_int("DMX_TABLE_MAXV", 2), minb = env_int("DMX_TABLE_MINB", 5);
        const int grid = grid_1d(n_snps * quads, threads, env_int("DMX_TABLE_CAP", 32));
#define DMX_TABLE_LAUNCH(MAXV, MINB)                                                                                 \
    probs_table_vec4_kernel<MAXV, MINB><<<grid, threads, 0, (cudaStream_t)stream_>>>(                                \
        betas, ld_betas, addition, ld_addition, quads, snp_offsets, snp_variants, n_snps, clip_lo, clip_hi, table,   \
        ld_table)
        if (maxv == 2 && minb >= 5) DMX_TABLE_LAUNCH(2, 5);
        else if (maxv == 2) DMX_TABLE_LAUNCH(2, 4);
        else if (minb >= 4) DMX_TABLE_LAUNCH(4, 4);
        else DMX_TABLE_LAUNCH(4, 1);
#undef DMX_TABLE_LAUNCH
        DMX_LAUNCH_CHECK();
        return 0;
    }
    probs_table_kernel<<<grid_1d(n_snps * ld_table, threads), threads, 0, (cudaStream_t)stream_>>>(
        betas, ld_betas, addition, ld_addition, n_genotypes, snp_offsets, snp_variants, n_snps, clip_lo, clip_hi,
        table, ld_table);
    DMX_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
