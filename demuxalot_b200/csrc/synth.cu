// Bench / test tooling, NOT part of the product ABI: device generator of synthetic count_snps output for the
// biobank-scale workload (BASELINE.json configs[3]: 200 donors, 100 k barcodes, 5 M variants, 500 M rows), which is
// too large to draw with numpy on the host ("rows generated on device per barcode shard with
// a counter-based RNG keyed by (seed, barcode, k) so any GPU count yields identical data").
//
// Every random decision is a pure integer function of (seed, barcode, group, molecule) -- splitmix64 mixing, 16/20-bit
// threshold comparisons, table look-ups for the float32 error values -- so `demuxalot_b200/synthetic_device.py`
// reproduces any barcode subset bit for bit with numpy uint64 arithmetic (the slice handed to the CPU oracle).
// Built into libdemux_synth.so (separate from libdemux_b200.so).
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t x) {
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__host__ __device__ inline uint64_t draw(uint64_t seed, uint64_t barcode, uint64_t group, uint64_t what) {
    return mix64(mix64(seed + barcode * 0x9e3779b97f4a7c15ull) + group * 0xd1b54a32d192ed03ull + what);
}

struct SynthParams {
    uint64_t seed;
    int64_t n_snps;
    int32_t n_genotypes;
    const int32_t* barcode_ids;     // [n_local] global barcode id of the k-th local barcode
    const int64_t* group_prefix;    // [n_local + 1] exclusive prefix of groups per local barcode
    int64_t n_local;
    const int32_t* donor_a;         // [n_barcodes_total] by global barcode id
    const int32_t* donor_b;         // -1: singlet
    const int8_t* dosage;           // [n_snps, n_genotypes] alt-allele dosage 0 / 1 / 2
    const uint8_t* ref_base;        // [n_snps]
    const uint8_t* alt_base;        // [n_snps]
    const int32_t* snp_position;    // [n_snps]
    const float* err_table;         // [12] single-read (3) and two-read (9) error probabilities
    const int32_t* flip_threshold;  // [12] floor(min(err, 0.04) * 2^20)
};

__device__ inline int64_t find_local_barcode(const int64_t* prefix, int64_t n_local, int64_t group) {
    int64_t lo = 0, hi = n_local;  // last k with prefix[k] <= group
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (prefix[mid] <= group) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ inline int molecules_in_group(uint64_t g2) {
    const uint32_t m = (uint32_t)(g2 >> 32) & 0xffffu;  // 1 + Geom(0.6), capped at 4
    return m < 39322u ? 1 : (m < 55050u ? 2 : (m < 61342u ? 3 : 4));
}

__global__ void synth_count_kernel(SynthParams p, int64_t n_groups, int32_t* __restrict__ molecules) {
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = find_local_barcode(p.group_prefix, p.n_local, g);
        const uint64_t b = (uint64_t)p.barcode_ids[k];
        molecules[g] = molecules_in_group(draw(p.seed, b, (uint64_t)(g - p.group_prefix[k]), 1));
    }
}

__global__ void synth_emit_kernel(SynthParams p, int64_t n_groups, const int64_t* __restrict__ call_offset,
                                  int32_t* __restrict__ out_pos, uint8_t* __restrict__ out_base,
                                  float* __restrict__ out_e, int32_t* __restrict__ out_cb,
                                  int64_t* __restrict__ out_key) {
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = find_local_barcode(p.group_prefix, p.n_local, g);
        const uint64_t b = (uint64_t)p.barcode_ids[k];
        const uint64_t grp = (uint64_t)(g - p.group_prefix[k]);
        const uint64_t g1 = draw(p.seed, b, grp, 0), g2 = draw(p.seed, b, grp, 1);
        // expression skew: product of three uniforms, then an affine permutation of the SNP ranks
        uint64_t t = ((g1 & 0xffffffffull) * (g1 >> 32)) >> 32;
        t = (t * (g2 & 0xffffffffull)) >> 32;
        const uint64_t rank = (t * (uint64_t)p.n_snps) >> 32;
        const int64_t snp = (int64_t)((rank * 2654435761ull + 12345ull) % (uint64_t)p.n_snps);
        const int m = molecules_in_group(g2);
        const int32_t da = p.donor_a[b], db = p.donor_b[b];
        const int64_t first = call_offset[g];
        for (int j = 0; j < m; ++j) {
            const uint64_t c = draw(p.seed, b, grp, 2 + (uint64_t)j);
            const uint64_t c2 = draw(p.seed, b, grp, 64 + (uint64_t)j);
            const uint32_t f0 = (uint32_t)c & 0xffffu, f1 = (uint32_t)(c >> 16) & 0xffffu;
            const uint32_t f2 = (uint32_t)(c >> 32) & 0xffffu, f3 = (uint32_t)(c >> 48) & 0xffffu;
            const int32_t donor = (db >= 0 && f1 < 32768u) ? db : da;
            const uint32_t dose = (uint32_t)p.dosage[snp * p.n_genotypes + donor];
            int base = f0 < dose * 32768u ? p.alt_base[snp] : p.ref_base[snp];
            const int q1 = f2 < 3277u ? 0 : (f2 < 9830u ? 1 : 2);
            const int err_index = f3 < 19661u ? 3 + 3 * q1 + (int)(f3 % 3u) : q1;
            if ((int32_t)(c2 & 0xfffffull) < p.flip_threshold[err_index])
                base = (base + 1 + (int)(((c2 >> 20) & 0xffull) % 3ull)) & 3;
            if (((c2 >> 28) & 0xffffull) < 131ull) base = 4;  // 'N'
            const int32_t off_target = ((c2 >> 44) & 0xffffull) < 1311ull ? 1 : 0;
            out_pos[first + j] = p.snp_position[snp] + off_target;
            out_base[first + j] = (uint8_t)base;
            out_e[first + j] = p.err_table[err_index];
            out_cb[first + j] = (int32_t)b;
            out_key[first + j] = (int64_t)(mix64(c ^ 0x9e3779b97f4a7c15ull) >> 1);  // order of the calls in the "file"
        }
    }
}

// records in the order given by `perm`: one molecule per call (molecule_index = position of the call)
__global__ void synth_pack_kernel(const int64_t* __restrict__ perm, int64_t n, const int32_t* __restrict__ pos,
                                  const uint8_t* __restrict__ base, const float* __restrict__ e,
                                  const int32_t* __restrict__ cb, uint8_t* __restrict__ records,
                                  int32_t* __restrict__ molecule_cb) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t src = perm[k];
        uint8_t* rec = records + 13 * k;
        const uint32_t mol = (uint32_t)k, ps = (uint32_t)pos[src], eb = __float_as_uint(e[src]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            rec[i] = (uint8_t)(mol >> (8 * i));
            rec[4 + i] = (uint8_t)(ps >> (8 * i));
            rec[9 + i] = (uint8_t)(eb >> (8 * i));
        }
        rec[8] = base[src];
        molecule_cb[k] = cb[src];
    }
}

inline int grid_for(int64_t n) {
    int64_t blocks = (n + 255) / 256;
    if (blocks < 1) blocks = 1;
    return (int)(blocks < 148 * 32 ? blocks : 148 * 32);
}

}  // namespace

extern "C" {

int dmxs_count(uint64_t seed, int64_t n_snps, int32_t n_genotypes, const int32_t* barcode_ids,
               const int64_t* group_prefix, int64_t n_local, int64_t n_groups, int32_t* molecules, void* stream) {
    SynthParams p{};
    p.seed = seed; p.n_snps = n_snps; p.n_genotypes = n_genotypes; p.barcode_ids = barcode_ids;
    p.group_prefix = group_prefix; p.n_local = n_local;
    if (n_groups > 0) synth_count_kernel<<<grid_for(n_groups), 256, 0, (cudaStream_t)stream>>>(p, n_groups, molecules);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int dmxs_emit(uint64_t seed, int64_t n_snps, int32_t n_genotypes, const int32_t* barcode_ids,
              const int64_t* group_prefix, int64_t n_local, int64_t n_groups, const int32_t* donor_a,
              const int32_t* donor_b, const int8_t* dosage, const uint8_t* ref_base, const uint8_t* alt_base,
              const int32_t* snp_position, const float* err_table, const int32_t* flip_threshold,
              const int64_t* call_offset, int32_t* out_pos, uint8_t* out_base, float* out_e, int32_t* out_cb,
              int64_t* out_key, void* stream) {
    SynthParams p{};
    p.seed = seed; p.n_snps = n_snps; p.n_genotypes = n_genotypes; p.barcode_ids = barcode_ids;
    p.group_prefix = group_prefix; p.n_local = n_local; p.donor_a = donor_a; p.donor_b = donor_b; p.dosage = dosage;
    p.ref_base = ref_base; p.alt_base = alt_base; p.snp_position = snp_position; p.err_table = err_table;
    p.flip_threshold = flip_threshold;
    if (n_groups > 0)
        synth_emit_kernel<<<grid_for(n_groups), 256, 0, (cudaStream_t)stream>>>(p, n_groups, call_offset, out_pos, out_base,
                                                                               out_e, out_cb, out_key);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int dmxs_pack(const int64_t* perm, int64_t n, const int32_t* pos, const uint8_t* base, const float* e,
              const int32_t* cb, uint8_t* records, int32_t* molecule_cb, void* stream) {
    if (n > 0) synth_pack_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(perm, n, pos, base, e, cb, records, molecule_cb);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
