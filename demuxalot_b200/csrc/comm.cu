// (f) of north_star: the M-step's partial variant x genotype sums of all barcode shards are combined over
// NVLink once per EM iteration.  dmx_mstep_allreduce runs the M-step kernel over tiles of the variant range and
// hands every finished tile to NCCL on a second stream, so the collective of tile k overlaps the kernel of tile k + 1.
//
// Wire formats:
//   float64: reduce-scatter of the unrounded float64 partials, each rank rounds ITS slice of the global sums to float32
//            once (the place where the single-GPU path rounds: N GPUs give the bits of one, up to float64 regrouping),
//            all-gather of the float32 slices -- 3/4 of the bytes of a float64 all-reduce and no rounding pass over
//            the whole table;
//   float32: in-place all-reduce of float32 partials (half the bytes again, one extra rounding per shard).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded by the host process, e.g. by torch), so
// libdemux_b200.so has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace dmx {

// the handful of NCCL declarations used here (nccl.h 2.x; values are ABI-stable)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_SUCCESS = 0, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct NcclApi {
    int (*GetUniqueId)(ncclUniqueId*);
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(int);
    bool ok;
};

static NcclApi* nccl_api() {
    static NcclApi api = {};
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host process already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
#define DMX_SYM(field, name) *(void**)(&api.field) = dlsym(h, name); if (!api.field) return nullptr
    DMX_SYM(GetUniqueId, "ncclGetUniqueId");
    DMX_SYM(CommInitRank, "ncclCommInitRank");
    DMX_SYM(CommDestroy, "ncclCommDestroy");
    DMX_SYM(AllReduce, "ncclAllReduce");
    DMX_SYM(ReduceScatter, "ncclReduceScatter");
    DMX_SYM(AllGather, "ncclAllGather");
    DMX_SYM(GetErrorString, "ncclGetErrorString");
#undef DMX_SYM
    api.ok = true;
    return &api;
}

#define DMX_NCCL(call)                                                                              \
    do {                                                                                            \
        const int rc_ = (call);                                                                     \
        if (rc_ != NCCL_SUCCESS) {                                                                  \
            set_error("%s failed: %s", #call, nccl_api()->GetErrorString(rc_));                     \
            return -3;                                                                              \
        }                                                                                           \
    } while (0)

struct Comm {
    ncclComm_t comm;
    cudaStream_t stream;       // collectives run here, beside the M-step kernels of the caller's stream
    cudaEvent_t tile_done[2];  // compute -> comm hand-over (alternating)
    cudaEvent_t all_done;      // comm -> compute
    int rank, world;
};

// ---- cross-GPU sum over peer memory (NVLink / NVSwitch), no library collective ---------------------------------------
// Every rank's M-step leaves its float32 partial [v_pad, G] in a buffer that all ranks of the node have mapped (CUDA IPC /
// fabric handles; the host maps them).  Rank r owns slice r of the flattened table: it loads that slice from every rank's
// partial (peer loads), adds the values in rank order in float64, rounds once and stores the float32 sum into slice r of
// EVERY rank's output table (peer stores) -- a reduce-scatter and an all-gather in one pass, deterministic, with the
// same result bits on all ranks.  The two barriers around it are the host's (device-side signal pads).
struct PeerPointers {
    const float* in[16];
    float* out[16];
};

template <int WORLD>
__global__ void __launch_bounds__(256) peer_sum_kernel(const PeerPointers p, int rank, int64_t slice_quads) {
    const int64_t base = (int64_t)rank * slice_quads;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < slice_quads;
         q += (int64_t)gridDim.x * blockDim.x) {
        float4 v[WORLD];
#pragma unroll
        for (int r = 0; r < WORLD; ++r) v[r] = __ldcg(reinterpret_cast<const float4*>(p.in[r]) + base + q);
        double x = 0.0, y = 0.0, z = 0.0, w = 0.0;
#pragma unroll
        for (int r = 0; r < WORLD; ++r) { x += (double)v[r].x; y += (double)v[r].y; z += (double)v[r].z; w += (double)v[r].w; }
        const float4 sum = make_float4((float)x, (float)y, (float)z, (float)w);
#pragma unroll
        for (int r = 0; r < WORLD; ++r) __stcg(reinterpret_cast<float4*>(p.out[r]) + base + q, sum);
    }
}

__global__ void round_slice_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        out[k] = (float)in[k];
}

}  // namespace dmx

extern "C" {

int dmx_comm_unique_id(uint8_t* h_id128) {
    using namespace dmx;
    NcclApi* api = nccl_api();
    DMX_REQUIRE(api, "libnccl.so.2 not found: multi-GPU EM needs NCCL");
    ncclUniqueId id;
    DMX_NCCL(api->GetUniqueId(&id));
    memcpy(h_id128, id.internal, 128);
    return 0;
}

int dmx_comm_init(const uint8_t* h_id128, int32_t rank, int32_t world, void** h_comm) {
    using namespace dmx;
    NcclApi* api = nccl_api();
    DMX_REQUIRE(api, "libnccl.so.2 not found: multi-GPU EM needs NCCL");
    DMX_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", (int)rank, (int)world);
    ncclUniqueId id;
    memcpy(id.internal, h_id128, 128);
    Comm* c = new Comm();
    c->rank = rank;
    c->world = world;
    DMX_NCCL(api->CommInitRank(&c->comm, world, id, rank));
    DMX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) DMX_CUDA(cudaEventCreateWithFlags(&c->tile_done[k], cudaEventDisableTiming));
    DMX_CUDA(cudaEventCreateWithFlags(&c->all_done, cudaEventDisableTiming));
    *h_comm = c;
    return 0;
}

int dmx_comm_destroy(void* comm) {
    using namespace dmx;
    if (!comm) return 0;
    Comm* c = (Comm*)comm;
    cudaStreamSynchronize(c->stream);
    nccl_api()->CommDestroy(c->comm);
    for (int k = 0; k < 2; ++k) cudaEventDestroy(c->tile_done[k]);
    cudaEventDestroy(c->all_done);
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int dmx_peer_sum_f32(const void* const* h_partials, void* const* h_outputs, int32_t rank, int32_t world,
                     int64_t n_elements, void* stream) {
    using namespace dmx;
    DMX_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, "bad rank %d / world %d (<= 16)", (int)rank, (int)world);
    DMX_REQUIRE(n_elements >= 0 && n_elements % (4 * (int64_t)world) == 0,
                "n_elements %lld must be a multiple of 4 * world", (long long)n_elements);
    if (n_elements == 0) return 0;
    PeerPointers p = {};
    for (int r = 0; r < world; ++r) {
        DMX_REQUIRE(h_partials[r] && h_outputs[r], "null peer pointer for rank %d", r);
        DMX_REQUIRE(((uintptr_t)h_partials[r] & 15) == 0 && ((uintptr_t)h_outputs[r] & 15) == 0, "peer buffers must be 16-byte aligned");
        p.in[r] = (const float*)h_partials[r];
        p.out[r] = (float*)h_outputs[r];
    }
    const int64_t slice_quads = n_elements / 4 / world;
    int64_t blocks = ceil_div(slice_quads, 256);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t s = (cudaStream_t)stream;
#define DMX_PEER(W) case W: peer_sum_kernel<W><<<(int)blocks, 256, 0, s>>>(p, rank, slice_quads); break
    switch (world) {
        DMX_PEER(1); DMX_PEER(2); DMX_PEER(3); DMX_PEER(4); DMX_PEER(5); DMX_PEER(6); DMX_PEER(7); DMX_PEER(8);
        DMX_PEER(9); DMX_PEER(10); DMX_PEER(11); DMX_PEER(12); DMX_PEER(13); DMX_PEER(14); DMX_PEER(15); DMX_PEER(16);
        default: break;
    }
#undef DMX_PEER
    DMX_LAUNCH_CHECK();
    return 0;
}

int64_t dmx_mstep_allreduce_padded_variants(int64_t n_variants, int32_t world) {
    return dmx::round_up(n_variants > 0 ? n_variants : 1, world > 0 ? world : 1);
}

int dmx_mstep_allreduce(const int64_t* variant_offsets, const int32_t* csc_cb, const float* csc_e,
                        const float* singlet_posteriors, int64_t ld_singlet, int32_t n_genotypes, double power,
                        float* addition, double* partial64, double* slice64, int64_t n_variants, const void* plan,
                        int64_t n_rows, int64_t n_medium, int64_t n_heavy_variants, int64_t n_heavy_items,
                        double* heavy_scratch, void* comm, int32_t n_tiles, int32_t wire_float64, void* stream_) {
    using namespace dmx;
    NcclApi* api = nccl_api();
    DMX_REQUIRE(api && comm, "no communicator");
    Comm* c = (Comm*)comm;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t G = n_genotypes;
    const int64_t v_pad = dmx_mstep_allreduce_padded_variants(n_variants, c->world);
    DMX_REQUIRE(!wire_float64 || (partial64 && slice64), "float64 wire format needs partial64 and slice64");
    if (n_tiles < 1) n_tiles = 1;
    // tile boundaries are multiples of world so that every tile splits into equal per-rank slices; rows in
    // [n_variants, v_pad) of `addition` / `partial64` are padding (the caller allocates and zeroes them once)
    const int64_t rows_per_tile = round_up(ceil_div(v_pad, n_tiles), c->world);
    int k = 0;
    for (int64_t v_lo = 0; v_lo < v_pad; v_lo += rows_per_tile, ++k) {
        const int64_t v_hi = v_lo + rows_per_tile < v_pad ? v_lo + rows_per_tile : v_pad;
        const int64_t k_hi = v_hi < n_variants ? v_hi : n_variants;
        if (v_lo < k_hi) {
            int rc;
            if (plan)
                rc = dmx_mstep_planned(variant_offsets, csc_cb, csc_e, singlet_posteriors, ld_singlet, n_genotypes, power,
                                       wire_float64 ? nullptr : addition, G, wire_float64 ? partial64 : nullptr, G, v_lo,
                                       k_hi, plan, n_rows, n_medium, n_heavy_variants, n_heavy_items, heavy_scratch,
                                       stream_);
            else
                rc = dmx_mstep(variant_offsets, csc_cb, csc_e, singlet_posteriors, ld_singlet, n_genotypes, power,
                               wire_float64 ? nullptr : addition, G, wire_float64 ? partial64 : nullptr, G, v_lo, k_hi,
                               stream_);
            if (rc) return rc;
        }
        cudaEvent_t ev = c->tile_done[k & 1];
        DMX_CUDA(cudaEventRecord(ev, stream));
        DMX_CUDA(cudaStreamWaitEvent(c->stream, ev, 0));
        const int64_t count = (v_hi - v_lo) * G;
        const int64_t slice = count / c->world;  // exact: v_hi - v_lo is a multiple of world
        if (wire_float64) {
            DMX_NCCL(api->ReduceScatter(partial64 + v_lo * G, slice64, (size_t)slice, NCCL_FLOAT64, NCCL_SUM, c->comm,
                                        c->stream));
            float* mine = addition + v_lo * G + (int64_t)c->rank * slice;
            int64_t blocks = ceil_div(slice, 256);
            if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
            round_slice_kernel<<<(int)blocks, 256, 0, c->stream>>>(slice64, mine, slice);
            DMX_LAUNCH_CHECK();
            DMX_NCCL(api->AllGather(mine, addition + v_lo * G, (size_t)slice, NCCL_FLOAT32, c->comm, c->stream));
        } else {
            DMX_NCCL(api->AllReduce(addition + v_lo * G, addition + v_lo * G, (size_t)count, NCCL_FLOAT32, NCCL_SUM,
                                    c->comm, c->stream));
        }
    }
    DMX_CUDA(cudaEventRecord(c->all_done, c->stream));
    DMX_CUDA(cudaStreamWaitEvent(stream, c->all_done, 0));
    return 0;
}

}  // extern "C"
