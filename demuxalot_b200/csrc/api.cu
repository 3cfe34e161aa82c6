// Boundary plumbing: ABI version, thread-local error text, device info, small utility kernels.
#include <stdarg.h>

#include <thread>
#include <vector>

#include <string.h>

#include "common.cuh"

namespace dmx {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
    va_list args;
    va_start(args, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, args);
    va_end(args);
}

__global__ void round_f64_to_f32_kernel(const double* __restrict__ in, int64_t ld_in, float* __restrict__ out,
                                        int64_t ld_out, int64_t n_rows, int n_cols) {
    const int64_t total = n_rows * n_cols;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = k / n_cols;
        const int c = (int)(k - r * n_cols);
        out[r * ld_out + c] = (float)in[r * ld_in + c];
    }
}

}  // namespace dmx

extern "C" {

int dmx_abi_version(void) { return DMX_ABI_VERSION; }

const char* dmx_last_error(void) { return dmx::g_error; }

int dmx_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes,
                    int64_t* total_mem_bytes) {
    cudaDeviceProp prop;
    DMX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (l2_bytes) *l2_bytes = prop.l2CacheSize;
    if (total_mem_bytes) *total_mem_bytes = (int64_t)prop.totalGlobalMem;
    return 0;
}

int dmx_host_gather_cb(const uint8_t* h_molecules_packed, int64_t n_molecules, int32_t* h_out_cb, int32_t n_threads) {
    if (n_molecules <= 0) return 0;
    DMX_REQUIRE(h_molecules_packed && h_out_cb, "null host pointer");
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    const int64_t min_per_thread = 1 << 16;
    if ((int64_t)n_threads * min_per_thread > n_molecules) n_threads = (int32_t)(n_molecules / min_per_thread) + 1;
    auto work = [=](int64_t lo, int64_t hi) {
        const uint8_t* src = h_molecules_packed + 12 * lo;  // record = (compressed_cb i4, compressed_ub i4, p f4)
        for (int64_t k = lo; k < hi; ++k, src += 12) {
            int32_t cb;
            memcpy(&cb, src, 4);
            h_out_cb[k] = cb;
        }
    };
    if (n_threads == 1) {
        work(0, n_molecules);
        return 0;
    }
    std::vector<std::thread> pool;
    pool.reserve(n_threads);
    const int64_t per = (n_molecules + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const int64_t lo = t * per, hi = lo + per < n_molecules ? lo + per : n_molecules;
        if (lo < hi) pool.emplace_back(work, lo, hi);
    }
    for (auto& th : pool) th.join();
    return 0;
}

int dmx_round_f64_to_f32(const double* in64, int64_t ld_in, float* out, int64_t ld_out, int64_t n_rows,
                         int32_t n_cols, void* stream) {
    if (n_rows <= 0 || n_cols <= 0) return 0;
    const int64_t total = n_rows * n_cols;
    const int threads = 256;
    int64_t blocks = dmx::ceil_div(total, threads);
    if (blocks > (int64_t)dmx::sm_count() * 16) blocks = (int64_t)dmx::sm_count() * 16;
    dmx::round_f64_to_f32_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(in64, ld_in, out, ld_out, n_rows, n_cols);
    DMX_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
