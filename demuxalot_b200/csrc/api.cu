// Boundary plumbing: ABI version, thread-local error text, device info, small utility kernels.
#include <stdarg.h>

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
