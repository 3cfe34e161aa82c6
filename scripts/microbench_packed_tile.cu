// Issue-rate microbenchmark shaped like the pair E-step's inner loop (sm_100a): every thread keeps NCH packed
// running products and applies  prod[c] = prod[c] * (a[c % 4] + b[c / 4])  with operands that keep all values at
// exactly 1.0 (no denormals / infinities, which distort the older microbench_fp32x2 numbers).
//   mode 0: FADD2 (scalar operand broadcast by the instruction) + FMUL2     -- the kernel's loop; the 8 x NCH adds of
//           an iteration take distinct register pairs (16 packed x 16 scalar operands) so that ptxas cannot merge them (it did in
//           the version that produced profiles/r01_microbench_packed_tile.log: SASS had 32 FADD2 per 256 FMUL2, so the
//           mode-0 rates of that log must be scaled by 288 / 512 -- 3.51 -> 1.97 warp-instructions/clk/SM)
//   mode 1: FMUL2 only            mode 2: FADD2 only            mode 3: scalar FADD + FMUL (2 x NCH chains)
//   mode 4: FFMA2 only
// Sweeps warps per SM (one warp per CTA, like the warp pair kernel) and chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_packed_tile microbench_packed_tile.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE, int NCH>
__global__ void __launch_bounds__(32) bench(float* out, const float* in, int iters, long long* cycles) {
    uint64_t prod[NCH];
    float sprod[2 * NCH];
    for (int c = 0; c < NCH; ++c) { prod[c] = pack2(1.f, 1.f); sprod[2 * c] = 1.f; sprod[2 * c + 1] = 1.f; }
    // operands from memory so nothing is constant-folded: a = 0.5, b = 0.5 -> factor 1.0
    float av[32], bv[16];
    for (int k = 0; k < 32; ++k) av[k] = in[k];
    for (int k = 0; k < 16; ++k) bv[k] = in[32 + k];
    uint64_t a2[16];
    for (int k = 0; k < 16; ++k) a2[k] = pack2(av[2 * k], av[2 * k + 1]);
    const uint64_t one2 = pack2(in[48], in[48]);  // 1.0
    const uint64_t zero2 = pack2(in[49], in[49]);  // 0.0
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (MODE == 0) prod[c] = mul2(prod[c], add2(a2[(rep * NCH + c) % 16], pack2(bv[((rep * NCH + c) / 16) % 16], bv[((rep * NCH + c) / 16) % 16])));
                if (MODE == 1) prod[c] = mul2(prod[c], one2);
                if (MODE == 2) prod[c] = add2(prod[c], zero2);
                if (MODE == 3) {
                    sprod[2 * c] = __fmul_rn(sprod[2 * c], __fadd_rn(av[c % 8], bv[(c / 4) % 8]));
                    sprod[2 * c + 1] = __fmul_rn(sprod[2 * c + 1], __fadd_rn(av[(c + 1) % 8], bv[(c / 4) % 8]));
                }
                if (MODE == 4) prod[c] = fma2(prod[c], one2, zero2);
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int c = 0; c < NCH; ++c) { float lo, hi; unpack2(prod[c], lo, hi); s += lo + hi + sprod[2 * c] + sprod[2 * c + 1]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int NCH>
void run(const char* name, int warps_per_sm, int instr_per_chain, int lanes_per_instr, const float* d_in) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int iters = 2000;
    const int blocks = sms * warps_per_sm;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * 32 * blocks);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    bench<MODE, NCH><<<blocks, 32>>>(out, d_in, 10, cyc);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<MODE, NCH><<<blocks, 32>>>(out, d_in, iters, cyc);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    long long* h = new long long[blocks];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; ++i) mean += (double)h[i]; mean /= blocks;
    const double warp_instr = (double)iters * 8 * NCH * instr_per_chain;  // per warp
    // by the device clock (all warps of an SM run concurrently) and by wall time at the nominal 1.965 GHz
    const double per_clk_sm = warp_instr * warps_per_sm / mean;
    const double per_clk_sm_wall = warp_instr * warps_per_sm / (ms * 1e-3 * 1.965e9);
    printf("%-34s chains %2d  warps/SM %2d  %.3f ms  %7.3f warp-instr/clk/SM (clock64)  %7.3f (wall@1.965GHz)  %7.1f lane-ops/clk/SM\n",
           name, NCH, warps_per_sm, ms, per_clk_sm, per_clk_sm_wall, per_clk_sm * lanes_per_instr * 32);
    delete[] h; cudaFree(out); cudaFree(cyc);
}

int main() {
    float h_in[50];
    for (int k = 0; k < 48; ++k) h_in[k] = 0.5f;
    h_in[48] = 1.f; h_in[49] = 0.f;
    float* d_in; cudaMalloc(&d_in, sizeof(h_in));
    cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    const int warps[] = {4, 8, 12, 16, 24, 32};
    for (int w : warps) {
        run<0, 32>("FADD2(bcast) -> FMUL2", w, 2, 2, d_in);
        run<0, 8>("FADD2(bcast) -> FMUL2", w, 2, 2, d_in);
        run<1, 32>("FMUL2", w, 1, 2, d_in);
        run<2, 32>("FADD2", w, 1, 2, d_in);
        run<4, 32>("FFMA2", w, 1, 2, d_in);
        run<3, 32>("scalar FADD -> FMUL", w, 4, 1, d_in);
    }
    return 0;
}
