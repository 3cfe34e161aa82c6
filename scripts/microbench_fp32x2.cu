// Issue-rate microbenchmark for the FP32 forms the pair E-step can use on sm_100a:
//   scalar FADD / FMUL / FFMA vs packed FADD2 / FMUL2 / FFMA2 (add/mul/fma.f32x2), MUFU.LG2, F2F.F64.F32 + DADD.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp32x2 microbench_fp32x2.cu
// Prints warp-instructions per clock per SM and the lane throughput for each form (16 warps per scheduler).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ float lo_of(uint64_t v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo + hi; }

template <int MODE>
__global__ void bench(float* out, float seed, long long* cycles) {
    float x[CHAINS];
    uint64_t x2[CHAINS];
    double d[CHAINS];
    for (int c = 0; c < CHAINS; ++c) { x[c] = seed + c; x2[c] = pack2(seed + c, seed - c); d[c] = 0.0; }
    const float k = seed * 0.5f;
    const uint64_t k2 = pack2(k, k + 1.f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(k));
            if (MODE == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(k));
            if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[c]) : "f"(k));
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x2[c]) : "l"(k2));
            if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x2[c]) : "l"(k2));
            if (MODE == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x2[c]) : "l"(k2));
            if (MODE == 6) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
            if (MODE == 7) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(x[c])); asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(t)); }
            if (MODE == 8) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"((double)k));
            if (MODE == 9) {  // the inner-loop pair: packed add feeding a packed multiply
                uint64_t t; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(k2), "l"(x2[(c + 1) % CHAINS]));
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x2[c]) : "l"(t));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int c = 0; c < CHAINS; ++c) s += x[c] + lo_of(x2[c]) + (float)d[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_chain_step, int lanes_per_instr) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 512, blocks_per_sm = 2;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * threads * sms * blocks_per_sm);
    cudaMalloc(&cyc, sizeof(long long) * sms * blocks_per_sm);
    bench<MODE><<<sms * blocks_per_sm, threads>>>(out, 1.0001f, cyc);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<MODE><<<sms * blocks_per_sm, threads>>>(out, 1.0001f, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[4096]; cudaMemcpy(h, cyc, sizeof(long long) * sms * blocks_per_sm, cudaMemcpyDeviceToHost);
    double mean_cycles = 0; for (int i = 0; i < sms * blocks_per_sm; ++i) mean_cycles += (double)h[i]; mean_cycles /= sms * blocks_per_sm;
    const double warp_instr_per_sm = (double)ITERS * CHAINS * instr_per_chain_step * (threads / 32) * blocks_per_sm;
    printf("%-22s %8.3f ms  %10.0f cycles/block  %6.3f warp-instr/clk/SM  %7.1f lane-ops/clk/SM\n", name, ms, mean_cycles,
           warp_instr_per_sm / mean_cycles, warp_instr_per_sm / mean_cycles * lanes_per_instr);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FADD", 1, 32);
    run<1>("FMUL", 1, 32);
    run<2>("FFMA", 1, 32);
    run<3>("FADD2 (add.f32x2)", 1, 64);
    run<4>("FMUL2 (mul.f32x2)", 1, 64);
    run<5>("FFMA2 (fma.f32x2)", 1, 64);
    run<6>("MUFU.LG2", 1, 32);
    run<7>("F2F.F64.F32 + DADD", 2, 32);
    run<8>("DADD", 1, 32);
    run<9>("FADD2 -> FMUL2 pair", 2, 64);
    return 0;
}
