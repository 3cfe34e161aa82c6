// Issue-rate microbenchmark shaped like the pair E-step's inner loop (sm_100a): which ceiling bounds
//     prod[c] = prod[c] * (a_i + a_j)          (one add + one multiply per update, packed f32x2 in the kernel)?
// Every chain is loop-carried (the add's input is its own previous output, the multiply consumes the fresh sum), so
// neither nvcc nor ptxas can hoist, merge or drop anything: the SASS instruction count per iteration equals the nominal
// count (checked with cuobjdump, listing in profiles/r02_microbench_sass_counts.txt).  All values stay at exactly 1.0 /
// 0.0 + 1.0 (no denormals, no infinities).
//   mode 0: FADD2 -> FMUL2, dependent, 1 : 1        (the kernel's pair)
//   mode 1: FMUL2 only      mode 2: FADD2 only      mode 4: FFMA2 only
//   mode 3: scalar FADD -> FMUL, dependent, 1 : 1   mode 5: scalar FFMA only
//   mode 6: mode 0 plus ONE independent integer LOP3 per packed pair (do ALU instructions ride along for free?)
//   mode 7: mode 0 plus TWO LOP3 per packed pair
// Sweeps warps per SM (one warp per CTA, like the warp pair kernel).  Rates are by wall time (CUDA events) at the SM
// clock read from NVML-free cudaDeviceGetAttribute(clockRate) -- printed -- and by clock64 of the slowest warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_packed_tile microbench_packed_tile.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float sadd(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float smul(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float sfma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ unsigned lop3(unsigned a, unsigned b, unsigned c) { unsigned d; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

constexpr int REPS = 8;  // unrolled repetitions of the chain set per loop iteration

template <int MODE, int NCH>
__global__ void __launch_bounds__(32) bench(float* out, const float* in, int iters, long long* cycles) {
    uint64_t prod[NCH], sum[NCH];
    float sprod[NCH], ssum[NCH];
    unsigned bits[NCH];
    const uint64_t one2 = pack2(in[0], in[0]);    // 1.0
    const uint64_t zero2 = pack2(in[1], in[1]);   // 0.0
    const float one = in[0], zero = in[1];
    const unsigned m0 = __float_as_uint(in[2]), m1 = __float_as_uint(in[3]);
    for (int c = 0; c < NCH; ++c) { prod[c] = one2; sum[c] = one2; sprod[c] = one; ssum[c] = one; bits[c] = m0 + c + threadIdx.x; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < REPS; ++rep) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (MODE == 0 || MODE == 6 || MODE == 7) { sum[c] = add2(sum[c], zero2); prod[c] = mul2(prod[c], sum[c]); }
                if (MODE == 6 || MODE == 7) bits[c] = lop3(bits[c], m0, m1);
                if (MODE == 7) bits[c] = lop3(bits[c], m1, m0);
                if (MODE == 1) prod[c] = mul2(prod[c], one2);
                if (MODE == 2) sum[c] = add2(sum[c], zero2);
                if (MODE == 4) prod[c] = fma2(prod[c], one2, zero2);
                if (MODE == 3) { ssum[c] = sadd(ssum[c], zero); sprod[c] = smul(sprod[c], ssum[c]); }
                if (MODE == 5) sprod[c] = sfma(sprod[c], one, zero);
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int c = 0; c < NCH; ++c) {
        float lo, hi, l2, h2;
        unpack2(prod[c], lo, hi);
        unpack2(sum[c], l2, h2);
        s += lo + hi + l2 + h2 + sprod[c] + ssum[c] + __uint_as_float(bits[c] & 0x3f800000u);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int NCH>
void run(const char* name, int warps_per_sm, int fp_instr_per_chain, int other_instr_per_chain, int lanes_per_instr,
         const float* d_in, double sm_hz) {
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
    long long slowest = 0; for (int i = 0; i < blocks; ++i) slowest = h[i] > slowest ? h[i] : slowest;
    const double fp_instr = (double)iters * REPS * NCH * fp_instr_per_chain;       // per warp
    const double all_instr = fp_instr + (double)iters * REPS * NCH * other_instr_per_chain;
    const double fp_rate_wall = fp_instr * warps_per_sm / (ms * 1e-3 * sm_hz);
    const double all_rate_wall = all_instr * warps_per_sm / (ms * 1e-3 * sm_hz);
    const double fp_rate_clk = fp_instr * warps_per_sm / (double)slowest;
    printf("%-30s chains %2d warps/SM %2d  %.3f ms  FP %6.3f warp-instr/clk/SM (wall)  %6.3f (clock64 of the slowest warp)  "
           "all %6.3f  -> %6.1f FP32 lane-ops/clk/SM\n",
           name, NCH, warps_per_sm, ms, fp_rate_wall, fp_rate_clk, all_rate_wall, fp_rate_wall * lanes_per_instr * 32);
    delete[] h; cudaFree(out); cudaFree(cyc);
}

int main() {
    float h_in[4] = {1.f, 0.f, 0.f, 0.f};
    float* d_in; cudaMalloc(&d_in, sizeof(h_in));
    cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double sm_hz = khz * 1e3;
    printf("SM clock used for the wall-time rates: %.0f MHz (cudaDevAttrClockRate)\n", sm_hz / 1e6);
    const int warps[] = {4, 8, 12, 16, 32};
    for (int w : warps) {
        run<0, 32>("FADD2 -> FMUL2 (1:1, dependent)", w, 2, 0, 2, d_in, sm_hz);
        run<1, 32>("FMUL2", w, 1, 0, 2, d_in, sm_hz);
        run<2, 32>("FADD2", w, 1, 0, 2, d_in, sm_hz);
        run<4, 32>("FFMA2", w, 1, 0, 2, d_in, sm_hz);
        run<3, 32>("FADD -> FMUL scalar (1:1)", w, 2, 0, 1, d_in, sm_hz);
        run<5, 32>("FFMA scalar", w, 1, 0, 1, d_in, sm_hz);
        run<6, 32>("FADD2 -> FMUL2 + 1 LOP3", w, 2, 1, 2, d_in, sm_hz);
        run<7, 32>("FADD2 -> FMUL2 + 2 LOP3", w, 2, 2, 2, d_in, sm_hz);
    }
    return 0;
}
