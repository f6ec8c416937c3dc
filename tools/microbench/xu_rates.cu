// Developer microbenchmark: per-SM throughput of the special-function forms a tanh MLP could use
// (results in profiles/; build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o xu_rates xu_rates.cu).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    float a[8]; uint32_t h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = seed + 0.01f * (threadIdx.x + j); h[j] = 0x3c003800u + threadIdx.x + j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[j]));
            if (MODE == 1) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h[j]));
            if (MODE == 2) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[j]));
            if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
            if (MODE == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
            if (MODE == 5) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[j]));            // 1 MUFU + 4 FFMA
                             float b = a[j]; b = fmaf(b, 1.0001f, 0.1f); b = fmaf(b, 0.999f, -0.1f); b = fmaf(b, 1.0001f, 0.1f); a[j] = fmaf(b, 0.999f, -0.1f); }
            if (MODE == 6) { float b = a[j]; b = fmaf(b, 1.0001f, 0.1f); b = fmaf(b, 0.999f, -0.1f); b = fmaf(b, 1.0001f, 0.1f); a[j] = fmaf(b, 0.999f, -0.1f); }  // 4 FFMA
            if (MODE == 7) { asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(h[j])); asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(h[j]));
                             asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(h[j])); asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(h[j])); }  // 4 HFMA2
            if (MODE == 8) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[j]));
                             uint32_t p; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(a[j]), "f"(a[(j + 1) & 7])); h[j] ^= p; }   // MUFU + F2FP
        }
    }
    float s = 0; uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += a[j]; x ^= h[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)x;
}

template <int MODE>
void run(const char* name, int ops_per_iter_per_thread) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(out, iters, 0.3f);
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, iters, 0.3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = (double)sms * 8 * 256 * iters * ops_per_iter_per_thread;
    printf("%-34s %8.3f ms  %7.2f lane-ops/ns/SM  (= %.2f per clk per SM at %.0f MHz nominal)\n", name, ms,
           lane_ops / (ms * 1e6) / sms, lane_ops / (ms * 1e6) / sms / (clk_khz * 1e-6), clk_khz * 1e-3);
    cudaFree(out);
}

int main() {
    run<0>("tanh.approx.f32", 8);
    run<1>("tanh.approx.bf16x2 (instr)", 8);
    run<2>("tanh.approx.f16x2 (instr)", 8);
    run<3>("ex2.approx.ftz.f32", 8);
    run<4>("rcp.approx.ftz.f32", 8);
    run<5>("tanh.f32 + 4 FFMA (MUFU count)", 8);
    run<6>("4 FFMA (FFMA count)", 32);
    run<7>("4 HFMA2.f16x2 (instr count)", 32);
    run<8>("tanh.f32 + cvt.bf16x2 (MUFU count)", 8);
    return 0;
}
