// Developer microbenchmark: issue rate and dependent-issue latency of the packed FP32 forms of sm_100
// (fma.rn.f32x2 -> SASS FFMA2) against scalar FFMA, as a function of the independent chains per thread
// (ILP) and the resident warps per scheduler.  Answers the question the two-envs-per-thread step kernel
// raises: does packing two envs halve the FP32 issue slots, and at what occupancy / ILP?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/ffma2_rates tools/microbench/ffma2_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: scalar FFMA (2 per chain step), 1: FFMA2 (one register-pair operand, two uniform operands),
// 2: FMUL2 / FADD2 alternating, 3: scalar FFMA with THREE register operands, 4: FFMA2 with three
// register-pair operands, 5: FFMA2 with two register-pair operands and one uniform
template <int ILP, int MODE>
__global__ void rate_kernel(float* out, int iters, float a, float b, long long* cycles) {
    float2 acc[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc[j] = make_float2(threadIdx.x * 1e-3f + j, threadIdx.x * 2e-3f - j);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (MODE == 0) { acc[j].x = fmaf(acc[j].x, a, b); acc[j].y = fmaf(acc[j].y, a, b); }
            else if (MODE == 1) acc[j] = __ffma2_rn(acc[j], A, B);
            else if (MODE == 2) acc[j] = (i & 1) ? __fmul2_rn(acc[j], A) : __fadd2_rn(acc[j], B);
            else if (MODE == 3) { acc[j].x = fmaf(acc[j].x, acc[(j + 1) % ILP].x, acc[(j + 2) % ILP].x); acc[j].y = fmaf(acc[j].y, acc[(j + 1) % ILP].y, acc[(j + 2) % ILP].y); }
            else if (MODE == 4) acc[j] = __ffma2_rn(acc[j], acc[(j + 1) % ILP], acc[(j + 2) % ILP]);
            else acc[j] = __ffma2_rn(acc[j], acc[(j + 1) % ILP], B);
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP, int MODE>
void run(int ctas_per_sm, int sms, float* out, long long* cyc) {
    const int iters = 4096, grid = sms * ctas_per_sm;
    rate_kernel<ILP, MODE><<<grid, 128>>>(out, iters, 0.999f, 1e-3f, cyc);
    cudaDeviceSynchronize();
    rate_kernel<ILP, MODE><<<grid, 128>>>(out, iters, 0.999f, 1e-3f, cyc);
    cudaDeviceSynchronize();
    long long h[4096], mx = 0;
    cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    // per scheduler: ctas_per_sm warps (a 128-thread CTA puts one warp on each of the 4 schedulers)
    const double fma_elems = (double)iters * ILP * 2 * 32 * ctas_per_sm;       // element FMAs per scheduler
    const double instr = (double)iters * ILP * ((MODE == 0 || MODE == 3) ? 2 : 1) * ctas_per_sm;
    printf("%-8s ILP %d warps/sched %2d: %9lld cycles, %.2f cycles/instr/sched, %.1f element-FMA/clk/sched\n",
           MODE == 0 ? "FFMA" : MODE == 1 ? "FFMA2" : MODE == 2 ? "FMUL2/FADD2" : MODE == 3 ? "FFMA-3reg" : MODE == 4 ? "FFMA2-3reg" : "FFMA2-2reg", ILP, ctas_per_sm, mx, mx / instr, fma_elems / mx);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * 4096 * 128); cudaMalloc(&cyc, sizeof(long long) * 4096);
    for (int w : {1, 2, 4, 8, 12}) {
        run<1, 0>(w, sms, out, cyc); run<1, 1>(w, sms, out, cyc);
        run<4, 0>(w, sms, out, cyc); run<4, 1>(w, sms, out, cyc); run<4, 2>(w, sms, out, cyc);
        run<8, 0>(w, sms, out, cyc); run<8, 1>(w, sms, out, cyc);
        run<8, 3>(w, sms, out, cyc); run<8, 4>(w, sms, out, cyc); run<8, 5>(w, sms, out, cyc);
    }
    return 0;
}
