// pcie_pipeline.cu -- developer microbenchmark: the copy pattern of copter_step_host_f32 for 2^24
// Lander3D envs (per chunk: actions H2D, a kernel touching the chunk, obs / reward / done D2H)
// under different stream arrangements, with an event timeline.  No library code involved.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/pcie_pipeline tools/microbench/pcie_pipeline.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void touch(const float4* act, float* obs, float* rew, unsigned char* done, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = act[i];
    for (int j = 0; j < 10; ++j) obs[i * 10 + j] = a.x + j;
    rew[i] = a.y;
    done[i] = a.z > 0;
}

int main(int argc, char** argv) {
    const long n = 1L << 24;
    const int O = 10;
    float4 *h_act, *d_act; float *h_obs, *d_obs, *h_rew, *d_rew; unsigned char *h_done, *d_done;
    CK(cudaHostAlloc(&h_act, n * 16, 0)); CK(cudaHostAlloc(&h_obs, n * O * 4, 0));
    CK(cudaHostAlloc(&h_rew, n * 4, 0)); CK(cudaHostAlloc(&h_done, n, 0));
    CK(cudaMalloc(&d_act, n * 16)); CK(cudaMalloc(&d_obs, n * O * 4)); CK(cudaMalloc(&d_rew, n * 4)); CK(cudaMalloc(&d_done, n));
    memset(h_act, 0, n * 16);
    cudaStream_t st[8], s_in, s_k, s_out;
    for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> ev_in(256), ev_k(256), ev_out(256);
    for (int i = 0; i < 256; ++i) { CK(cudaEventCreate(&ev_in[i])); CK(cudaEventCreate(&ev_k[i])); CK(cudaEventCreate(&ev_out[i])); }
    cudaEvent_t e0; CK(cudaEventCreate(&e0));

    auto run = [&](int mode, long chunk, int ns, bool timeline) {
        // mode 0: round-robin streams, each chunk's H2D -> kernel -> 3 D2H on one stream (the shipped pipeline)
        // mode 1: one stream per stage (H2D stream, kernel stream, D2H stream), events between stages
        // mode 2: like 0, but only the obs D2H per chunk; reward and done leave as two whole-array copies at the end
        // mode 3: like 1, reward/done whole-array at the end
        double best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaDeviceSynchronize());
            auto t0 = std::chrono::steady_clock::now();
            CK(cudaEventRecord(e0, st[0]));
            int c = 0;
            for (long lo = 0; lo < n; lo += chunk, ++c) {
                const long m = (n - lo < chunk) ? n - lo : chunk;
                cudaStream_t a, k, o;
                if (mode == 0 || mode == 2) a = k = o = st[c % ns]; else { a = s_in; k = s_k; o = s_out; }
                CK(cudaMemcpyAsync(d_act + lo, h_act + lo, m * 16, cudaMemcpyHostToDevice, a));
                if (a != k) { CK(cudaEventRecord(ev_in[c], a)); CK(cudaStreamWaitEvent(k, ev_in[c], 0)); }
                else if (timeline) CK(cudaEventRecord(ev_in[c], a));
                touch<<<(unsigned)((m + 127) / 128), 128, 0, k>>>(d_act + lo, d_obs + lo * O, d_rew + lo, d_done + lo, m);
                if (k != o) { CK(cudaEventRecord(ev_k[c], k)); CK(cudaStreamWaitEvent(o, ev_k[c], 0)); }
                else if (timeline) CK(cudaEventRecord(ev_k[c], k));
                CK(cudaMemcpyAsync(h_obs + lo * O, d_obs + lo * O, m * O * 4, cudaMemcpyDeviceToHost, o));
                if (mode == 0 || mode == 1) {
                    CK(cudaMemcpyAsync(h_rew + lo, d_rew + lo, m * 4, cudaMemcpyDeviceToHost, o));
                    CK(cudaMemcpyAsync(h_done + lo, d_done + lo, m, cudaMemcpyDeviceToHost, o));
                }
                if (timeline) CK(cudaEventRecord(ev_out[c], o));
            }
            if (mode == 2 || mode == 3) {
                cudaStream_t o = (mode == 2) ? st[0] : s_out;
                if (mode == 2) for (int s = 1; s < ns; ++s) { CK(cudaEventRecord(ev_k[200 + s], st[s])); CK(cudaStreamWaitEvent(o, ev_k[200 + s], 0)); }
                CK(cudaMemcpyAsync(h_rew, d_rew, n * 4, cudaMemcpyDeviceToHost, o));
                CK(cudaMemcpyAsync(h_done, d_done, n, cudaMemcpyDeviceToHost, o));
            }
            auto t1 = std::chrono::steady_clock::now();
            CK(cudaDeviceSynchronize());
            auto t2 = std::chrono::steady_clock::now();
            const double ms = std::chrono::duration<double, std::milli>(t2 - t0).count();
            const double issue = std::chrono::duration<double, std::milli>(t1 - t0).count();
            if (ms < best) best = ms;
            if (timeline && rep == 5) {
                printf("  issue loop %.3f ms, total %.3f ms\n", issue, ms);
                for (int i = 0; i < c; ++i) {
                    float a, b, d;
                    CK(cudaEventElapsedTime(&a, e0, ev_in[i])); CK(cudaEventElapsedTime(&b, e0, ev_k[i])); CK(cudaEventElapsedTime(&d, e0, ev_out[i]));
                    printf("  chunk %2d: h2d done %.3f  kernel done %.3f  d2h done %.3f\n", i, a, b, d);
                }
            }
        }
        return best;
    };
    for (int mode = 0; mode < 4; ++mode)
        for (long cl : {19L, 20L, 21L})
            for (int ns : {2, 4}) {
                if ((mode == 1 || mode == 3) && ns != 2) continue;
                printf("mode %d chunk 2^%ld streams %d: %.3f ms\n", mode, cl, ns, run(mode, 1L << cl, ns, false));
            }
    printf("timeline mode 0, 2^20, 4 streams\n"); run(0, 1L << 20, 4, true);
    printf("timeline mode 1, 2^20\n"); run(1, 1L << 20, 2, true);
    return 0;
}
