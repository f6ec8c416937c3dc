// copter_policy.cuh -- the small tanh MLP policy of BASELINE.json configs[4] (O -> 64 -> 64 -> A,
// SURVEY.md 8d "config 5") as ONE kernel that reads the env's fp32 state planes in place and
// writes the action rows the step kernel consumes.  In PyTorch this policy costs 4.2 ms per
// step for 2^23 envs (three GEMMs with [N,64] intermediates through HBM plus separate tanh
// passes) against 0.21 ms for the env step itself; here the 64-wide activations never leave
// registers.
//
// One warp = 32 envs = two 16-row tiles of warp-level tensor-core MMAs
// (mma.sync.m16n8k16, bf16 inputs, fp32 accumulation).  The accumulator fragment of layer L is
// exactly the A-operand fragment of layer L+1 (n-tiles 2k, 2k+1 -> k-tile k), so between layers
// there is only bias + tanh (MUFU.TANH) + bf16 packing, all in registers.  The kernel is bound
// by the MUFU (XU) pipe -- 132 tanh per env; ncu: XU 62 %, tensor pipe 40 %, LSU 29 % busy, HBM
// idle -- so the warp-level MMA is enough here: tcgen05/TMEM would not move the XU floor.
// Weights live in shared memory as bf16, rows padded by 8 elements so that the B-fragment loads
// of a warp hit 32 distinct banks.  Persistent CTAs (weights are loaded once per CTA).
#pragma once

#include <cuda_bf16.h>

namespace copter {

constexpr int kPolH = 64;            // hidden width
constexpr int kPolIn = 16;           // observation padded to one k-tile
constexpr int kPolOut = 8;           // actions padded to one n-tile
constexpr int kPolW1Stride = kPolIn + 8, kPolW2Stride = kPolH + 8, kPolXStride = kPolIn + 8;

struct PolicySmem {
    __nv_bfloat16 w1[kPolH * kPolW1Stride];      // [64][16 (+8)]
    __nv_bfloat16 w2[kPolH * kPolW2Stride];      // [64][64 (+8)]
    __nv_bfloat16 w3[kPolOut * kPolW2Stride];    // [ 8][64 (+8)]
    float b1[kPolH], b2[kPolH], b3[kPolOut];
    __nv_bfloat16 x[4][32 * kPolXStride];        // per warp: 32 envs x 16 inputs (+8)
};

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// hidden activations: bias add in fp32, then tanh.  COPTER_POLICY_TANH_BF16X2 (A/B knob) rounds
// the pair to bf16 first and uses tanh.approx.bf16x2; it is not faster on B200 (the packed form
// costs two XU slots) and is less accurate, so the default is two fp32 MUFU.TANH.
#ifndef COPTER_POLICY_TANH_BF16X2
#define COPTER_POLICY_TANH_BF16X2 0      // measured on B200: 0.435 ms vs 0.412 ms for two fp32 tanh (2^23 envs)
#endif
__device__ __forceinline__ uint32_t tanh_pack(float lo, float hi) {
#if COPTER_POLICY_TANH_BF16X2
    uint32_t y;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(pack_bf16(lo, hi)));
    return y;
#else
    return pack_bf16(tanh_fast(lo), tanh_fast(hi));
#endif
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct PolicyArgs {
    const float* state; int64_t stride, n;           // fp32 state planes [3][stride][4]
    const float *w1, *b1, *w2, *b2, *w3, *b3;        // torch.nn.Linear layouts: W[out][in], b[out]
    float out_scale, out_offset;                     // action = out_offset + out_scale * tanh(.)
    float* action;                                   // [n][act]
};

// FIRST / OBS / ACT: observation window into the 12-state and action size of the env variant.
#ifndef COPTER_POLICY_CTAS_PER_SM
#define COPTER_POLICY_CTAS_PER_SM 3
#endif
template <int FIRST, int OBS, int ACT>
__global__ void __launch_bounds__(128, COPTER_POLICY_CTAS_PER_SM)
copter_mlp_policy_kernel(const __grid_constant__ PolicyArgs a) {
    static_assert(OBS <= kPolIn && ACT <= kPolOut && FIRST + OBS <= 12, "policy tile shapes");
    __shared__ __align__(16) PolicySmem sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;

    // weights -> shared memory (bf16), zero-padded to the tile shapes
    for (int e = threadIdx.x; e < kPolH * kPolW1Stride; e += blockDim.x) {
        const int r = e / kPolW1Stride, c = e % kPolW1Stride;
        sm.w1[e] = __float2bfloat16(c < OBS ? a.w1[r * OBS + c] : 0.0f);
    }
    for (int e = threadIdx.x; e < kPolH * kPolW2Stride; e += blockDim.x) {
        const int r = e / kPolW2Stride, c = e % kPolW2Stride;
        sm.w2[e] = __float2bfloat16(c < kPolH ? a.w2[r * kPolH + c] : 0.0f);
    }
    for (int e = threadIdx.x; e < kPolOut * kPolW2Stride; e += blockDim.x) {
        const int r = e / kPolW2Stride, c = e % kPolW2Stride;
        sm.w3[e] = __float2bfloat16((r < ACT && c < kPolH) ? a.w3[r * kPolH + c] : 0.0f);
    }
    for (int e = threadIdx.x; e < kPolH; e += blockDim.x) { sm.b1[e] = a.b1[e]; sm.b2[e] = a.b2[e]; }
    if (threadIdx.x < kPolOut) sm.b3[threadIdx.x] = threadIdx.x < ACT ? a.b3[threadIdx.x] : 0.0f;
    __syncthreads();

    __nv_bfloat16* x = sm.x[warp];
    const uint32_t* w1 = reinterpret_cast<const uint32_t*>(sm.w1);
    const uint32_t* w2 = reinterpret_cast<const uint32_t*>(sm.w2);
    const uint32_t* w3 = reinterpret_cast<const uint32_t*>(sm.w3);
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(x);

    const int64_t n_warp_tiles = (a.n + 31) / 32;
    for (int64_t wt = (int64_t)blockIdx.x * 4 + warp; wt < n_warp_tiles; wt += (int64_t)gridDim.x * 4) {
        const int64_t row0 = wt * 32, i = row0 + lane;
        // this lane's env: 12 state components -> 16 bf16 inputs (observation window, zero padded)
        float s[12];
        if (i < a.n) {
            const float4* planes = reinterpret_cast<const float4*>(a.state);
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
                const float4 v = planes[(int64_t)pl * a.stride + i];
                s[4 * pl] = v.x; s[4 * pl + 1] = v.y; s[4 * pl + 2] = v.z; s[4 * pl + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 12; ++j) s[j] = 0.0f;
        }
        uint32_t xin[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float lo = (2 * j < OBS) ? s[(FIRST + 2 * j) % 12] : 0.0f;
            const float hi = (2 * j + 1 < OBS) ? s[(FIRST + 2 * j + 1) % 12] : 0.0f;
            xin[j] = pack_bf16(lo, hi);
        }
        uint4* xrow = reinterpret_cast<uint4*>(x + lane * kPolXStride);
        xrow[0] = make_uint4(xin[0], xin[1], xin[2], xin[3]);
        xrow[1] = make_uint4(xin[4], xin[5], xin[6], xin[7]);
        __syncwarp();

#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {                       // rows 16*mt .. 16*mt+15 of the warp's 32 envs
            // ---- layer 1: [16 x 16] x [16 x 64] ------------------------------------------
            uint32_t afrag[4];
            const int r0 = (16 * mt + g) * (kPolXStride / 2), r1 = (16 * mt + g + 8) * (kPolXStride / 2);
            afrag[0] = xw[r0 + t]; afrag[1] = xw[r1 + t]; afrag[2] = xw[r0 + t + 4]; afrag[3] = xw[r1 + t + 4];
            uint32_t h[4][4];                                  // activations as A fragments of the next layer, per k-tile
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                const int wr = (8 * nt + g) * (kPolW1Stride / 2);
                mma_bf16(c, afrag, w1[wr + t], w1[wr + t + 4]);
                const float bx = sm.b1[8 * nt + 2 * t], by = sm.b1[8 * nt + 2 * t + 1];
                h[nt >> 1][(nt & 1) * 2 + 0] = tanh_pack(c[0] + bx, c[1] + by);   // rows g
                h[nt >> 1][(nt & 1) * 2 + 1] = tanh_pack(c[2] + bx, c[3] + by);   // rows g+8
            }
            // ---- layer 2: [16 x 64] x [64 x 64] ------------------------------------------
            uint32_t h2[4][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                const int wr = (8 * nt + g) * (kPolW2Stride / 2);
#pragma unroll
                for (int kt = 0; kt < 4; ++kt) mma_bf16(c, h[kt], w2[wr + 8 * kt + t], w2[wr + 8 * kt + t + 4]);
                const float bx = sm.b2[8 * nt + 2 * t], by = sm.b2[8 * nt + 2 * t + 1];
                h2[nt >> 1][(nt & 1) * 2 + 0] = tanh_pack(c[0] + bx, c[1] + by);
                h2[nt >> 1][(nt & 1) * 2 + 1] = tanh_pack(c[2] + bx, c[3] + by);
            }
            // ---- layer 3: [16 x 64] x [64 x 8] -------------------------------------------
            float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            const int wr = g * (kPolW2Stride / 2);
#pragma unroll
            for (int kt = 0; kt < 4; ++kt) mma_bf16(c, h2[kt], w3[wr + 8 * kt + t], w3[wr + 8 * kt + t + 4]);
            // columns 2t, 2t+1 of rows g and g+8
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int64_t row = row0 + 16 * mt + g + 8 * half;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int col = 2 * t + q;
                    if (row < a.n && col < ACT)
                        a.action[row * ACT + col] = a.out_offset + a.out_scale * tanh_fast(c[2 * half + q] + sm.b3[col]);
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace copter
