// The tone-latent recursion step in isolation (profiling aid, not product): CPL=4 tokens x K=4 tones per lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I ssnt-tts-rust_b200/csrc tools/tone_chain_mb.cu -o tools/tone_chain_mb
#include <cstdio>
#include "fb_split.cuh"
using namespace ssnt::lattice;
constexpr int CPL = 4, K = 4, W = 16, RW = 512, S4 = 34, RWP = CPL * S4 * 4;

template <int MODE>  // 0: full step; 1: no state store; 2: no loads (registers only)
__global__ void __launch_bounds__(32, 1) mb(float* out, long long* cyc, int rows) {
    extern __shared__ __align__(128) float sm[];
    const int lane = threadIdx.x;
    float* es = sm;                 // 8 padded e rows | 8 padded s rows
    float* st = sm + 16 * RWP;      // 8 state rows
    for (int i = lane; i < 16 * RWP; i += 32) es[i] = i < 8 * RWP ? 0.6f : 0.1f;
    __syncwarp();
    float tone[W], v[W];
    for (int i = 0; i < W; ++i) { tone[i] = 0.25f; v[i] = lane == 0 && i < K ? 0.25f : 0.0f; }
    const float g = lane == 0 ? 0.f : 1.f;
    auto ldpp = [&](const float* row, float (&x)[W]) {
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(row + (q * S4 + lane) * 4);
            x[4 * q] = w.x; x[4 * q + 1] = w.y; x[4 * q + 2] = w.z; x[4 * q + 3] = w.w;
        }
    };
    float E[2][W], S[2][W];
    ldpp(es, E[0]); ldpp(es + 8 * RWP, S[0]);
    long long t0 = clock64();
#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
        const int q = r & 7, qn = (r + 1) & 7, cb = r & 1;
        if (MODE != 2) {
            if (cb == 0) { ldpp(es + qn * RWP, E[1]); ldpp(es + 8 * RWP + qn * RWP, S[1]); }
            else { ldpp(es + qn * RWP, E[0]); ldpp(es + 8 * RWP + qn * RWP, S[0]); }
        }
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                *reinterpret_cast<float4*>(st + q * RW + (c * 32 + lane) * 4) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
        auto step = [&](const float (&Ec)[W], const float (&Sc)[W]) {
            float X[CPL];
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                float x = 0.0f;
#pragma unroll
                for (int kk = 0; kk < K; ++kk) x = fmaf(v[i * K + kk], Sc[i * K + kk], x);
                X[i] = x;
            }
            const float in = __shfl_up_sync(kFull, X[CPL - 1], 1) * g;
#pragma unroll
            for (int i = CPL - 1; i >= 1; --i)
#pragma unroll
                for (int kk = 0; kk < K; ++kk) v[i * K + kk] = fmaf(tone[i * K + kk], X[i - 1], v[i * K + kk] * Ec[i * K + kk]);
#pragma unroll
            for (int kk = 0; kk < K; ++kk) v[kk] = fmaf(tone[kk], in, v[kk] * Ec[kk]);
        };
        if (cb == 0) step(E[0], S[0]); else step(E[1], S[1]);
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x] = t1 - t0;
    float acc = 0; for (int i = 0; i < W; ++i) acc += v[i];
    out[blockIdx.x * 32 + lane] = acc;
}
template <int MODE> void run(const char* name) {
    float* out; long long* cyc; cudaMalloc(&out, 64 * 32 * 4); cudaMalloc(&cyc, 64 * 8);
    const size_t smem = (16 * RWP + 8 * RW) * 4;
    cudaFuncSetAttribute(mb<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int rows = 1600;
    for (int i = 0; i < 2; ++i) mb<MODE><<<64, 32, smem>>>(out, cyc, rows);
    cudaDeviceSynchronize();
    long long h[64]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 64; ++i) avg += h[i];
    printf("%-40s %7.1f cycles/row (%s)\n", name, avg / 64 / rows, cudaGetErrorString(cudaGetLastError()));
}
int main() { run<0>("tone step, loads + state store"); run<1>("tone step, loads, no state store"); run<2>("tone step, registers only"); return 0; }
