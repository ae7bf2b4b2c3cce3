// Micro-benchmark of the block-float recursion warp in isolation (profiling aid, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ssnt-tts-rust_b200/csrc tools/chain_microbench.cu -o tools/chain_mb
#include <cstdio>
#include "fb_bf.cuh"
using namespace ssnt::lattice;

// WAITMODE 0: no barrier; 1: all lanes poll an (already complete) mbarrier; 2: lane 0 polls + __syncwarp;
// 3: all lanes poll + __syncwarp
template <int CPL, int WAITMODE>
__global__ void mb(float* out, long long* cyc, int stages, float* gscr) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar;
    const int lane = threadIdx.x & 31;
    const int max_u = 32 * CPL, SU = max_u + 32, UP = max_u;
    float* e = sm;
    float* s = sm + 8 * max_u;
    for (int i = threadIdx.x; i < 8 * max_u; i += blockDim.x) { e[i] = 0.6f; s[i] = 0.4f; }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    __syncthreads();
    float v[CPL]; int ex = 0;
    for (int i = 0; i < CPL; ++i) v[i] = lane == 0 && i == 0 ? 1.0f : 0.0f;
    float g = lane == 0 ? 0.f : 1.0f;
    int own = 0, nbmag = 0, ex_dec = 0, nbex = 0;
    auto d1 = [&]() {
        float mx = v[0];
        for (int i = 1; i < CPL; ++i) mx = fmaxf(mx, v[i]);
        own = mx > 0.0f ? ex + ilogb_pos(mx) - kTarget : kNoMass;
        const float edge = v[CPL - 1];
        const int amag = edge > 0.0f ? ex + ilogb_pos(edge) : kNoMass;
        nbmag = __shfl_up_sync(kFull, amag, 1);
    };
    auto d2 = [&]() {
        int nw = max(own, nbmag - kTarget - kSlack);
        if (nw <= kNoMass / 2) nw = ex;
        ex_dec = nw;
        nbex = __shfl_up_sync(kFull, nw, 1);
    };
    unsigned par = 0;
    long long t0 = clock64();
    long long tw = 0, trw = 0;
    for (int k = 0; k < stages; ++k) {
        const long long a0 = clock64();
        if (WAITMODE == 4) { if (lane == 0) mbar_arrive(smem_u32(&bar)); __syncwarp(); mbar_wait_warp(smem_u32(&bar), par); par ^= 1u; }
        else if (WAITMODE == 5) { if (lane == 0) mbar_arrive_relaxed_n(smem_u32(&bar), 1); __syncwarp(); mbar_wait_warp(smem_u32(&bar), par); par ^= 1u; }
        else if (WAITMODE) {
            if (lane == 0) mbar_arrive(smem_u32(&bar));
            __syncwarp();
            if (WAITMODE == 1) mbar_wait(smem_u32(&bar), par);
            if (WAITMODE == 2) { if (lane == 0) mbar_wait(smem_u32(&bar), par); __syncwarp(); }
            if (WAITMODE == 3) { mbar_wait(smem_u32(&bar), par); __syncwarp(); }
            par ^= 1u;
        }
        const long long a1 = clock64();
        tw += a1 - a0;
        const int shift = ex - ex_dec;
        for (int i = 0; i < CPL; ++i) v[i] = scale_pow2(v[i], shift);
        ex = ex_dec;
        g = lane == 0 ? 0.0f : pow2i(max(-126, min(126, nbex - ex_dec)));
        float* st0 = gscr + (size_t)(blockIdx.x * 64 + (k % 64)) * 8 * SU;
        chain_stage<CPL, 0, false, true, 0>(v, g, e, s, max_u, st0, (long long)SU, ex, UP, lane, lane * CPL, max_u, d1, d2);
        trw += clock64() - a1;
    }
    long long t1 = clock64();
    if (lane == 0) { cyc[blockIdx.x] = t1 - t0; cyc[64 + blockIdx.x] = tw; cyc[128 + blockIdx.x] = trw; }
    out[blockIdx.x * 32 + lane] = v[0] + ex;
}

template <int CPL, int WAITMODE>
void run(const char* name) {
    float* out; long long* cyc; float* gscr;
    cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 192 * 8); cudaMalloc(&gscr, (size_t)64 * 64 * 8 * (32 * CPL + 32) * 4);
    const int stages = 100;
    const size_t smem = (16 * 32 * CPL) * 4 + 1024;
    cudaFuncSetAttribute(mb<CPL, WAITMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) mb<CPL, WAITMODE><<<64, 32, smem>>>(out, cyc, stages, gscr);
    cudaDeviceSynchronize();
    long long h[192]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0, aw = 0, ar = 0; for (int i = 0; i < 64; ++i) { avg += h[i]; aw += h[64 + i]; ar += h[128 + i]; }
    printf("%-64s %7.1f cycles/stage = wait %6.1f + rows %6.1f (%s)\n", name, avg / 64 / stages, aw / 64 / stages, ar / 64 / stages, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc); cudaFree(gscr);
}

int main() {
    run<4, 0>("CPL4 chain_stage, no barrier");
    run<4, 1>("CPL4 chain_stage, all lanes poll mbarrier");
    run<4, 2>("CPL4 chain_stage, lane 0 polls + __syncwarp");
    run<4, 3>("CPL4 chain_stage, all lanes poll + __syncwarp");
    run<4, 4>("CPL4 chain_stage, arrive + warp-uniform wait");
    run<4, 5>("CPL4 chain_stage, relaxed arrive + warp-uniform wait");
    run<8, 0>("CPL8 chain_stage, no barrier");
    run<8, 3>("CPL8 chain_stage, all lanes poll + __syncwarp");
    run<2, 0>("CPL2 chain_stage, no barrier");
    run<1, 0>("CPL1 chain_stage, no barrier");
    return 0;
}
