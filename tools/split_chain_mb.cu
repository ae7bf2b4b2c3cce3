// The split-role kernel's recursion round in isolation (profiling aid, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I ssnt-tts-rust_b200/csrc tools/split_chain_mb.cu -o tools/split_chain_mb
#include <cstdio>
#include "fb_split.cuh"
using namespace ssnt::lattice;

template <int CPL, int RANK>
__global__ void __launch_bounds__(32, 1) mb(float* out, long long* cyc, int rounds) {
    extern __shared__ __align__(128) float sm[];
    const int lane = threadIdx.x & 31;
    constexpr int max_u = 32 * CPL;
    constexpr int slot_floats = 3 * kG * max_u + 32;
    for (int i = threadIdx.x; i < 4 * slot_floats; i += blockDim.x) sm[i] = (i % slot_floats) < kG * max_u ? 0.6f : 0.4f;
    __syncthreads();
    ChainState<CPL> cs;
    cs.init(RANK, lane, 32 * CPL);
    float g = (RANK == 0 ? lane == 0 : lane == 31) ? 0.f : 1.0f;
    float* sp[4] = {sm, sm + slot_floats, sm + 2 * slot_floats, sm + 3 * slot_floats};
    long long t0 = clock64();
    for (int k = 0; k < rounds; ++k) {
        chain_round_split<CPL, RANK, 4>(cs, g, sp, lane, NoHook(), NoHook());
        for (int i = 0; i < CPL; ++i) cs.a[i] = fmaxf(fminf(cs.a[i], 1.0f), 1e-30f);
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 32 + lane] = cs.a[0] + cs.inA + cs.inB;
}

template <int CPL, int RANK>
void run(const char* name, int rounds) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 192 * 8);
    const size_t smem = 4 * (3 * 8 * 32 * CPL + 32) * 4;
    cudaFuncSetAttribute(mb<CPL, RANK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) mb<CPL, RANK><<<64, 32, smem>>>(out, cyc, rounds);
    cudaDeviceSynchronize();
    long long h[64]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 64; ++i) avg += h[i];
    printf("%-40s %7.1f cycles/row (%s)\n", name, avg / 64 / rounds / 32, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main(int argc, char** argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 100;
    run<4, 0>("CPL4 alpha, 4-stage rounds", rounds);
    run<4, 1>("CPL4 beta, 4-stage rounds", rounds);
    run<8, 0>("CPL8 alpha, 4-stage rounds", rounds);
    run<2, 0>("CPL2 alpha, 4-stage rounds", rounds);
    return 0;
}
