// Micro-benchmark of the SKEWED block-float recursion (profiling aid, not product): does delaying the
// consumption of the cross-lane shuffle by two rows remove the per-row shuffle stall?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ssnt-tts-rust_b200/csrc tools/skew_microbench.cu -o tools/skew_mb
#include <cstdio>
#include "fb_bf.cuh"
using namespace ssnt::lattice;

template <int CPL, int RANK, bool TO_SMEM>
__global__ void mb(float* out, long long* cyc, int rounds, float* gscr) {
    extern __shared__ __align__(128) float sm[];
    const int lane = threadIdx.x & 31;
    constexpr int max_u = 32 * CPL, SU = max_u + 32;
    float* e = sm;
    float* s = sm + 16 * max_u;
    float* stsm = sm + 32 * max_u;
    for (int i = threadIdx.x; i < 16 * max_u; i += blockDim.x) { e[i] = 0.6f; s[i] = 0.4f; }
    __syncthreads();
    ChainState<CPL> cs;
    cs.init(RANK, lane, 32 * CPL);
    float g = (RANK == 0 ? lane == 0 : lane == 31) ? 0.f : 1.0f;
    int ex = 0;
    long long t0 = clock64();
    for (int k = 0; k < rounds; ++k) {
        float* base = TO_SMEM ? stsm : gscr + (size_t)(blockIdx.x * 32 + (k % 32)) * 16 * SU;
        const bool rev = RANK == 1 && !TO_SMEM;  // beta walks the global scratch backwards
        float* st0 = rev ? base + 15 * SU : base;
        float* st1 = rev ? base + 7 * SU : base + 8 * (TO_SMEM ? max_u : SU);
        const float* eA = RANK == 0 ? e : e + 7 * max_u;
        const float* eB = RANK == 0 ? e + 8 * max_u : e + 15 * max_u;
        chain_round_skew<CPL, RANK, TO_SMEM, 16>(cs, g, eA, eA + 16 * max_u, st0, ex, lane, NoHook(), NoHook(), eB,
                                                 eB + 16 * max_u, st1);
        // crude renormalisation so the values stay finite
        for (int i = 0; i < CPL; ++i) cs.a[i] = fminf(cs.a[i], 1.0f);
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 32 + lane] = cs.a[0] + cs.inA + cs.inB;
}

template <int CPL, int RANK, bool TO_SMEM>
void run(const char* name) {
    float* out; long long* cyc; float* gscr;
    cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 192 * 8);
    cudaMalloc(&gscr, (size_t)64 * 32 * 16 * (32 * CPL + 32) * 4);
    const int rounds = 50;
    const size_t smem = (48 * 32 * CPL) * 4 + 1024;
    cudaFuncSetAttribute(mb<CPL, RANK, TO_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) mb<CPL, RANK, TO_SMEM><<<64, 32, smem>>>(out, cyc, rounds, gscr);
    cudaDeviceSynchronize();
    long long h[64]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 64; ++i) avg += h[i];
    printf("%-48s %7.1f cycles/row (%s)\n", name, avg / 64 / rounds / 16, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc); cudaFree(gscr);
}

int main() {
    run<4, 0, false>("CPL4 alpha skew, state -> global");
    run<4, 0, true>("CPL4 alpha skew, state -> shared");
    run<4, 1, false>("CPL4 beta skew, state -> global");
    run<4, 1, true>("CPL4 beta skew, state -> shared");
    run<8, 0, false>("CPL8 alpha skew, state -> global");
    run<8, 1, true>("CPL8 beta skew, state -> shared");
    run<2, 0, false>("CPL2 alpha skew, state -> global");
    run<2, 1, true>("CPL2 beta skew, state -> shared");
    return 0;
}
