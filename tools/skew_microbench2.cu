// Which neighbour activity slows the recursion warp?  (profiling aid, not product)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ssnt-tts-rust_b200/csrc tools/skew_microbench2.cu -o tools/skew_mb2
#include <cstdio>
#include "fb_bf.cuh"
using namespace ssnt::lattice;

// NOISE: 0 none (other warps exit), 1 spin on an mbarrier (warp-uniform try_wait), 2 LDS/STS traffic,
// 3 MUFU, 4 LDS/STS + MUFU (prep-like), 5 global STG traffic, 6 same as 1 but with nanosleep backoff
template <int CPL, int NOISE>
__global__ void __launch_bounds__(256, 1) mb(float* out, long long* cyc, int rounds, float* gscr, float* gsink) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar;
    __shared__ volatile int stop;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int max_u = 32 * CPL, SU = max_u + 32;
    float* e = sm;
    float* s = sm + 16 * max_u;
    float* noise = sm + 32 * max_u;  // 32 KB of noise area
    for (int i = threadIdx.x; i < 16 * max_u; i += blockDim.x) { e[i] = 0.6f; s[i] = 0.4f; }
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) noise[i] = -0.5f;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); stop = 0; }
    __syncthreads();
    if (warp == 0) {
        ChainState<CPL> cs;
        cs.init(0, lane, 32 * CPL);
        float g = lane == 0 ? 0.f : 1.0f;
        int ex = 0;
        long long t0 = clock64();
        for (int k = 0; k < rounds; ++k) {
            float* base = gscr + (size_t)(blockIdx.x * 32 + (k % 32)) * 16 * SU;
            chain_round_skew<CPL, 0, false, 16>(cs, g, e, e + 16 * max_u, base, ex, lane, NoHook(), NoHook(), e + 8 * max_u,
                                                e + 24 * max_u, base + 8 * SU);
            for (int i = 0; i < CPL; ++i) cs.a[i] = fminf(cs.a[i], 1.0f);
        }
        long long t1 = clock64();
        if (lane == 0) cyc[blockIdx.x] = t1 - t0;
        out[blockIdx.x * 32 + lane] = cs.a[0] + cs.inA + cs.inB;
        __syncwarp();
        if (lane == 0) { stop = 1; mbar_arrive(smem_u32(&bar)); }
    } else {
        if (NOISE == 0) return;
        if (NOISE == 1) { mbar_wait_warp(smem_u32(&bar), 0); return; }
        if (NOISE == 6) { mbar_wait_backoff(smem_u32(&bar), 0, 200); return; }
        float acc = 0.f;
        float* mine = noise + (warp - 1) * 1024;
        int it = 0;
        while (!stop) {
            if (NOISE == 2 || NOISE == 4) {
                float4 v[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) v[r] = *reinterpret_cast<float4*>(mine + r * 128 + lane * 4);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (NOISE == 4) { v[r].x = ex2(v[r].x); v[r].y = ex2(v[r].y); v[r].z = ex2(v[r].z); v[r].w = ex2(v[r].w); }
                    v[r].x = -fabsf(v[r].x) - 0.1f;
                    *reinterpret_cast<float4*>(mine + 512 + r * 128 + lane * 4) = v[r];
                    acc += v[r].y;
                }
            } else if (NOISE == 3) {
#pragma unroll
                for (int r = 0; r < 16; ++r) acc = ex2(acc * 0.5f - 1.0f);
            } else if (NOISE == 7) {
                // large straight-line code footprint on the other sub-partitions (I-cache pressure)
#define F4(k) acc = fmaf(acc, 1.0001f + (k) * 1e-7f, 0.5f + (k) * 1e-6f); acc = fmaf(acc, 0.9999f - (k) * 1e-7f, -0.5f + (k) * 1e-6f);
#define F16(k) F4(k) F4(k + 1) F4(k + 2) F4(k + 3) F4(k + 4) F4(k + 5) F4(k + 6) F4(k + 7)
#define F128(k) F16(k) F16(k + 8) F16(k + 16) F16(k + 24) F16(k + 32) F16(k + 40) F16(k + 48) F16(k + 56)
#define F1K(k) F128(k) F128(k + 64) F128(k + 128) F128(k + 192) F128(k + 256) F128(k + 320) F128(k + 384) F128(k + 448)
                F1K(0) F1K(512) F1K(1024) F1K(1536)
            } else if (NOISE == 5) {
                float4 v = make_float4(acc, 1.f, 2.f, 3.f);
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    __stcs(reinterpret_cast<float4*>(gsink + ((size_t)(blockIdx.x * 8 + warp) * 64 + ((it * 4 + r) & 63)) * 128 + lane * 4), v);
                acc += 1.f;
            }
            ++it;
        }
        if (acc == 12345.f) out[0] = acc;
    }
}

template <int CPL, int NOISE>
void run(const char* name) {
    float* out; long long* cyc; float* gscr; float* gsink;
    cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 192 * 8);
    cudaMalloc(&gscr, (size_t)64 * 32 * 16 * (32 * CPL + 32) * 4);
    cudaMalloc(&gsink, (size_t)64 * 8 * 64 * 128 * 4);
    const int rounds = 50;
    const size_t smem = (32 * 32 * CPL) * 4 + 32768 + 1024;
    cudaFuncSetAttribute(mb<CPL, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) mb<CPL, NOISE><<<64, 256, smem>>>(out, cyc, rounds, gscr, gsink);
    cudaDeviceSynchronize();
    long long h[64]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 64; ++i) avg += h[i];
    printf("%-60s %7.1f cycles/row (%s)\n", name, avg / 64 / rounds / 16, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc); cudaFree(gscr); cudaFree(gsink);
}

int main() {
    run<4, 0>("CPL4 alpha->global, other warps exit");
    run<4, 1>("CPL4, 7 warps blocked on an mbarrier (try_wait+vote)");
    run<4, 6>("CPL4, 7 warps polling an mbarrier with nanosleep(200)");
    run<4, 2>("CPL4, 7 warps LDS.128/STS.128 traffic");
    run<4, 3>("CPL4, 7 warps MUFU.EX2");
    run<4, 4>("CPL4, 7 warps LDS+EX2+STS (prep-like)");
    run<4, 5>("CPL4, 7 warps streaming STG.128");
    run<4, 7>("CPL4, 7 warps looping over 64 KB of straight-line FFMA code");
    return 0;
}
