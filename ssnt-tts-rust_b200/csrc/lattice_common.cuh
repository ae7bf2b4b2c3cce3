// Device-side building blocks shared by the lattice kernels (log2-domain math, mbarrier / TMA
// bulk-copy wrappers, per-lane vector row access, warp reductions, deterministic loss reduction).
#pragma once
#include <cooperative_groups.h>

#include "ssnt_common.cuh"

namespace ssnt {
namespace lattice {

constexpr float kNeg = -1.0e30f;      // finite stand-in for -inf
constexpr float kNegTest = -1.0e29f;  // anything below counts as -inf
constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLn2 = 0.6931471805599453;
constexpr int kG = 8;                 // rows per pipeline stage (also the offset re-centring period)
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// log2(2^x + 2^y); operands are finite (kNeg sentinel), so n - m is never NaN.
__device__ __forceinline__ float lae2(float x, float y) {
    const float m = fmaxf(x, y);
    const float n = fminf(x, y);
    return m + lg2(1.0f + ex2(n - m));
}
__device__ __forceinline__ float to_log2(float v) { return fmaxf(v * kLog2e, kNeg); }

// ---- mbarrier / TMA bulk copy (1-D) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Wait executed by ALL 32 lanes of a warp.  The loop condition is made warp-uniform with a vote:
// if every lane spun on its own predicate, lanes could leave the loop in different iterations and
// the warp would stay diverged afterwards — every later instruction then issues once per group of
// lanes (measured: the recursion warp ran 2-12x slower).
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    while (!__all_sync(kFull, mbar_try_wait(bar, parity))) {
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// ---- per-lane row access ----------------------------------------------------------------------
// Loads this lane's CPL consecutive floats of a row (shared or global); columns >= max_u → fill.
template <int CPL>
__device__ __forceinline__ void load_cells(const float* row, int c0, int max_u, float fill, float (&v)[CPL]) {
    if constexpr (CPL >= 4) {
#pragma unroll
        for (int q = 0; q < CPL / 4; ++q) {
            if (c0 + 4 * q < max_u) {
                const float4 w = *reinterpret_cast<const float4*>(row + c0 + 4 * q);
                v[4 * q + 0] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
            } else {
                v[4 * q + 0] = fill; v[4 * q + 1] = fill; v[4 * q + 2] = fill; v[4 * q + 3] = fill;
            }
        }
    } else if constexpr (CPL == 2) {
        if (c0 < max_u) {
            const float2 w = *reinterpret_cast<const float2*>(row + c0);
            v[0] = w.x; v[1] = w.y;
        } else {
            v[0] = fill; v[1] = fill;
        }
    } else {
        v[0] = c0 < max_u ? row[c0] : fill;
    }
}
template <int CPL>
__device__ __forceinline__ void store_cells(float* row, int c0, int max_u, const float (&v)[CPL]) {
    if constexpr (CPL >= 4) {
#pragma unroll
        for (int q = 0; q < CPL / 4; ++q)
            if (c0 + 4 * q < max_u)
                *reinterpret_cast<float4*>(row + c0 + 4 * q) =
                    make_float4(v[4 * q + 0], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else if constexpr (CPL == 2) {
        if (c0 < max_u) *reinterpret_cast<float2*>(row + c0) = make_float2(v[0], v[1]);
    } else {
        if (c0 < max_u) row[c0] = v[0];
    }
}
// Streaming (evict-first) variant for the gradient tensors, which are written once.
template <int CPL>
__device__ __forceinline__ void store_cells_cs(float* row, int c0, int max_u, const float (&v)[CPL]) {
    if constexpr (CPL >= 4) {
#pragma unroll
        for (int q = 0; q < CPL / 4; ++q)
            if (c0 + 4 * q < max_u)
                __stcs(reinterpret_cast<float4*>(row + c0 + 4 * q),
                       make_float4(v[4 * q + 0], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    } else if constexpr (CPL == 2) {
        if (c0 < max_u) __stcs(reinterpret_cast<float2*>(row + c0), make_float2(v[0], v[1]));
    } else {
        if (c0 < max_u) __stcs(row + c0, v[0]);
    }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Deterministic loss = -sum_b ll[b]: the last CTA to finish adds the B values in index order.  With a connected
// loss exchange (multi-GPU, runtime.cu) the same warp then stores {loss, call number} into every rank's slot buffer
// over NVLink — lane r one posted 64-bit store to rank r, nothing else is launched for the all-reduce.
static __device__ void finish_loss(const float* ll, float* loss, int B, unsigned* counter, int lane, int nthreads,
                                   LossExchange* xchg = nullptr) {
    __shared__ unsigned s_last;
    __threadfence();
    if (lane == 0) s_last = (atomicAdd(counter, 1u) == (unsigned)(B - 1)) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (lane < 32) {  // first warp
        double acc = 0.0;
        for (int i = lane; i < B; i += 32) acc -= (double)__ldcg(ll + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
        if (lane == 0) {
            if (loss) *loss = (float)acc;
            *counter = 0u;  // hand the counter back zeroed
        }
        if (xchg && loss) loss_exchange_publish(xchg, (float)acc, lane);
    }
    (void)nthreads;
}

}  // namespace lattice
}  // namespace ssnt
