// SSNT lattice forward-backward (log-likelihood + posterior-occupancy gradients) for sm_100a.
//
// Specification: SURVEY.md §8 a-FB (the reference crate has no forward-backward; the Emit/Shift
// semantics are the decoding rules of src/lib.rs:187-225 — a frame either Emits (stay on token
// u) or Shifts (u→u+1), Shift from the last token is prohibited, the last frame must Emit at
// (T-1, U-1)).  T = output frames (serial), U = input tokens (across lanes).
//
// Kernels (selected by launch_forward_backward):
//  * kind 2  fb_bf_kernel<CPL>        the hot path: warp-specialised block-floating-point recursion
//            (fb_bf.cuh) with an in-kernel log-domain re-run of utterances it cannot hold.
//  * kind 1  fb_log_warp_kernel<CPL>  cluster of two single-warp CTAs per utterance, log2 domain
//            with per-lane offsets, TMA-fed ring (fb_log_warp.cuh).  Numerically unconditional.
//  * kind 0  fb_generic_kernel        any shape/alignment (max_u % 4 != 0, U > 1024, unaligned
//            bases): one CTA per utterance, one thread per token, rows in shared memory, log2
//            domain with one integer offset per token.
#include <algorithm>
#include <cmath>
#include "fb_split.cuh"
#include "fb_tp.cuh"
#include "fb_ws.cuh"

namespace ssnt {
namespace {

using namespace lattice;

thread_local int tls_force_kind = -1;
thread_local int tls_last_kind = -1;
thread_local long long* tls_stats = nullptr;

template <int CPL>
__global__ void __launch_bounds__(32) fb_log_warp_kernel(const LogParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();  // 0 = alpha sweep, 1 = beta sweep
    const int b = blockIdx.x >> 1;
    const FbArgs& a = p.a;
    int T = a.t_len ? a.t_len[b] : a.max_t;
    int U = a.u_len ? a.u_len[b] : a.max_u;
    T = min(max(T, 0), a.max_t);
    U = min(max(U, 0), a.max_u);
    tp_pdl_trigger();           // a following time-parallel call's first kernel may be launched (it waits for this grid)
    if (p.only) tp_pdl_wait();  // launched as a dependent of the time-parallel kernels: wait for their status words
    if (p.only && p.only[b] == 0u) {
        // re-run mode: this utterance's block-float results stand
    } else if (T <= 0 || U <= 0 || U > T) {
        // No monotonic path: ll = -inf, every gradient 0.  Uniform for both CTAs of the cluster.
        const size_t slab = (size_t)a.max_t * a.max_u;
        float* g = a.logits ? (rank == 0 ? a.grad_logits : nullptr) : (rank == 0 ? a.grad_emit : a.grad_shift);
        const float zeros[CPL] = {};
        if (g) {
            g += (size_t)b * slab;
            for (int t = 0; t < a.max_t; ++t) store_cells_cs<CPL>(g + (size_t)t * a.max_u, lane * CPL, a.max_u, zeros);
        }
        if (rank == 0 && lane == 0) a.log_likelihood[b] = -INFINITY;
    } else {
        if (p.only && rank == 0 && lane == 0) atomicAdd(p.fallbacks, 1u);
        log_lattice_cta<CPL>(p, b, rank, lane, T, U, reinterpret_cast<uint64_t*>(smem_raw),
                             reinterpret_cast<float*>(smem_raw + 128), cluster);
    }
    if (rank == 0) finish_loss(a.log_likelihood, a.loss, a.batch_size, p.counter, lane, 32, a.xchg);
}

// MINB = 2: throughput variant, two CTAs (of different utterances) per SM — 128 registers per thread and
// half the shared-memory ring each; their phases interleave, which fills the helpers' idle time.
template <int CPL, int MINB = 1>
__global__ void __launch_bounds__(kBfThreads, MINB) fb_bf_kernel(const BfParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();  // 0 = alpha sweep, 1 = beta sweep
    const int b = blockIdx.x >> 1;
    const FbArgs& a = p.a;
    int T = a.t_len ? a.t_len[b] : a.max_t;
    int U = a.u_len ? a.u_len[b] : a.max_u;
    T = min(max(T, 0), a.max_t);
    U = min(max(U, 0), a.max_u);
    const size_t slab = (size_t)a.max_t * a.max_u;
    if (T <= 0 || U <= 0 || U > T) {
        // No monotonic path: ll = -inf, every gradient 0.  Uniform for both CTAs of the cluster.
        float4* g = reinterpret_cast<float4*>((rank == 0 ? a.grad_emit : a.grad_shift) + (size_t)b * slab);
        for (size_t i = tid; i < slab / 4; i += kBfThreads) __stcs(g + i, make_float4(0.f, 0.f, 0.f, 0.f));
        if (rank == 0 && tid == 0) a.log_likelihood[b] = -INFINITY;
    } else {
        bf_lattice_cta<CPL>(p, b, rank, T, U, smem_raw, cluster);
        // Padded frames t >= T (rank 0 clears grad_emit, rank 1 grad_shift).
        {
            float4* g = reinterpret_cast<float4*>((rank == 0 ? a.grad_emit : a.grad_shift) + (size_t)b * slab +
                                                  (size_t)T * a.max_u);
            const size_t n4 = (size_t)(a.max_t - T) * a.max_u / 4;
            for (size_t i = tid; i < n4; i += kBfThreads) __stcs(g + i, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        // Did either CTA flag the utterance?  If so the same cluster redoes it in the log domain.
        cluster.sync();
        const unsigned st = *reinterpret_cast<volatile unsigned*>(p.status + b);
        if (st && !(p.debug_skip & 16)) {
            if (rank == 0 && tid == 0) atomicAdd(p.fallbacks, 1u);
            if (warp == 0) {
                LogParams lp;
                lp.a = a;
                lp.scratch = p.scratch;
                lp.SU = p.SU;
                lp.NS = p.NS < 8 ? p.NS : 8;
                lp.counter = p.counter;
                log_lattice_cta<CPL>(lp, b, rank, lane, T, U, reinterpret_cast<uint64_t*>(smem_raw + 576),
                                     reinterpret_cast<float*>(smem_raw + kBfHeaderBytes), cluster);
            } else {
                cluster.sync();
            }
        }
    }
    if (rank == 0) finish_loss(a.log_likelihood, a.loss, a.batch_size, p.counter, tid, kBfThreads, a.xchg);
}


template <int CPL>
__global__ void __launch_bounds__(kSplitThreads, 1) fb_split_kernel(const SplitParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();  // 0/1 recursion CTAs (alpha/beta), 2/3 their helper CTAs
    const int b = blockIdx.x >> 2;
    const FbArgs& a = p.a;
    auto gtime = []() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return (long long)t; };
    long long* tl = p.stats ? p.stats + (size_t)gridDim.x * 16 * 16 + (size_t)blockIdx.x * 4 : nullptr;  // launch timeline (ns)
    if (tl && tid == 0) tl[0] = gtime();
    int T = a.t_len ? a.t_len[b] : a.max_t;
    int U = a.u_len ? a.u_len[b] : a.max_u;
    T = min(max(T, 0), a.max_t);
    U = min(max(U, 0), a.max_u);
    const size_t slab = (size_t)a.max_t * a.max_u;
    if (T <= 0 || U <= 0 || U > T) {
        // No monotonic path: ll = -inf, every gradient 0.  Uniform for the four CTAs of the cluster.
        if (rank >= 2) {
            float4* g = reinterpret_cast<float4*>((rank == 2 ? a.grad_emit : a.grad_shift) + (size_t)b * slab);
            for (size_t i = tid; i < slab / 4; i += kSplitThreads) __stcs(g + i, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        if (rank == 0 && tid == 0) a.log_likelihood[b] = -INFINITY;
    } else {
        // Touch what the first stages will read before the cluster barrier: the address translations of a
        // fresh launch (all recursion CTAs miss the TLB at once, ~3 us) then overlap the barrier.
        if (rank < 2 && tid < 64) {
            const int rows = min(2 * kG, T);
            const int r0 = rank == 0 ? 0 : T - rows;
            const float* base = (tid < 32 ? a.log_emit : a.log_shift) + (size_t)b * slab + (size_t)r0 * a.max_u;
            const int nlines = rows * a.max_u / 32;
            for (int l = tid & 31; l < nlines; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + l * 32));
        }
        for (int i = tid; i < (kSplitHeaderBytes - 128) / 4; i += kSplitThreads) reinterpret_cast<int*>(smem_raw + 128)[i] = 0;
        if (tid == 0) {
            if (rank == 0) p.status[b] = p.force_fallback ? (unsigned)kBfForced : 0u;
        }
        __syncthreads();
        cluster.sync();  // flags are zero everywhere before anybody writes into a neighbour's shared memory
        if (tl && tid == 0) tl[1] = gtime();
        if (rank < 2) split_chain_cta<CPL>(p, b, rank, T, U, smem_raw);
        else split_helper_cta<CPL>(p, b, rank, T, U, smem_raw);
        // Padded frames t >= T (helper 0 clears grad_emit, helper 1 grad_shift).
        if (rank >= 2) {
            float4* g = reinterpret_cast<float4*>((rank == 2 ? a.grad_emit : a.grad_shift) + (size_t)b * slab +
                                                  (size_t)T * a.max_u);
            const size_t n4 = (size_t)(a.max_t - T) * a.max_u / 4;
            for (size_t i = tid; i < n4; i += kSplitThreads) __stcs(g + i, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        // Did a helper flag the utterance?  If so the recursion CTAs redo it in the log domain.
        if (tl && tid == 0) tl[2] = gtime();
        __threadfence();
        cluster.sync();
        const unsigned st = *reinterpret_cast<volatile unsigned*>(p.status + b);
        if (st) {
            if (rank == 0 && tid == 0) atomicAdd(p.fallbacks, 1u);
            if (rank < 2 && warp == 0) {
                LogParams lp;
                lp.a = a;
                // this utterance's own A region doubles as the re-run's scratch rows
                float* mine = p.A + (size_t)b * 2 * p.nstp * kG * p.SU;
                lp.scratch = mine - (size_t)b * (a.max_t + 1) * p.SU;
                lp.SU = p.SU;
                {
                    const int log_stage_bytes = kG * (2 * a.max_u + p.SU) * (int)sizeof(float);
                    const int fit = p.ring_bytes / log_stage_bytes;
                    lp.NS = fit < 8 ? fit : 8;  // >= 2: a split slot is 3 rows wide per row, a log stage 3.1
                }
                lp.counter = p.counter;
                log_lattice_cta<CPL>(lp, b, rank, lane, T, U, reinterpret_cast<uint64_t*>(smem_raw + 512),
                                     reinterpret_cast<float*>(smem_raw + kSplitHeaderBytes), cluster);
            } else {
                cluster.sync();
            }
        }
    }
    if (rank == 0) finish_loss(a.log_likelihood, a.loss, a.batch_size, p.counter, tid, kSplitThreads, a.xchg);
    if (tl && tid == 0) tl[3] = gtime();
}

// ===============================================================================================
// Generic path: one CTA per utterance, thread per token (strided when U > blockDim).
// Per token: value in (-0.5, 0.5] + integer offset, i.e. an explicit integer/fraction split of the
// log2 quantity, so fp32 rounding never sees a large magnitude.
// ===============================================================================================
struct GenericParams {
    FbArgs a;
    float* sval;  // [B][max_t][max_u] alpha~ fractions
    float* soff;  // [B][max_t][max_u] alpha~ integer offsets
    unsigned* counter;
};

__global__ void fb_generic_kernel(const GenericParams p) {
    extern __shared__ float sm[];
    const FbArgs& a = p.a;
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int max_t = a.max_t, max_u = a.max_u;
    int T = a.t_len ? a.t_len[b] : max_t;
    int U = a.u_len ? a.u_len[b] : max_u;
    T = min(max(T, 0), max_t);
    U = min(max(U, 0), max_u);
    const size_t slab = (size_t)max_t * max_u;
    const float* le = a.log_emit + (size_t)b * slab;
    const float* ls = a.log_shift + (size_t)b * slab;
    float* ge = a.grad_emit + (size_t)b * slab;
    float* gs = a.grad_shift + (size_t)b * slab;
    float* sval = p.sval + (size_t)b * slab;
    float* soff = p.soff + (size_t)b * slab;

    if (T <= 0 || U <= 0 || U > T) {
        for (size_t i = tid; i < slab; i += nt) { ge[i] = 0.0f; gs[i] = 0.0f; }
        if (tid == 0) a.log_likelihood[b] = -INFINITY;
        finish_loss(a.log_likelihood, a.loss, a.batch_size, p.counter, tid, nt, a.xchg);
        return;
    }
    // four rows of (max_u + 2) floats with a guard cell on each side: index u+1 <-> token u
    const int RW = max_u + 2;
    float* cv = sm;            // current values
    float* co = sm + RW;       // current offsets
    float* nv = sm + 2 * RW;   // next values
    float* no = sm + 3 * RW;   // next offsets
    for (int i = tid; i < RW; i += nt) { cv[i] = kNeg; nv[i] = kNeg; co[i] = 0.0f; no[i] = 0.0f; }
    __syncthreads();
    if (tid == 0) cv[1] = 0.0f;  // alpha(0,0) = 0
    __syncthreads();
    // ---- forward: store alpha(t) as (fraction, offset), then advance ----
    for (int t = 0; t < T; ++t) {
        for (int u = tid; u < U; u += nt) {
            sval[(size_t)t * max_u + u] = cv[u + 1];
            soff[(size_t)t * max_u + u] = co[u + 1];
        }
        if (t < T - 1) {
            for (int u = tid; u < U; u += nt) {
                const float e = to_log2(le[(size_t)t * max_u + u]);
                const float stay = cv[u + 1] + e;
                const float sh = u > 0 ? (cv[u] + (co[u] - co[u + 1])) + to_log2(ls[(size_t)t * max_u + u - 1]) : kNeg;
                const float r = lae2(stay, sh);
                const float c = r > kNegTest ? rintf(r) : 0.0f;
                nv[u + 1] = fmaxf(r - c, kNeg);
                no[u + 1] = co[u + 1] + c;
            }
            __syncthreads();
            float* tmp = cv; cv = nv; nv = tmp;
            tmp = co; co = no; no = tmp;
        }
    }
    __syncthreads();
    // LL2 = ref + llt with ref the (integer) offset of token U-1 at the last frame
    const float ref = co[U];
    const float llt = cv[U] + to_log2(le[(size_t)(T - 1) * max_u + U - 1]);
    const bool dead = !(llt > kNegTest);
    if (tid == 0) {
        const double ll2 = (double)llt + (double)ref;
        a.log_likelihood[b] = dead ? -INFINITY : (float)(ll2 * kLn2);
    }
    __syncthreads();
    // ---- backward with fused gradients: cv/co := beta(t+1, .), virtual terminal row at t = T ----
    for (int i = tid; i < RW; i += nt) { cv[i] = kNeg; nv[i] = kNeg; co[i] = 0.0f; no[i] = 0.0f; }
    __syncthreads();
    if (tid == 0) cv[U] = 0.0f;  // beta(T, U-1) = 0
    __syncthreads();
    for (int t = T - 1; t >= 0; --t) {
        for (int u = tid; u < max_u; u += nt) {
            float g1 = 0.0f, g2 = 0.0f;
            if (u < U) {
                const float e = to_log2(le[(size_t)t * max_u + u]);
                const float s = (t == T - 1 || u == U - 1) ? kNeg : to_log2(ls[(size_t)t * max_u + u]);
                const float av = sval[(size_t)t * max_u + u];
                const float ao = soff[(size_t)t * max_u + u];
                const float x = e + cv[u + 1];
                const float y = s + (cv[u + 2] + (co[u + 2] - co[u + 1]));
                const float kt = ((ao + co[u + 1]) - ref) - llt;
                if (!dead) {
                    g1 = ex2((av + x) + kt);
                    g2 = ex2((av + y) + kt);
                }
                const float r = lae2(x, y);
                const float c = r > kNegTest ? rintf(r) : 0.0f;
                nv[u + 1] = fmaxf(r - c, kNeg);
                no[u + 1] = co[u + 1] + c;
            }
            ge[(size_t)t * max_u + u] = g1;
            gs[(size_t)t * max_u + u] = g2;
        }
        __syncthreads();
        float* tmp = cv; cv = nv; nv = tmp;
        tmp = co; co = no; no = tmp;
    }
    for (size_t i = (size_t)T * max_u + tid; i < slab; i += nt) { ge[i] = 0.0f; gs[i] = 0.0f; }
    finish_loss(a.log_likelihood, a.loss, a.batch_size, p.counter, tid, nt, a.xchg);
}

template <int CPL>
void launch_warp(const LogParams& p, size_t smem, cudaStream_t stream, bool dependent = false) {
    static size_t configured_[64] = {};  // per instantiation and device: largest opt-in requested so far
    size_t& configured = configured_[device_ordinal()];
    if (configured == 0) configured = 48 * 1024;
    if (smem > configured) {
        SSNT_CUDA(cudaFuncSetAttribute(fb_log_warp_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)p.a.batch_size * 2u);
    cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // re-run after the time-parallel kernels
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = dependent ? 2 : 1;
    SSNT_CUDA(cudaLaunchKernelEx(&cfg, fb_log_warp_kernel<CPL>, p));
}

template <int CPL, int MINB = 1>
void launch_bf(const BfParams& p, size_t smem, cudaStream_t stream) {
    static size_t configured_[64] = {};  // per device
    size_t& configured = configured_[device_ordinal()];
    if (configured == 0) configured = 48 * 1024;
    if (smem > configured) {
        SSNT_CUDA(cudaFuncSetAttribute(fb_bf_kernel<CPL, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)p.a.batch_size * 2u);
    cfg.blockDim = dim3(kBfThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SSNT_CUDA(cudaLaunchKernelEx(&cfg, fb_bf_kernel<CPL, MINB>, p));
}

template <int CPL>
void launch_split(const SplitParams& p, size_t smem, cudaStream_t stream) {
    static size_t configured_[64] = {};  // per device
    size_t& configured = configured_[device_ordinal()];
    if (configured == 0) configured = 48 * 1024;
    if (smem > configured) {
        SSNT_CUDA(cudaFuncSetAttribute(fb_split_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)p.a.batch_size * 4u);
    cfg.blockDim = dim3(kSplitThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SSNT_CUDA(cudaLaunchKernelEx(&cfg, fb_split_kernel<CPL>, p));
}

inline int round_up4(int x) { return (x + 3) & ~3; }

// Time-parallel path (kind 6): chunk operators, boundary vectors, chunk interiors; one warp per CTA throughout.
template <int CPL, int L = kTpL, int G = (32 * CPL <= 128 ? 32 : 16)>
void launch_tp(const TpParams& p, cudaStream_t stream) {
    const FbArgs& a = p.a;
    const size_t chunk_smem = 128 + (size_t)2 * L * a.max_u * sizeof(float);
    constexpr int NT = 32 * CPL;
    constexpr int NS0 = NT <= 128 ? 16 : 8;   // 139 KB of operators in flight per (utterance, direction)
    // tuning aid: half the ring (leaves room for more early-launched fill CTAs next to a combine CTA)
    static const int ring_sel = [] { const char* e = std::getenv("SSNT_TP_RING"); return e ? std::atoi(e) : 0; }();  // 1 half, 2 one and a half
    const int NS = ring_sel == 1 ? NS0 / 2 : (ring_sel == 2 ? NS0 * 3 / 2 : NS0);
    // Exponent granularity of the boundary vectors (G): a warp's 32 tokens for short sweeps of narrow lattices, half a
    // warp otherwise — long sweeps lose likelihood with 32-token groups (T = 2000 at U = 256: ~1e-4; T = 1000 at U = 128:
    // 5 of 1184 random utterances missed the 3e-5 agreement of the two sweeps and were re-run), see the call sites.
    const size_t ring_smem = 576 + ((size_t)2 * (2 * L + NT) + 4) * sizeof(float) + (size_t)NS * (L + 1) * NT * sizeof(float);
    static size_t configured_[64] = {};  // per device
    size_t& configured = configured_[device_ordinal()];
    if (configured == 0) configured = 48 * 1024;
    if (ring_smem > configured) {
        SSNT_CUDA(cudaFuncSetAttribute(tp_combine_kernel<NT, L, NS0, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
        SSNT_CUDA(cudaFuncSetAttribute(tp_combine_kernel<NT, L, NS0 / 2, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
        SSNT_CUDA(cudaFuncSetAttribute(tp_combine_kernel<NT, L, NS0 * 3 / 2, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
        configured = ring_smem;
    }
    const unsigned tasks = (unsigned)a.batch_size * (unsigned)p.C;
#ifdef SSNT_BF_DEBUG_VARIANTS  // stop after the first n kernels (profiling aid; results are incomplete)
    static const int stages = [] { const char* e = std::getenv("SSNT_TP_DEBUG_STAGES"); return e ? std::atoi(e) : 3; }();
#else
    constexpr int stages = 3;
#endif
    TpParams pg = p;
    pg.G = G;
    const TpParams& p2 = pg;
    const bool lg = a.logits != nullptr;  // raw-logit mode: one input tensor, one gradient tensor
    // every lane's cells exist (max_u = 32 * CPL: 64, 128, 256): chunk kernels without bounds checks on their row accesses
    static const bool no_full = [] { const char* e = std::getenv("SSNT_TP_NO_FULL"); return e && std::atoi(e) != 0; }();  // A/B aid
    const bool full = a.max_u == 32 * CPL && p.UP == 32 * CPL && !no_full;
    // dependent launches: each kernel's CTAs start while its predecessor drains and block in griddepcontrol.wait
    cudaLaunchAttribute pdl[1];
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.stream = stream;
    cfg.attrs = pdl;
    static const int pdl_mask = [] { const char* e = std::getenv("SSNT_TP_PDL"); return e ? std::atoi(e) : 7; }();  // tuning aid: 1 combine, 2 fill (its row conversions run before it waits: 0.9 us at cfg2, 3.5 us at B=148), 4 re-run, 8 build as a dependent of the previous call's re-run kernel (measured 1 us slower at cfg2: off)
    cfg.numAttrs = (pdl_mask & 8) ? 1 : 0;
    if (stages >= 1) {
        cfg.gridDim = dim3(tasks);
        cfg.blockDim = dim3(32);
        cfg.dynamicSmemBytes = chunk_smem;
        if (lg && full) SSNT_CUDA(cudaLaunchKernelEx(&cfg, tp_build_kernel<CPL, L, true, true>, pg));
        else if (lg) SSNT_CUDA(cudaLaunchKernelEx(&cfg, tp_build_kernel<CPL, L, true>, pg));
        else if (full) SSNT_CUDA(cudaLaunchKernelEx(&cfg, tp_build_kernel<CPL, L, false, true>, pg));
        else SSNT_CUDA(cudaLaunchKernelEx(&cfg, tp_build_kernel<CPL, L, false>, pg));
    }
    cfg.numAttrs = (pdl_mask & 1) ? 1 : 0;
    if (stages >= 2) {
        cfg.gridDim = dim3((unsigned)a.batch_size * 2u);
        cfg.blockDim = dim3(NT + 32);
        cfg.dynamicSmemBytes = ring_smem;
        if (ring_sel == 1) SSNT_CUDA(cudaLaunchKernelEx(&cfg, tp_combine_kernel<NT, L, NS0 / 2, G>, pg));
        else if (ring_sel == 2) SSNT_CUDA(cudaLaunchKernelEx(&cfg, tp_combine_kernel<NT, L, NS0 * 3 / 2, G>, pg));
        else SSNT_CUDA(cudaLaunchKernelEx(&cfg, tp_combine_kernel<NT, L, NS0, G>, pg));
    }
    cfg.numAttrs = (pdl_mask & 2) ? 1 : 0;
    if (stages >= 3) {
        cfg.gridDim = dim3(tasks);
        cfg.blockDim = dim3(32);
        cfg.dynamicSmemBytes = chunk_smem;
        if (lg && full) SSNT_CUDA(cudaLaunchKernelEx(&cfg, tp_fill_kernel<CPL, L, true, true>, pg));
        else if (lg) SSNT_CUDA(cudaLaunchKernelEx(&cfg, tp_fill_kernel<CPL, L, true>, pg));
        else if (full) SSNT_CUDA(cudaLaunchKernelEx(&cfg, tp_fill_kernel<CPL, L, false, true>, pg));
        else SSNT_CUDA(cudaLaunchKernelEx(&cfg, tp_fill_kernel<CPL, L, false>, pg));
    }
}

// Resident warps per SM for the warp-serial kernels.  All utterances of a wave finish together, so what counts is how
// evenly the batch divides into waves: efficiency(r) = per_sm / (ceil(per_sm / r) * r).  The count that fits is kept
// unless it wastes more than 30 % of its slots and a smaller one (not below half) is clearly more even.
inline int ws_resident_cap(size_t batch, int sms, int fit) {
    if (fit < 4) return fit;
    const double per_sm = (double)batch / (double)sms;
    if (per_sm <= (double)fit) return fit;  // one wave anyway
    auto eff = [&](int r) { return per_sm / (std::ceil(per_sm / r) * r); };
    const double e0 = eff(fit);
    if (e0 >= 0.7) return fit;  // measured: switching pays at 0.53 and 0.65 (B = 2048, 2500: +20 %), costs 3-10 % at 0.71-0.78
    int best = fit;
    double best_e = e0 + 0.1;
    for (int r = fit - 1; r >= (fit + 1) / 2 && r >= 4; --r)
        if (eff(r) > best_e) { best_e = eff(r); best = r; }
    return best;
}

// Warp-serial throughput path (kind 8): alpha checkpoints forward, chunked beta + gradients backward.
template <int CPL, int L>
void launch_ws(const WsParams& p, cudaStream_t stream) {
    const FbArgs& a = p.a;
    size_t fwd_smem = 128 + (size_t)p.NS * 2 * p.R * a.max_u * sizeof(float);
    size_t bwd_smem = 128 + (size_t)kWsBwdStages * 2 * L * a.max_u * sizeof(float);
    // Even waves: all utterances of a wave finish together, so B = 2048 at 13 resident warps per SM is one full wave
    // plus a sliver (124 utterances on an otherwise idle GPU).  Asking for more shared memory than needed caps the
    // resident warps per SM at a count that divides the batch into equal waves (ws_resident_cap).
    {
        const size_t sm_bytes = 228 * 1024, per_cta = 1024;  // shared memory per SM, reserved per resident CTA
        static const int env_cap = [] { const char* e = std::getenv("SSNT_WS_RESIDENT"); return e ? std::atoi(e) : -1; }();  // tuning aid: 0 off, n forces
        const int fit_f = (int)(sm_bytes / (fwd_smem + per_cta)), fit_b = (int)(sm_bytes / (bwd_smem + per_cta));
        const int cap_f = env_cap > 0 ? env_cap : (env_cap == 0 ? fit_f : ws_resident_cap((size_t)a.batch_size, sm_count(), fit_f));
        const int cap_b = env_cap > 0 ? env_cap : (env_cap == 0 ? fit_b : ws_resident_cap((size_t)a.batch_size, sm_count(), fit_b));
        if (cap_f >= 3 && cap_f < fit_f) fwd_smem = std::min<size_t>(sm_bytes / cap_f - per_cta - 256, 100 * 1024);
        if (cap_b >= 3 && cap_b < fit_b) bwd_smem = std::min<size_t>(sm_bytes / cap_b - per_cta - 256, 100 * 1024);
    }
    static bool configured_[64] = {};  // per device
    bool& configured = configured_[device_ordinal()];
    if (!configured) {
        SSNT_CUDA(cudaFuncSetAttribute(ws_forward_kernel<CPL, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        SSNT_CUDA(cudaFuncSetAttribute(ws_backward_kernel<CPL, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        SSNT_CUDA(cudaFuncSetAttribute(ws_backward_kernel<CPL, L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured = true;
    }
    ws_forward_kernel<CPL, L><<<(unsigned)a.batch_size, 32, fwd_smem, stream>>>(p);
    SSNT_CUDA(cudaGetLastError());
    cudaLaunchAttribute pdl[1];
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.stream = stream;
    cfg.attrs = pdl;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)a.batch_size);
    cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = bwd_smem;
    // every lane's cells exist (max_u = 32 * CPL): no bounds checks on the row accesses, compile-time strides
    static const bool no_full = [] { const char* e = std::getenv("SSNT_TP_NO_FULL"); return e && std::atoi(e) != 0; }();  // A/B aid
    if (a.max_u == 32 * CPL && p.UP == 32 * CPL && !no_full) SSNT_CUDA(cudaLaunchKernelEx(&cfg, ws_backward_kernel<CPL, L, true>, p));
    else SSNT_CUDA(cudaLaunchKernelEx(&cfg, ws_backward_kernel<CPL, L>, p));
}

}  // namespace

// Time-parallel path (kind 6): region 0 holds the chunk operators Q [B][C][L+1][UP] and is re-used as the scratch
// rows of the log-domain re-run; then the boundary vectors A, Bv [B][C+1][UP+32], the two likelihood estimates and
// the status words.
struct TpLayout {
    int CPL, UP, C, SU;
    size_t region0, vec, total;
};
constexpr int kTpLShort = 8;  // chunk length for wide lattices at medium batch sizes (half the operator work)
static bool tp_layout(int B, int max_t, int max_u, TpLayout& l, int L = kTpL) {
    if (max_u % 4 != 0 || max_u > 256 || max_u <= 0 || max_t <= 0 || B <= 0) return false;
    l.CPL = max_u <= 64 ? 2 : (max_u <= 128 ? 4 : 8);
    l.UP = 32 * l.CPL;
    l.C = (max_t + L - 1) / L;
    l.SU = max_u + 32;
    const size_t q = (size_t)B * l.C * (L + 1) * l.UP * sizeof(float);
    const size_t scr = (size_t)B * (max_t + 1) * l.SU * sizeof(float);
    l.region0 = ((q > scr ? q : scr) + 255) & ~(size_t)255;
    l.vec = (((size_t)B * (l.C + 1) * (l.UP + 32) * sizeof(float)) + 255) & ~(size_t)255;
    l.total = l.region0 + 2 * l.vec + (((size_t)B * 4 * sizeof(float) + 255) & ~(size_t)255) +
              (((size_t)B * sizeof(unsigned) + 255) & ~(size_t)255);
    return true;
}

// Warp-serial path (kind 8): region 0 = the scratch rows of the log-domain re-run; then the alpha checkpoints
// A [B][C+1][UP+32], the forward likelihoods and the status words.
struct WsLayout {
    int CPL, L, UP, C, SU, R;
    size_t region0, vec, total;
};
static bool ws_layout(int B, int max_t, int max_u, WsLayout& l) {
    if (max_u % 4 != 0 || max_u > 256 || max_u <= 0 || max_t <= 0 || B <= 0) return false;
    l.CPL = max_u <= 64 ? 2 : (max_u <= 128 ? 4 : 8);
    l.L = l.CPL == 8 ? 8 : 16;
    l.UP = 32 * l.CPL;
    l.C = (max_t + l.L - 1) / l.L;
    l.SU = max_u + 32;
    l.R = (kWsStageBytes / (8 * max_u)) & ~3;  // whole groups of four rows
    if (l.R < 4) l.R = 4;
    l.region0 = (((size_t)B * (max_t + 1) * l.SU * sizeof(float)) + 255) & ~(size_t)255;
    l.vec = (((size_t)B * (l.C + 1) * (l.UP + 32) * sizeof(float)) + 255) & ~(size_t)255;
    l.total = l.region0 + l.vec + (((size_t)B * 2 * sizeof(float) + 255) & ~(size_t)255) +
              (((size_t)B * sizeof(unsigned) + 255) & ~(size_t)255);
    return true;
}

// Split kernel (kind 4): state rows A [B][2][nstp*8][SU] (both sweeps, sweep order) and status [B].
static size_t split_workspace_bytes(int B, int max_t, int max_u) {
    if (max_u != 64 && max_u != 128 && max_u != 256) return 0;
    if (B > 64) return 0;  // one wave of 4-CTA clusters only: never taken (nor forced in a test) for larger batches
    const size_t SU = (size_t)max_u + 32;
    const size_t nstp = ((size_t)max_t + kG - 1) / kG;
    const size_t A = (size_t)B * 2 * nstp * kG * SU * sizeof(float);
    return A + (((size_t)B * sizeof(unsigned) + 255) & ~(size_t)255) + 512;
}

// Workspace: scratch rows (+1 virtual row, stride max_u+4 rounded to 4) and per-row offsets for
// the generic kernel; sized for whichever kernel is picked.
size_t fb_workspace_bytes(int B, int max_t, int max_u) {
    if (B <= 0 || max_t <= 0 || max_u <= 0) return 256;
    const size_t SU = (size_t)round_up4(max_u) + 32;
    size_t warp_bytes = (size_t)B * (max_t + 1) * SU * sizeof(float) + (((size_t)B * sizeof(unsigned) + 255) & ~(size_t)255);
    size_t gen_bytes = (size_t)2 * B * max_t * max_u * sizeof(float);
    size_t n = warp_bytes > gen_bytes ? warp_bytes : gen_bytes;
    const size_t split_bytes = split_workspace_bytes(B, max_t, max_u);
    n = n > split_bytes ? n : split_bytes;
    TpLayout tl;
    if (tp_layout(B, max_t, max_u, tl)) n = n > tl.total ? n : tl.total;
    if (tp_layout(B, max_t, max_u, tl, kTpLShort)) n = n > tl.total ? n : tl.total;
    WsLayout wl;
    if (ws_layout(B, max_t, max_u, wl)) n = n > wl.total ? n : wl.total;
    return (n + 255) & ~(size_t)255;
}

int fb_last_kernel_kind() { return tls_last_kind; }
void fb_set_stats_buffer(long long* dev) { tls_stats = dev; }
long long* fb_get_stats_buffer() { return tls_stats; }
void fb_force_kernel_kind(int kind) { tls_force_kind = kind; }

// ---- raw logits on shapes the fused kernels do not take ---------------------------------------------------------
namespace {
__global__ void logits_to_logprobs_kernel(const float* __restrict__ z, float* __restrict__ le, float* __restrict__ ls, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = z[i], sp = log1pf(expf(-fabsf(v)));
        le[i] = fminf(v, 0.0f) - sp;   // log sigmoid(z)
        ls[i] = fminf(-v, 0.0f) - sp;  // log sigmoid(-z)
    }
}
__global__ void chain_logit_gradient_kernel(const float* __restrict__ z, const float* __restrict__ ge, const float* __restrict__ gs,
                                            float* __restrict__ gz, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = z[i], t = expf(-fabsf(v)), r = 1.0f / (1.0f + t);
        const float p = v >= 0.0f ? r : t * r, q = v >= 0.0f ? t * r : r;  // sigmoid(z), sigmoid(-z)
        gz[i] = ge[i] * q - gs[i] * p;
    }
}
}  // namespace

static size_t logits_unfused_bytes(int B, int max_t, int max_u) {
    return (((size_t)4 * B * max_t * max_u * sizeof(float)) + 255) & ~(size_t)255;
}

size_t fb_logits_workspace_bytes(int B, int max_t, int max_u) {
    if (B <= 0 || max_t <= 0 || max_u <= 0) return 256;
    size_t n = fb_workspace_bytes(B, max_t, max_u);
    if (max_u % 4 != 0 || max_u > 1024) n += logits_unfused_bytes(B, max_t, max_u);
    return n;
}

void launch_forward_backward(const FbArgs& a_in, cudaStream_t stream);
static void launch_logits_unfused(const FbArgs& a, void* ws, cudaStream_t stream) {
    const size_t n = (size_t)a.batch_size * a.max_t * a.max_u;
    const size_t base = fb_workspace_bytes(a.batch_size, a.max_t, a.max_u);
    float* tail;
    if (a.workspace && a.workspace_bytes >= base + logits_unfused_bytes(a.batch_size, a.max_t, a.max_u)) {
        tail = (float*)((char*)ws + base);
    } else {  // e.g. an aligned shape with unaligned buffers: the query did not reserve the tail
        tail = (float*)device_scratch(2, 4 * n * sizeof(float) + 16);
    }
    float *le = tail, *ls = tail + n, *ge = tail + 2 * n, *gs = tail + 3 * n;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 16);
    logits_to_logprobs_kernel<<<blocks, 256, 0, stream>>>(a.logits, le, ls, n);
    SSNT_CUDA(cudaGetLastError());
    FbArgs b = a;
    b.logits = nullptr; b.grad_logits = nullptr;
    b.log_emit = le; b.log_shift = ls; b.grad_emit = ge; b.grad_shift = gs;
    b.workspace = ws; b.workspace_bytes = base;
    launch_forward_backward(b, stream);
    chain_logit_gradient_kernel<<<blocks, 256, 0, stream>>>(a.logits, ge, gs, a.grad_logits, n);
    SSNT_CUDA(cudaGetLastError());
}

void launch_forward_backward(const FbArgs& a_in, cudaStream_t stream) {
    FbArgs a = a_in;
    a.xchg = loss_exchange_device();  // multi-GPU: the loss reduction also stores into the peers' slot buffers
    if (a.batch_size <= 0) {
        if (a.loss) SSNT_CUDA(cudaMemsetAsync(a.loss, 0, sizeof(float), stream));
        return;
    }
    SSNT_ASSERT(a.max_t >= 0 && a.max_u >= 0, "negative lattice size");
    if (a.max_t == 0 || a.max_u == 0) {
        // Empty lattices: ll = -inf for every utterance, loss = +inf, no gradient elements.
        static const float ninf = -INFINITY, pinf = INFINITY;
        for (int b = 0; b < a.batch_size; ++b)
            SSNT_CUDA(cudaMemcpyAsync(a.log_likelihood + b, &ninf, sizeof(float), cudaMemcpyHostToDevice, stream));
        if (a.loss) SSNT_CUDA(cudaMemcpyAsync(a.loss, &pinf, sizeof(float), cudaMemcpyHostToDevice, stream));
        return;
    }
    void* ws = a.workspace;
    const size_t need = fb_workspace_bytes(a.batch_size, a.max_t, a.max_u);
    if (ws) {
        SSNT_ASSERT(a.workspace_bytes >= need, "forward_backward: workspace too small");
    } else {
        ws = device_scratch(0, need);
    }
    unsigned* counter = done_counter_for(ws);

    auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    bool warp_ok = (a.max_u % 4 == 0) && a.max_u <= 1024 && aligned16(ws) &&
                   (a.logits ? aligned16(a.logits) && aligned16(a.grad_logits)
                             : aligned16(a.log_emit) && aligned16(a.log_shift) && aligned16(a.grad_emit) && aligned16(a.grad_shift));
    if (a.logits && !warp_ok) {
        // Raw logits on a shape the fused kernels do not take (max_u not a multiple of 4, > 1024, unaligned buffers):
        // form the log-probabilities, run the ordinary path, chain the gradient — through the tail of the workspace.
        launch_logits_unfused(a, ws, stream);
        return;
    }
    const bool bf_ok = warp_ok && a.max_u <= 256;
    const bool split_ok = bf_ok && (a.max_u == 64 || a.max_u == 128 || a.max_u == 256);
    int kind = tls_force_kind;
    // the time-parallel kernels (kind 6) win at every batch size measured (B = 4 .. 512, U = 64 / 128 / 256); the
    // single-kernel block-float paths (2: fused, 4: split-role) remain selectable
    if (kind < 0) kind = bf_ok ? 6 : (warp_ok ? 1 : 0);
    // Large batches: the warp-serial kernels (kind 8) move 25 bytes per cell instead of ~37 and win once every SM holds
    // enough utterances to hide one warp's row latency.  Measured on B200 (G cells/s, kind 6 vs 8), T=2000 U=256:
    // B=512 135/137, 1024 137/212, 4096 140/232; T=800 U=128: B=1024 157/191, 4096 163/224.
    if (tls_force_kind < 0 && kind == 6 && !a.logits) {
        const size_t per_sm = (size_t)a.batch_size / (size_t)sm_count();
        if ((a.max_u > 128 && per_sm >= 4) || (a.max_u > 64 && a.max_u <= 128 && per_sm >= 5) ||
            (a.max_u <= 64 && per_sm >= 10)) kind = 8;  // U=64 T=800: B=888 0.37 (kind 6) vs 0.34, B=2048 0.39 vs 0.46 (kind 8)
        // Very narrow (half the lanes of the time-parallel kernels idle) or very long lattices (T/U >= 20: with unbiased
        // rows the fronts are so steep that the chunk scheme re-runs utterances in the log domain; the warp-serial
        // kernels re-derive their frames every four rows and never did): B=1024 U=32 T=800 581 -> 337 us,
        // B=1024 U=32 T=1600 2050 -> 662 us, B=1024 U=64 T=1600 1122 -> 659 us.
        if ((a.max_u <= 32 && per_sm >= 6) || ((long long)a.max_t >= 20ll * a.max_u && per_sm >= 4)) kind = 8;
    }
    if (a.logits) {  // the raw-logit mode lives in the time-parallel kernels and in the log-domain warp kernel
        if (kind != 1 && kind != 6 && kind != 7 && kind != 10) kind = bf_ok ? 6 : 1;
    }
    if (kind == 6 || kind == 7 || kind == 10) {
        // Chunk length: 16 frames, except 8 for wide lattices (max_u > 128) once the sweeps no longer have an SM each —
        // the operators then cost half as much to build and the 2x longer sweep hides behind other utterances
        // (B=512 U=256 T=2000: 1.57 vs 1.95 ms).  Kind 10 forces the short chunks (max_u > 128 only).
        const bool short_chunks = a.max_u > 128 && (kind == 10 || (kind == 6 && tls_force_kind < 0 && (size_t)a.batch_size * 2 > (size_t)sm_count()));
        SSNT_ASSERT(kind != 10 || a.max_u > 128, "forward_backward: kind 10 (8-frame chunks) needs max_u > 128");
        const int L = short_chunks ? kTpLShort : kTpL;
        TpLayout tl;
        SSNT_ASSERT(bf_ok && tp_layout(a.batch_size, a.max_t, a.max_u, tl, L), "forward_backward: time-parallel kernels forced on an unsupported shape");
        tls_last_kind = kind;
        TpParams p;
        p.a = a;
        char* base = (char*)ws;
        p.Q = (float*)base;
        p.A = (float*)(base + tl.region0);
        p.Bv = (float*)(base + tl.region0 + tl.vec);
        p.zlg = (float*)(base + tl.region0 + 2 * tl.vec);
        p.status = (unsigned*)(base + tl.region0 + 2 * tl.vec + (((size_t)a.batch_size * 4 * sizeof(float) + 255) & ~(size_t)255));
        p.C = tl.C;
        p.UP = tl.UP;
        const size_t stage = (size_t)(L + 1) * tl.UP * sizeof(float);
        int NS = (int)((size_t)(200 * 1024) / stage);
        p.NS = NS > 16 ? 16 : NS;
        p.debug = 0;
#ifdef SSNT_BF_DEBUG_VARIANTS
        static const int k2dbg = [] { const char* e = std::getenv("SSNT_TP_DEBUG_K2"); return e ? std::atoi(e) : 0; }();
        p.debug = k2dbg;
#endif
        p.force_fallback = kind == 7 ? 1 : 0;  // kind 7: run the time-parallel kernels but force the log-domain re-run
        // 32-token exponent groups only while the sweeps are short (<= 50 steps: no re-run in 4800 random utterances at
        // T = 800; the per-step cost is ~10 % lower than with 16-token groups)
        // — and only while the lattice is not much longer than wide: with T / U beyond ~6 the fronts get so steep that
        // 32-token groups lose the sweeps' agreement (B=32 U=64 T=800, random inputs: every utterance was re-run in the
        // log domain, 265 us per step; with 16-token groups none, 33 us; U=96 T=800 likewise)
        const bool wide_groups = a.max_t <= 50 * kTpL && 4 * (long long)a.max_t <= 25 * (long long)a.max_u;
        if (tl.CPL == 2) { if (wide_groups) launch_tp<2>(p, stream); else launch_tp<2, kTpL, 16>(p, stream); }
        else if (tl.CPL == 4) { if (wide_groups) launch_tp<4>(p, stream); else launch_tp<4, kTpL, 16>(p, stream); }
        else if (short_chunks) launch_tp<8, kTpLShort>(p, stream);
        else launch_tp<8>(p, stream);
        // the log-domain kernel re-runs what was flagged (status != 0) and reduces the loss
#ifdef SSNT_BF_DEBUG_VARIANTS
        static const bool skip_log = [] { const char* e = std::getenv("SSNT_TP_DEBUG_STAGES"); return e && std::atoi(e) < 4; }();
        if (skip_log) return;
#endif
        LogParams lp;
        lp.a = a;
        lp.scratch = (float*)ws;
        lp.SU = tl.SU;
        lp.counter = counter;
        lp.only = p.status;
        lp.fallbacks = device_fallback_counter();
        const size_t stage_bytes = (size_t)kG * (2 * a.max_u + lp.SU) * sizeof(float);
        const bool latency_mode = (size_t)a.batch_size * 2 <= (size_t)sm_count();
        int LNS = (int)((latency_mode ? 192 * 1024 : 52 * 1024) / stage_bytes);
        lp.NS = LNS < 2 ? 2 : (LNS > 8 ? 8 : LNS);
        const size_t smem = 128 + (size_t)lp.NS * stage_bytes;
        static const bool pdl4 = [] { const char* e = std::getenv("SSNT_TP_PDL"); return e ? (std::atoi(e) & 4) != 0 : true; }();
        if (tl.CPL == 2) launch_warp<2>(lp, smem, stream, pdl4);
        else if (tl.CPL == 4) launch_warp<4>(lp, smem, stream, pdl4);
        else launch_warp<8>(lp, smem, stream, pdl4);
        return;
    }
    if (kind == 8 || kind == 9) {
        WsLayout wl;
        SSNT_ASSERT(bf_ok && ws_layout(a.batch_size, a.max_t, a.max_u, wl), "forward_backward: warp-serial kernels forced on an unsupported shape");
        tls_last_kind = kind;
        WsParams p;
        p.a = a;
        char* base = (char*)ws;
        p.A = (float*)(base + wl.region0);
        p.zlg = (float*)(base + wl.region0 + wl.vec);
        p.status = (unsigned*)(base + wl.region0 + wl.vec + (((size_t)a.batch_size * 2 * sizeof(float) + 255) & ~(size_t)255));
        p.C = wl.C;
        p.UP = wl.UP;
        p.R = wl.R;
        // forward ring depth: 3 stages once an SM holds twenty utterances (B=4096 U=256: 232 vs 216 G cells/s), else 2
        // so that more warps are resident (B=1024: 212 vs 201)
        p.NS = (size_t)a.batch_size >= (size_t)20 * sm_count() ? 3 : 2;
        p.force_fallback = kind == 9 ? 1 : 0;  // kind 9: run the kernels but force the log-domain re-run
        if (wl.CPL == 2) launch_ws<2, 16>(p, stream);
        else if (wl.CPL == 4) launch_ws<4, 16>(p, stream);
        else launch_ws<8, 8>(p, stream);
        // the log-domain kernel re-runs what was flagged (status != 0) and reduces the loss
        LogParams lp;
        lp.a = a;
        lp.scratch = (float*)ws;
        lp.SU = wl.SU;
        lp.counter = counter;
        lp.only = p.status;
        lp.fallbacks = device_fallback_counter();
        const size_t stage_bytes = (size_t)kG * (2 * a.max_u + lp.SU) * sizeof(float);
        const bool latency_mode = (size_t)a.batch_size * 2 <= (size_t)sm_count();
        int LNS = (int)((latency_mode ? 192 * 1024 : 52 * 1024) / stage_bytes);
        lp.NS = LNS < 2 ? 2 : (LNS > 8 ? 8 : LNS);
        const size_t smem = 128 + (size_t)lp.NS * stage_bytes;
        if (wl.CPL == 2) launch_warp<2>(lp, smem, stream, true);
        else if (wl.CPL == 4) launch_warp<4>(lp, smem, stream, true);
        else launch_warp<8>(lp, smem, stream, true);
        return;
    }
    SSNT_ASSERT(kind <= 7, "forward_backward: unknown kernel kind (0 .. 10)");
    if (kind >= 4) SSNT_ASSERT(split_ok && a.batch_size <= 64, "forward_backward: split kernel forced on an unsupported shape");
    if (kind == 1) SSNT_ASSERT(warp_ok, "forward_backward: warp kernel forced on an unsupported shape");
    if (kind >= 2) SSNT_ASSERT(bf_ok, "forward_backward: block-float kernel forced on an unsupported shape");
    tls_last_kind = kind;

    if (kind >= 4) {
        SplitParams p;
        p.a = a;
        p.SU = a.max_u + 32;
        p.nstp = (a.max_t + kG - 1) / kG;
        const size_t Af = (size_t)a.batch_size * 2 * p.nstp * kG * p.SU;
        p.A = (float*)ws;
        p.status = (unsigned*)(p.A + Af);
        p.fallbacks = device_fallback_counter();
        p.force_fallback = kind == 5 ? 1 : 0;  // kind 5: run the split kernel but force the log-domain re-run
        p.debug = 0;
#ifdef SSNT_BF_DEBUG_VARIANTS  // result-altering profiling knobs exist in debug builds only
        static const int split_debug = [] { const char* e = std::getenv("SSNT_SPLIT_DEBUG"); return e ? std::atoi(e) : 0; }();
        p.debug = split_debug;
#endif
        p.counter = counter;
        p.stats = tls_stats;
        const size_t slot_bytes = ((size_t)2 * kG * a.max_u) * sizeof(float);
        int NS = (int)((size_t)(224 * 1024 - kSplitHeaderBytes) / slot_bytes);
        NS = NS >= 16 ? 16 : (NS >= 8 ? 8 : 4);  // power of two (the recursion indexes the ring with masks)
        SSNT_ASSERT((size_t)NS * slot_bytes + kSplitHeaderBytes <= 224 * 1024, "forward_backward: ring does not fit shared memory");
        p.NS = NS;
        p.ring_bytes = (int)((size_t)NS * slot_bytes);
        const size_t smem = kSplitHeaderBytes + (size_t)NS * slot_bytes;
        if (a.max_u == 64) launch_split<2>(p, smem, stream);
        else if (a.max_u == 128) launch_split<4>(p, smem, stream);
        else launch_split<8>(p, smem, stream);
        return;
    }

    if (kind >= 2) {
        BfParams p;
        p.a = a;
        p.scratch = (float*)ws;
        p.SU = round_up4(a.max_u) + 32;
        p.status = (unsigned*)((char*)ws + (size_t)a.batch_size * (a.max_t + 1) * p.SU * sizeof(float));
        SSNT_CUDA(cudaMemsetAsync(p.status, 0, (size_t)a.batch_size * sizeof(unsigned), stream));  // the CTAs only OR into it
        p.fallbacks = device_fallback_counter();
        p.force_fallback = kind == 3 ? 1 : 0;  // kind 3: run the block-float kernel but force the log-domain re-run
        p.counter = counter;
        p.stats = tls_stats;
        const size_t stage_bytes = ((size_t)kG * (3 * a.max_u + p.SU) + 32) * sizeof(float);
        int NS = (int)((size_t)(224 * 1024 - kBfHeaderBytes) / stage_bytes);
        NS = NS > 12 ? 12 : NS;
        SSNT_ASSERT(NS >= 4, "forward_backward: ring does not fit shared memory");
        // Throughput mode (more CTAs than SMs): two CTAs per SM with a 6-stage ring each — for max_u <= 64
        // only.  Measured at B=1024 U=64: 70 vs 63 G cells/s; at B=512 U=128 the 128-register cap spills
        // in the recursion and the helpers are the bottleneck anyway: 86 vs 106 G cells/s.
        bool two_per_sm = (size_t)a.batch_size * 2 > (size_t)sm_count() && a.max_u <= 64 &&
                          kBfHeaderBytes + 6 * stage_bytes <= 112 * 1024;
        static const int env_two = [] { const char* e = std::getenv("SSNT_BF_TWO_PER_SM"); return e ? std::atoi(e) : 1; }();  // tuning aid
        two_per_sm = two_per_sm && env_two != 0;
        if (two_per_sm) NS = 6;
        p.NS = NS;
        // Few utterances (latency mode): one CTA per SM has to cover the whole HBM latency by
        // itself, so prefetch far ahead; many utterances: neighbours share the L2, stay modest.
        p.pf_rows = (size_t)a.batch_size * 2 <= (size_t)sm_count() ? 256 : 64;
        static const int env_pf = [] { const char* e = std::getenv("SSNT_BF_PF_ROWS"); return e ? std::atoi(e) : -1; }();  // tuning aid
        if (env_pf >= 0) p.pf_rows = env_pf;
        p.pf_sleep_ns = 64;
        static const int env_sleep = [] { const char* e = std::getenv("SSNT_BF_SLEEP_NS"); return e ? std::atoi(e) : -1; }();  // tuning aid
        if (env_sleep >= 0) p.pf_sleep_ns = env_sleep;
        p.debug_skip = 0;
#ifdef SSNT_BF_DEBUG_VARIANTS
        static const int env_skip = [] { const char* e = std::getenv("SSNT_BF_DEBUG_SKIP"); return e ? std::atoi(e) : 0; }();
        p.debug_skip = env_skip;
#endif
        const size_t smem = kBfHeaderBytes + (size_t)NS * stage_bytes;
        const int U = a.max_u;
        if (two_per_sm) {
            launch_bf<2, 2>(p, smem, stream);
        } else if (U <= 64) launch_bf<2>(p, smem, stream);
        else if (U <= 128) launch_bf<4>(p, smem, stream);
        else launch_bf<8>(p, smem, stream);
        return;
    }

    if (kind == 1) {
        LogParams p;
        p.a = a;
        p.scratch = (float*)ws;
        p.SU = round_up4(a.max_u) + 32;
        p.counter = counter;
        const size_t stage_bytes = (size_t)kG * (2 * a.max_u + p.SU) * sizeof(float);
        // Latency mode (few utterances): deep ring; throughput mode: keep several CTAs per SM.
        const bool latency_mode = (size_t)a.batch_size * 2 <= (size_t)sm_count();
        const size_t budget = latency_mode ? 192 * 1024 : 52 * 1024;
        int NS = (int)(budget / stage_bytes);
        NS = NS < 2 ? 2 : (NS > 8 ? 8 : NS);
        p.NS = NS;
        const size_t smem = 128 + (size_t)NS * stage_bytes;
        SSNT_ASSERT(smem <= 226 * 1024, "forward_backward: ring does not fit shared memory");
        const int U = a.max_u;
        if (U <= 32) launch_warp<1>(p, smem, stream);
        else if (U <= 64) launch_warp<2>(p, smem, stream);
        else if (U <= 128) launch_warp<4>(p, smem, stream);
        else if (U <= 256) launch_warp<8>(p, smem, stream);
        else if (U <= 512) launch_warp<16>(p, smem, stream);
        else launch_warp<32>(p, smem, stream);
    } else {
        GenericParams p;
        p.a = a;
        p.sval = (float*)ws;
        p.soff = (float*)ws + (size_t)a.batch_size * a.max_t * a.max_u;
        p.counter = counter;
        int threads = ((a.max_u + 31) / 32) * 32;
        threads = threads > 1024 ? 1024 : threads;
        const size_t smem = 4 * (size_t)(a.max_u + 2) * sizeof(float);
        if (smem > 48 * 1024)
            SSNT_CUDA(cudaFuncSetAttribute(fb_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        fb_generic_kernel<<<a.batch_size, threads, smem, stream>>>(p);
        SSNT_CUDA(cudaGetLastError());
    }
}

// ---- all-reduced loss of the latest call: the sum of the `world` entries the ranks' kernels stored here ------------
namespace {
__global__ void loss_allreduce_kernel(LossExchange* x, float* out, unsigned* err) {
    const int lane = threadIdx.x;
    const unsigned seq = *(volatile unsigned*)&x->seq;
    const int world = x->world;
    const unsigned long long* mine = x->peers[x->rank] + (size_t)(seq % (unsigned)kLossRing) * kLossMaxWorld;
    float v = 0.0f;
    bool ok = true;
    if (lane < world) {
        unsigned long long e = 0;
        unsigned spins = 0;
        for (;;) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(e) : "l"(mine + lane) : "memory");
            if ((unsigned)(e >> 32) == seq) break;
            if (++spins > (1u << 24)) { ok = false; break; }  // a peer never delivered (seconds): report, do not hang
            __nanosleep(100);
        }
        v = __uint_as_float((unsigned)e);
    }
    ok = __all_sync(0xffffffffu, ok);
    float acc = 0.0f;
    for (int r = 0; r < world; ++r) acc += __shfl_sync(0xffffffffu, v, r);  // rank order: identical bits on every rank
    if (lane == 0) {
        *out = ok ? acc : __int_as_float(0x7fc00000);
        if (!ok) atomicOr(err, (unsigned)kErrLossExchange);
    }
}
}  // namespace

void launch_loss_allreduce(float* out_device, cudaStream_t stream) {
    LossExchange* x = loss_exchange_device();
    SSNT_ASSERT(x != nullptr, "ssnt_tts_loss_allreduce: no loss exchange connected");
    loss_allreduce_kernel<<<1, 32, 0, stream>>>(x, out_device, device_error_flag());
    SSNT_CUDA(cudaGetLastError());
}

}  // namespace ssnt
