// SSNT lattice forward-backward (log-likelihood + posterior-occupancy gradients) for sm_100a.
//
// Specification: SURVEY.md §8 a-FB (the reference crate has no forward-backward; the Emit/Shift
// semantics are the decoding rules of src/lib.rs:187-225 — a frame either Emits (stay on token
// u) or Shifts (u→u+1), Shift from the last token is prohibited, the last frame must Emit at
// (T-1, U-1)).  T = output frames (serial), U = input tokens (across lanes).
//
// Numerics.  All lattice quantities are kept in the log2 domain (inputs are scaled by log2(e)
// on load, ex2.approx/lg2.approx do the log-add-exp).  Each stored row carries an
// integer-valued offset: true value = stored + offset, the offset being re-centred on the row
// maximum every 8 rows, so fp32 rounding acts on numbers of magnitude ~|distance to the row
// max| instead of ~|log-likelihood| (this is what keeps gradients ~1e-5 from the fp64 oracle
// at T=800).  -inf is represented by the finite sentinel kNeg so that x-y never produces NaN.
//
// Two kernels:
//  * fb_warp_kernel<CPL>  — the hot path.  One thread-block CLUSTER of two single-warp CTAs
//    per utterance: CTA rank 0 sweeps alpha forward from t=0, rank 1 sweeps beta backward from
//    t=T, concurrently (critical path T instead of 2T).  Each lane owns CPL consecutive tokens
//    in registers, so the u-1 / u+1 neighbour is one __shfl per row and no __syncthreads sits
//    in the recursion.  Phase 1 stores the first half of each sweep to scratch; after one
//    cluster barrier both CTAs know log-likelihood (sum over the meeting row) and phase 2
//    emits the gradients of the rows it walks, reading the partner's stored half, so every
//    lattice cell is read once per pass.  Rows of log_emit/log_shift (and of the partner's
//    scratch) are prefetched NS stages x 8 rows ahead by TMA bulk copies (cp.async.bulk →
//    mbarrier complete_tx) into a shared-memory ring; lanes read them with 128-bit LDS and
//    write gradients with 128-bit coalesced STG.
//  * fb_generic_kernel    — any shape/alignment (max_u % 4 != 0, U > 1024, unaligned bases):
//    one CTA per utterance, one thread per token, alpha row double-buffered in shared memory.
#include <cooperative_groups.h>

#include "ssnt_common.cuh"

namespace cg = cooperative_groups;

namespace ssnt {
namespace {

constexpr float kNeg = -1.0e30f;      // finite stand-in for -inf
constexpr float kNegTest = -1.0e29f;  // anything below counts as -inf
constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLn2 = 0.6931471805599453;
constexpr int kG = 8;                 // rows per pipeline stage (also the offset re-centring period)
constexpr unsigned kFull = 0xffffffffu;

thread_local int tls_force_kind = -1;
thread_local int tls_last_kind = -1;

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

// Deterministic loss = -sum_b ll[b]: the last CTA to finish adds the B values in index order.
__device__ void finish_loss(const float* ll, float* loss, int B, unsigned* counter, int lane, int nthreads) {
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
            *counter = 0u;  // hand the ticket back zeroed
        }
    }
    (void)nthreads;
}

// ===============================================================================================
// Hot path: cluster-of-two, warp-per-sweep kernel.
// ===============================================================================================
struct WarpParams {
    FbArgs a;
    float* scratch;      // [B][max_t + 1][SU]; row t = alpha(t) for t < m, beta(t) for t >= m
    int SU;              // scratch row stride in floats (multiple of 4); offset lives at [SU - 4]
    int NS;              // pipeline stages
    unsigned* counter;   // loss ticket
};

enum Mode { kAlpha1 = 0, kBeta1 = 1, kAlpha2 = 2, kBeta2 = 3 };

template <int CPL>
struct Sweep {
    // registers
    float v[CPL];   // alpha~ (rank 0) or beta~ (rank 1) of the current row, log2 domain
    float off;      // integer-valued offset: true = v + off
    // normalisation pipeline
    float nm;
    // log-likelihood bookkeeping (phase 2)
    float llt;      // LL~ (relative to offA(m-1) + offB(m))
    float offA_m1, offB_m;
    bool dead;      // no path with finite probability
};

template <int CPL>
__global__ void __launch_bounds__(32) fb_warp_kernel(const WarpParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();  // 0 = alpha sweep, 1 = beta sweep
    const int b = blockIdx.x >> 1;
    const FbArgs& a = p.a;
    const int max_t = a.max_t, max_u = a.max_u, SU = p.SU, NS = p.NS;
    int T = a.t_len ? a.t_len[b] : max_t;
    int U = a.u_len ? a.u_len[b] : max_u;
    T = min(max(T, 0), max_t);
    U = min(max(U, 0), max_u);
    const size_t slab = (size_t)max_t * max_u;
    const float* le = a.log_emit + (size_t)b * slab;
    const float* ls = a.log_shift + (size_t)b * slab;
    float* ge = a.grad_emit + (size_t)b * slab;
    float* gs = a.grad_shift + (size_t)b * slab;
    float* scr = p.scratch + (size_t)b * (max_t + 1) * SU;
    const int c0 = lane * CPL;

    const float zeros[CPL] = {};
    if (T <= 0 || U <= 0 || U > T) {
        // No monotonic path: ll = -inf, every gradient 0.  Uniform for both CTAs of the cluster.
        float* g = rank == 0 ? ge : gs;
        for (int t = 0; t < max_t; ++t) store_cells_cs<CPL>(g + (size_t)t * max_u, c0, max_u, zeros);
        if (rank == 0) {
            if (lane == 0) a.log_likelihood[b] = -INFINITY;
            finish_loss(a.log_likelihood, a.loss, a.batch_size, p.counter, lane, 32);
        }
        return;
    }

    // ---- shared-memory ring -------------------------------------------------------------------
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    float* ring = reinterpret_cast<float*>(smem_raw + 128);
    const int stage_floats = kG * (2 * max_u + SU);
    const int off_e = 0, off_s = kG * max_u, off_x = 2 * kG * max_u;
    if (lane == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(smem_u32(bars + s), 1);
        fence_mbar_init();
    }
    __syncwarp();

    const int m = (T + 1) >> 1;  // alpha phase 1: rows 0..m-1; beta phase 1: rows T..m
    unsigned kg = 0;             // global stage counter (slot = kg % NS, parity = (kg / NS) & 1)

    // One sweep = n consumption rows row(j) = t0 + dir*j.  Arrays: log_emit, log_shift (row r)
    // and, in phase 2, scratch (row r + xoff).
    auto issue = [&](int k, int n, int t0, int dir, bool with_x, int xoff, unsigned kbase) {
        const int j0 = k * kG;
        const int cnt = min(kG, n - j0);
        const int r0 = dir > 0 ? t0 + j0 : t0 - j0 - cnt + 1;
        const unsigned kk = kbase + (unsigned)k;
        const int slot = (int)(kk % (unsigned)NS);
        const uint32_t bar = smem_u32(bars + slot);
        float* dst = ring + (size_t)slot * stage_floats;
        const uint32_t row_bytes = (uint32_t)max_u * 4u;
        const uint32_t bytes_e = (uint32_t)cnt * row_bytes;
        const uint32_t bytes_x = with_x ? (uint32_t)cnt * (uint32_t)SU * 4u : 0u;
        mbar_expect_tx(bar, 2u * bytes_e + bytes_x);
        bulk_g2s(smem_u32(dst + off_e), le + (size_t)r0 * max_u, bytes_e, bar);
        bulk_g2s(smem_u32(dst + off_s), ls + (size_t)r0 * max_u, bytes_e, bar);
        if (with_x) bulk_g2s(smem_u32(dst + off_x), scr + (size_t)(r0 + xoff) * SU, bytes_x, bar);
    };

    Sweep<CPL> S;
    S.off = 0.0f;
    S.nm = kNeg;
    S.llt = 0.0f;
    S.offA_m1 = 0.0f;
    S.offB_m = 0.0f;
    S.dead = false;

    // Column masks are applied when rows are converted: tokens >= U read as log-prob -inf.
    auto convert = [&](float (&x)[CPL], bool all_masked) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) x[i] = (c0 + i < U && !all_masked) ? to_log2(x[i]) : kNeg;
    };

    // alpha step: v(t) → v(t+1) given row t.
    auto alpha_step = [&](const float (&E)[CPL], const float (&Sh)[CPL]) {
        float y[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) y[i] = S.v[i] + Sh[i];
        float yin = __shfl_up_sync(kFull, y[CPL - 1], 1);
        if (lane == 0) yin = kNeg;
#pragma unroll
        for (int i = CPL - 1; i >= 1; --i) S.v[i] = lae2(S.v[i] + E[i], y[i - 1]);
        S.v[0] = lae2(S.v[0] + E[0], yin);
    };
    // Offset re-centring, software-pipelined over the kG rows of a full stage so that the
    // 5-step warp max never sits on the recursion's dependency chain.
    auto renorm = [&](int q, bool full_stage) {
        if (!full_stage) return;
        if (q == 0) {
            float mx = S.v[0];
#pragma unroll
            for (int i = 1; i < CPL; ++i) mx = fmaxf(mx, S.v[i]);
            S.nm = mx;
        } else if (q <= 5) {
            S.nm = fmaxf(S.nm, __shfl_xor_sync(kFull, S.nm, 32 >> q));
        } else if (q == 6) {
            S.nm = S.nm > kNegTest ? rintf(S.nm) : 0.0f;
        } else {
#pragma unroll
            for (int i = 0; i < CPL; ++i) S.v[i] = fmaxf(S.v[i] - S.nm, kNeg);
            S.off += S.nm;
        }
    };
    auto store_state_row = [&](int t) {
        float* row = scr + (size_t)t * SU;
        store_cells<CPL>(row, c0, max_u, S.v);
        if (lane == 0) row[SU - 4] = S.off;
    };

    const int nphase1 = rank == 0 ? (m - 1) : (T - m);

    // =========================== phase 1 ===========================
    if (rank == 0) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) S.v[i] = (c0 + i == 0) ? 0.0f : kNeg;
        store_state_row(0);
    } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) S.v[i] = (c0 + i == U - 1) ? 0.0f : kNeg;
        store_state_row(T);  // virtual terminal row beta(T, .)
    }
    {
        const int n = nphase1;
        const int t0 = rank == 0 ? 0 : T - 1;
        const int dir = rank == 0 ? 1 : -1;
        const int nst = (n + kG - 1) / kG;
        if (lane == 0)
            for (int k = 0; k < min(NS, nst); ++k) issue(k, n, t0, dir, false, 0, kg);
        for (int k = 0; k < nst; ++k) {
            const unsigned kk = kg + (unsigned)k;
            const int slot = (int)(kk % (unsigned)NS);
            mbar_wait(smem_u32(bars + slot), (kk / (unsigned)NS) & 1u);
            const float* st = ring + (size_t)slot * stage_floats;
            const int j0 = k * kG;
            const int cnt = min(kG, n - j0);
            const bool full_stage = cnt == kG;
#pragma unroll
            for (int q = 0; q < kG; ++q) {
                if (q < cnt) {
                    const int t = t0 + dir * (j0 + q);
                    const int idx = dir > 0 ? q : cnt - 1 - q;
                    float E[CPL], Sh[CPL];
                    load_cells<CPL>(st + off_e + idx * max_u, c0, max_u, 0.0f, E);
                    load_cells<CPL>(st + off_s + idx * max_u, c0, max_u, 0.0f, Sh);
                    convert(E, false);
                    convert(Sh, t == T - 1);  // the last frame must emit
                    renorm(q, full_stage);
                    if (rank == 0) {
                        alpha_step(E, Sh);
                        store_state_row(t + 1);
                    } else {
                        float bs = __shfl_down_sync(kFull, S.v[0], 1);
                        if (lane == 31) bs = kNeg;
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            const float nb = (i + 1 < CPL) ? S.v[i + 1] : bs;
                            S.v[i] = lae2(E[i] + S.v[i], Sh[i] + nb);
                        }
                        store_state_row(t);
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && k + NS < nst) issue(k + NS, n, t0, dir, false, 0, kg);
        }
        kg += (unsigned)nst;
    }

    // Make this CTA's scratch rows visible to the partner's TMA reads, then meet.
    __threadfence();
    fence_proxy_async();
    cluster.sync();
    fence_proxy_async();

    // =========================== phase 2 ===========================
    // rank 0: rows t = m-1 .. T-1, scratch row t+1 = beta(t+1); first row only yields LL.
    // rank 1: rows t = m-1 .. 0,   scratch row t   = alpha(t);  every row emits gradients.
    {
        const int n = rank == 0 ? (T - m + 1) : m;
        const int t0 = m - 1;
        const int dir = rank == 0 ? 1 : -1;
        const int xoff = rank == 0 ? 1 : 0;
        const int nst = (n + kG - 1) / kG;
        if (lane == 0)
            for (int k = 0; k < min(NS, nst); ++k) issue(k, n, t0, dir, true, xoff, kg);
        for (int k = 0; k < nst; ++k) {
            const unsigned kk = kg + (unsigned)k;
            const int slot = (int)(kk % (unsigned)NS);
            mbar_wait(smem_u32(bars + slot), (kk / (unsigned)NS) & 1u);
            const float* st = ring + (size_t)slot * stage_floats;
            const int j0 = k * kG;
            const int cnt = min(kG, n - j0);
            const bool full_stage = cnt == kG;
#pragma unroll
            for (int q = 0; q < kG; ++q) {
                if (q < cnt) {
                    const int t = t0 + dir * (j0 + q);
                    const int idx = dir > 0 ? q : cnt - 1 - q;
                    float E[CPL], Sh[CPL], X[CPL];
                    load_cells<CPL>(st + off_e + idx * max_u, c0, max_u, 0.0f, E);
                    load_cells<CPL>(st + off_s + idx * max_u, c0, max_u, 0.0f, Sh);
                    const float* xrow = st + off_x + idx * SU;
                    load_cells<CPL>(xrow, c0, max_u, kNeg, X);
                    const float xoffv = xrow[SU - 4];  // partner's offset of that row (broadcast)
                    convert(E, false);
                    convert(Sh, t == T - 1);
                    const bool first = (j0 + q) == 0;  // t == m-1: the meeting row
                    renorm(q, full_stage);
                    // beta(t+1, u) and beta(t+1, u+1) (rank 0: from scratch; rank 1: registers)
                    float bn[CPL], x[CPL], y[CPL];
                    if (rank == 0) {
#pragma unroll
                        for (int i = 0; i < CPL; ++i) bn[i] = X[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < CPL; ++i) bn[i] = S.v[i];
                    }
                    float bsh = __shfl_down_sync(kFull, bn[0], 1);
                    if (lane == 31) bsh = kNeg;
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        const float nb = (i + 1 < CPL) ? bn[i + 1] : bsh;
                        x[i] = E[i] + bn[i];
                        y[i] = Sh[i] + nb;
                    }
                    if (first) {
                        // LL~ = LSE_u( alpha~(m-1,u) + beta~(m-1,u) ), identical bits in both CTAs.
                        float term[CPL];
                        float mx = kNeg;
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            const float av = rank == 0 ? S.v[i] : X[i];
                            term[i] = fmaxf(av + lae2(x[i], y[i]), kNeg);
                            mx = fmaxf(mx, term[i]);
                        }
                        mx = warp_max(mx);
                        float sum = 0.0f;
#pragma unroll
                        for (int i = 0; i < CPL; ++i) sum += ex2(term[i] - mx);
                        sum = warp_sum(sum);
                        S.llt = mx + lg2(sum);
                        S.dead = !(mx > kNegTest);
                        S.offA_m1 = rank == 0 ? S.off : xoffv;
                        S.offB_m = rank == 0 ? xoffv : S.off;
                        if (rank == 0 && lane == 0) {
                            const double ll2 = (double)S.llt + (double)S.offA_m1 + (double)S.offB_m;
                            a.log_likelihood[b] = S.dead ? -INFINITY : (float)(ll2 * kLn2);
                        }
                    }
                    const bool emit = rank == 1 || !first;
                    if (emit) {
                        const float offA_t = rank == 0 ? S.off : xoffv;
                        const float offB_t1 = rank == 0 ? xoffv : S.off;
                        const float kt = ((offA_t - S.offA_m1) + (offB_t1 - S.offB_m)) - S.llt;
                        float g1[CPL], g2[CPL];
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            const float av = rank == 0 ? S.v[i] : X[i];
                            g1[i] = S.dead ? 0.0f : ex2((av + x[i]) + kt);
                            g2[i] = S.dead ? 0.0f : ex2((av + y[i]) + kt);
                        }
                        store_cells_cs<CPL>(ge + (size_t)t * max_u, c0, max_u, g1);
                        store_cells_cs<CPL>(gs + (size_t)t * max_u, c0, max_u, g2);
                    }
                    if (rank == 0) {
                        if (t < T - 1) alpha_step(E, Sh);
                    } else {
#pragma unroll
                        for (int i = 0; i < CPL; ++i) S.v[i] = lae2(x[i], y[i]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && k + NS < nst) issue(k + NS, n, t0, dir, true, xoff, kg);
        }
        kg += (unsigned)nst;
    }

    // Padded frames t >= T: gradients are exactly 0 (rank 0 clears grad_emit, rank 1 grad_shift).
    {
        float* g = rank == 0 ? ge : gs;
        for (int t = T; t < max_t; ++t) store_cells_cs<CPL>(g + (size_t)t * max_u, c0, max_u, zeros);
    }
    if (rank == 0) finish_loss(a.log_likelihood, a.loss, a.batch_size, p.counter, lane, 32);
}

// ===============================================================================================
// Generic path: one CTA per utterance, thread per token (strided when U > blockDim).
// ===============================================================================================
struct GenericParams {
    FbArgs a;
    float* scratch;  // [B][max_t][max_u] alpha~ rows
    float* offs;     // [B][max_t] per-row offsets
    unsigned* counter;
};

__device__ float block_max(float v, float* red /*[32]*/) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float r = l < ((blockDim.x + 31) >> 5) ? red[l] : kNeg;
    r = warp_max(r);
    return r;
}
__global__ void fb_generic_kernel(const GenericParams p) {
    extern __shared__ float sm[];
    __shared__ float red[32];
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
    float* scr = p.scratch + (size_t)b * slab;
    float* offs = p.offs + (size_t)b * max_t;

    if (T <= 0 || U <= 0 || U > T) {
        for (size_t i = tid; i < slab; i += nt) { ge[i] = 0.0f; gs[i] = 0.0f; }
        if (tid == 0) a.log_likelihood[b] = -INFINITY;
        finish_loss(a.log_likelihood, a.loss, a.batch_size, p.counter, tid, nt);
        return;
    }
    // two rows of (U + 2) floats with a guard cell on each side: index u+1 ↔ token u
    float* cur = sm;
    float* nxt = sm + (max_u + 2);
    for (int i = tid; i < max_u + 2; i += nt) { cur[i] = kNeg; nxt[i] = kNeg; }
    __syncthreads();
    if (tid == 0) cur[1] = 0.0f;  // alpha(0,0) = 0
    __syncthreads();
    float off = 0.0f;
    // ---- forward: store alpha~(t) and its offset, then advance ----
    for (int t = 0; t < T; ++t) {
        if ((t & 15) == 15) {  // re-centre on the row maximum
            float mx = kNeg;
            for (int u = tid; u < U; u += nt) mx = fmaxf(mx, cur[u + 1]);
            mx = block_max(mx, red);
            const float c = mx > kNegTest ? rintf(mx) : 0.0f;
            for (int u = tid; u < U; u += nt) cur[u + 1] = fmaxf(cur[u + 1] - c, kNeg);
            off += c;
            __syncthreads();
        }
        for (int u = tid; u < U; u += nt) scr[(size_t)t * max_u + u] = cur[u + 1];
        if (tid == 0) offs[t] = off;
        if (t < T - 1) {
            for (int u = tid; u < U; u += nt) {
                const float e = to_log2(le[(size_t)t * max_u + u]);
                const float stay = cur[u + 1] + e;
                const float sh = u > 0 ? cur[u] + to_log2(ls[(size_t)t * max_u + u - 1]) : kNeg;
                nxt[u + 1] = lae2(stay, sh);
            }
            __syncthreads();
            float* tmp = cur; cur = nxt; nxt = tmp;
        }
    }
    // LL~ = alpha~(T-1,U-1) + le(T-1,U-1), relative to offA(T-1)
    __syncthreads();
    const float llt = cur[U] + to_log2(le[(size_t)(T - 1) * max_u + U - 1]);
    const float offA_last = off;
    const bool dead = !(llt > kNegTest);
    if (tid == 0) {
        const double ll2 = (double)llt + (double)offA_last;
        a.log_likelihood[b] = dead ? -INFINITY : (float)(ll2 * kLn2);
    }
    __syncthreads();
    // ---- backward with fused gradients: cur := beta~(t+1, .), virtual terminal row at t = T ----
    for (int i = tid; i < max_u + 2; i += nt) { cur[i] = kNeg; nxt[i] = kNeg; }
    __syncthreads();
    if (tid == 0) cur[U] = 0.0f;  // beta(T, U-1) = 0
    __syncthreads();
    float offB = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
        const float kt = ((offs[t] - offA_last) + offB) - llt;
        for (int u = tid; u < max_u; u += nt) {
            float g1 = 0.0f, g2 = 0.0f;
            if (u < U) {
                const float e = to_log2(le[(size_t)t * max_u + u]);
                const float s = (t == T - 1 || u == U - 1) ? kNeg : to_log2(ls[(size_t)t * max_u + u]);
                const float av = scr[(size_t)t * max_u + u];
                const float x = e + cur[u + 1];
                const float y = s + cur[u + 2];
                if (!dead) {
                    g1 = ex2((av + x) + kt);
                    g2 = ex2((av + y) + kt);
                }
                nxt[u + 1] = lae2(x, y);
            }
            ge[(size_t)t * max_u + u] = g1;
            gs[(size_t)t * max_u + u] = g2;
        }
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
        if ((t & 15) == 0 && t > 0) {
            float mx = kNeg;
            for (int u = tid; u < U; u += nt) mx = fmaxf(mx, cur[u + 1]);
            mx = block_max(mx, red);
            const float c = mx > kNegTest ? rintf(mx) : 0.0f;
            for (int u = tid; u < U; u += nt) cur[u + 1] = fmaxf(cur[u + 1] - c, kNeg);
            offB += c;
            __syncthreads();
        }
    }
    for (size_t i = (size_t)T * max_u + tid; i < slab; i += nt) { ge[i] = 0.0f; gs[i] = 0.0f; }
    finish_loss(a.log_likelihood, a.loss, a.batch_size, p.counter, tid, nt);
}

template <int CPL>
void launch_warp(const WarpParams& p, size_t smem, cudaStream_t stream) {
    static bool configured = false;  // per instantiation; idempotent attribute
    if (!configured || smem > 48 * 1024) {
        SSNT_CUDA(cudaFuncSetAttribute(fb_warp_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)p.a.batch_size * 2u);
    cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SSNT_CUDA(cudaLaunchKernelEx(&cfg, fb_warp_kernel<CPL>, p));
}

inline int round_up4(int x) { return (x + 3) & ~3; }

}  // namespace

// Workspace: scratch rows (+1 virtual row, stride max_u+4 rounded to 4) and per-row offsets for
// the generic kernel; sized for whichever kernel is picked.
size_t fb_workspace_bytes(int B, int max_t, int max_u) {
    if (B <= 0 || max_t <= 0 || max_u <= 0) return 256;
    const size_t SU = (size_t)round_up4(max_u) + 4;
    size_t warp_bytes = (size_t)B * (max_t + 1) * SU * sizeof(float);
    size_t gen_bytes = (size_t)B * max_t * max_u * sizeof(float) + (size_t)B * max_t * sizeof(float);
    size_t n = warp_bytes > gen_bytes ? warp_bytes : gen_bytes;
    return (n + 255) & ~(size_t)255;
}

int fb_last_kernel_kind() { return tls_last_kind; }
void fb_force_kernel_kind(int kind) { tls_force_kind = kind; }

void launch_forward_backward(const FbArgs& a, cudaStream_t stream) {
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
    unsigned* counter = next_done_counter();

    auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    bool warp_ok = (a.max_u % 4 == 0) && a.max_u <= 1024 && aligned16(a.log_emit) &&
                   aligned16(a.log_shift) && aligned16(a.grad_emit) && aligned16(a.grad_shift) &&
                   aligned16(ws);
    int kind = tls_force_kind;
    if (kind < 0) kind = warp_ok ? 1 : 0;
    if (kind == 1) SSNT_ASSERT(warp_ok, "forward_backward: warp kernel forced on an unsupported shape");
    tls_last_kind = kind;

    if (kind == 1) {
        WarpParams p;
        p.a = a;
        p.scratch = (float*)ws;
        p.SU = round_up4(a.max_u) + 4;
        p.counter = counter;
        const size_t stage_bytes = (size_t)kG * (2 * a.max_u + p.SU) * sizeof(float);
        // Latency mode (few utterances): deep ring; throughput mode: keep several CTAs per SM.
        const bool latency_mode = (size_t)a.batch_size * 2 <= (size_t)sm_count();
        const size_t budget = latency_mode ? 200 * 1024 : 52 * 1024;
        int NS = (int)(budget / stage_bytes);
        NS = NS < 2 ? 2 : (NS > 8 ? 8 : NS);
        p.NS = NS;
        const size_t smem = 128 + (size_t)NS * stage_bytes;
        SSNT_ASSERT(smem <= 227 * 1024, "forward_backward: ring does not fit shared memory");
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
        p.scratch = (float*)ws;
        p.offs = (float*)ws + (size_t)a.batch_size * a.max_t * a.max_u;
        p.counter = counter;
        int threads = ((a.max_u + 31) / 32) * 32;
        threads = threads > 1024 ? 1024 : threads;
        const size_t smem = 2 * (size_t)(a.max_u + 2) * sizeof(float);
        if (smem > 48 * 1024)
            SSNT_CUDA(cudaFuncSetAttribute(fb_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        fb_generic_kernel<<<a.batch_size, threads, smem, stream>>>(p);
        SSNT_CUDA(cudaGetLastError());
    }
}

}  // namespace ssnt
