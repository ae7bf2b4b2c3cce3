// Time-parallel lattice forward-backward (kernel kind 6): the T-serial recursion is cut into chunks
// of L frames whose transfer operators are built concurrently.
//
// The probability-domain recursion is linear:  alpha(t+1) = M_t alpha(t)  with the bidiagonal
//   (M_t x)(u) = e(t,u) x(u) + s(t,u-1) x(u-1),      e = exp(log_emit), s = exp(log_shift),
// and beta(t) = M_t^T beta(t+1) (SURVEY.md §8 a-FB; Emit/Shift semantics of src/lib.rs:187-225).
// For chunk c = frames [cL, cL+L) the product  P_c = M_{cL+L-1} ... M_{cL}  is lower-banded with
// bandwidth L+1, so it is held as L+1 diagonals  Q_c(i, d) = P_c(i, i-d).  Three kernels:
//
//   tp_build_kernel    one warp per (utterance, chunk): TMA-loads the chunk's 2·L rows, converts them
//                      (EX2, length masks) and runs the L-row recursion on all L+1 diagonals at once
//                      (registers; one shuffle per diagonal and row); writes Q_c.  Every chunk of
//                      every utterance is independent: B·T/L warps, one wave.
//   tp_combine_kernel  one warp per (utterance, direction): alpha_{c+1} = P_c alpha_c forward and
//                      beta_c = P_c^T beta_{c+1} backward — the SAME operator serves both sweeps —
//                      T/L banded mat-vecs instead of T dependent rows; operators stream through a TMA
//                      ring.  Boundary vectors carry one power-of-two exponent per lane (block float).
//   tp_fill_kernel     one warp per (utterance, chunk): from alpha_c and beta_{c+1} re-runs the L
//                      rows of its chunk (alpha forward into registers, beta backward with the
//                      gradients fused) and writes grad_emit / grad_shift with streaming stores.
//
// Critical path: ~2L rows + T/L mat-vecs instead of 2T rows.  Range: operators are plain fp32
// products of L probabilities; boundary vectors are block-float.  Every frame's occupancies must sum
// to 1 and the two sweeps' likelihoods must agree, else the utterance is flagged (status word) and
// re-run by the log-domain kernel (fb_log_warp.cuh), like the other block-float kernels do.
#pragma once
#include "fb_log_warp.cuh"

namespace ssnt {
namespace lattice {

#ifndef SSNT_TP_L
#define SSNT_TP_L 16
#endif
constexpr int kTpL = SSNT_TP_L;        // frames per chunk
constexpr float kTpRowTol = 2e-5f;     // |sum_u occupancy(t,u) - 1| beyond this flags the utterance
constexpr float kTpZTol = 3e-5f;       // |log2 Z_forward - log2 Z_backward| beyond this flags it

enum TpStatus : unsigned { kTpBadZ = 1u, kTpBadRow = 2u, kTpForced = 4u };

struct TpParams {
    FbArgs a;
    float* Q;          // [B][C][L+1][UP]   chunk operators, diagonal-major
    float* A;          // [B][C+1][UP+32]   alpha at chunk boundaries: UP mantissas + 32 lane exponents (int)
    float* Bv;         // [B][C+1][UP+32]   beta at chunk boundaries
    float* zlg;        // [B][2][2]         (log2 mantissa, exponent as float) of Z from the forward / backward sweep
    unsigned* status;  // [B]
    int C;             // chunks per utterance = ceil(max_t / L)
    int UP;            // padded token count = 32 * CPL
    int NS;            // combine ring stages
    int force_fallback;
    int G;             // tokens per block-float exponent of the boundary vectors (16 or 32)
    int debug;         // profiling aid (SSNT_TP_DEBUG_K2): see tp_combine_kernel; results are wrong when non-zero
};

// mbarrier wait by all 32 lanes that gives up (device printf + trap, i.e. a launch failure the host sees) instead of
// spinning forever if the awaited bulk copy never completes.
__device__ __forceinline__ void tp_wait(uint32_t bar, uint32_t parity, int code) {
    unsigned spins = 0;
    while (!__all_sync(kFull, mbar_try_wait(bar, parity))) {
        if (++spins > (1u << 22)) {
            if ((threadIdx.x & 31) == 0) printf("ssnt_tts_c: tp kernel wait %d timed out (block %d)\n", code, (int)blockIdx.x);
            __trap();
        }
    }
}

// Programmatic dependent launch: the four kernels of one call are chained with
// cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel's CTAs are launched (and run their prologue) while
// its predecessor drains; tp_pdl_wait() blocks until the predecessor has completed and its writes are visible.
__device__ __forceinline__ void tp_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void tp_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// 2^dd for dd <= 0 as an exact float; 0 below the normal range (flush).
__device__ __forceinline__ float tp_pow2_neg(int dd) {
    return __int_as_float(max(dd + 127, 0) << 23);
}
// 2^dd for dd in [-127, 127]; 0 below, clamped above.
__device__ __forceinline__ float tp_pow2(int dd) {
    return __int_as_float(min(max(dd + 127, 0), 254) << 23);
}

template <int CPL>
__device__ __forceinline__ void tp_load(const float* p, float (&v)[CPL]) {
    if constexpr (CPL >= 4) {
#pragma unroll
        for (int q = 0; q < CPL / 4; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(p + 4 * q);
            v[4 * q + 0] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
        }
    } else {
        const float2 w = *reinterpret_cast<const float2*>(p);
        v[0] = w.x; v[1] = w.y;
    }
}
template <int CPL>
__device__ __forceinline__ void tp_store(float* p, const float (&v)[CPL]) {
    if constexpr (CPL >= 4) {
#pragma unroll
        for (int q = 0; q < CPL / 4; ++q)
            *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q + 0], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    }
}

// Row accesses of the chunk kernels.  FULL: max_u == 32 * CPL, every lane's cells exist — plain 128-bit accesses, no
// bounds check (the checked form costs a branch with a reconvergence point per access, ~20 % of the fill kernel's
// instructions); otherwise the checked helpers of lattice_common.cuh.
template <int CPL, bool FULL>
__device__ __forceinline__ void tp_ld(const float* row, int c0, int max_u, float (&v)[CPL]) {
    if constexpr (FULL) tp_load<CPL>(row + c0, v);
    else load_cells<CPL>(row, c0, max_u, 0.0f, v);
}
template <int CPL, bool FULL>
__device__ __forceinline__ void tp_st(float* row, int c0, int max_u, const float (&v)[CPL]) {
    if constexpr (FULL) tp_store<CPL>(row + c0, v);
    else store_cells<CPL>(row, c0, max_u, v);
}
template <int CPL, bool FULL>
__device__ __forceinline__ void tp_st_cs(float* row, int c0, int max_u, const float (&v)[CPL]) {
    if constexpr (FULL) {
        if constexpr (CPL >= 4) {
#pragma unroll
            for (int q = 0; q < CPL / 4; ++q)
                __stcs(reinterpret_cast<float4*>(row + c0 + 4 * q), make_float4(v[4 * q + 0], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
        } else {
            __stcs(reinterpret_cast<float2*>(row + c0), make_float2(v[0], v[1]));
        }
    } else {
        store_cells_cs<CPL>(row, c0, max_u, v);
    }
}

// Sums N per-lane values over the 32 lanes, N values at once: at every butterfly level a lane keeps one half of the
// values and hands the other half to its partner, so the N sums cost N - 1 + (levels left) shuffles instead of 5 N and
// sit on no dependency chain.  On return v[0] holds the complete sum of ONE of the N values (which one depends on the
// lane; every value is held by 32 / N lanes).
template <int N>
__device__ __forceinline__ void tp_sum_many(float (&v)[N], int lane) {
    static_assert(N == 1 || N == 2 || N == 4 || N == 8 || N == 16 || N == 32, "power of two");
    int n = N;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        if (n > 1) {
            const bool hi = (lane & o) != 0;
            const int h = n / 2;
#pragma unroll
            for (int i = 0; i < N / 2; ++i) {
                if (i < h) {
                    const float send = hi ? v[i] : v[i + h];
                    const float keep = hi ? v[i + h] : v[i];
                    v[i] = keep + __shfl_xor_sync(kFull, send, o);
                }
            }
            n = h;
        } else {
            v[0] += __shfl_xor_sync(kFull, v[0], o);
        }
    }
}

// One lattice row of a chunk as probabilities, from the raw log-prob rows in shared memory.
// Frames t >= T act as the identity (e = 1, s = 0) so that the virtual terminal vector passes through
// a partial last chunk unchanged; tokens >= U and the prohibited shifts (last token, last frame) are 0.
template <int CPL, bool FULL = false>
__device__ __forceinline__ void tp_row_probs(const float* se, const float* ss, int l, int t, int T, int U, int max_u,
                                             int c0, float (&e)[CPL], float (&s)[CPL]) {
    if (t < T) {
        float re[CPL], rs[CPL];
        tp_ld<CPL, FULL>(se + l * max_u, c0, max_u, re);
        tp_ld<CPL, FULL>(ss + l * max_u, c0, max_u, rs);
        const bool last = t == T - 1;
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            e[r] = (c0 + r < U) ? ex2(re[r] * kLog2e) : 0.0f;   // ex2(-inf) = 0: no clamp needed (same bits as ex2(to_log2(.)))
            s[r] = (c0 + r < U - 1 && !last) ? ex2(rs[r] * kLog2e) : 0.0f;
        }
    } else {
#pragma unroll
        for (int r = 0; r < CPL; ++r) { e[r] = 1.0f; s[r] = 0.0f; }
    }
}

// sigmoid(z) and sigmoid(-z) from one EX2 and one RCP, both to full relative accuracy:
// t = 2^-|z'|, r = 1 / (1 + t):  sigmoid(|z|) = r, sigmoid(-|z|) = t r.
__device__ __forceinline__ void sigmoid_pair(float z, float& p, float& q) {
    const float zz = z * kLog2e;
    const float t = ex2(-fabsf(zz));
    const float r = __frcp_rn(1.0f + t);
    const float tr = t * r;
    p = zz >= 0.0f ? r : tr;
    q = zz >= 0.0f ? tr : r;
}

// Raw-logit mode: one lattice row as probabilities from the logit row in shared memory; (e, s) carry the lattice's
// masks like tp_row_probs, (pu, qu) are the unmasked sigmoid(z), sigmoid(-z) the gradient is chained through.
template <int CPL, bool FULL = false>
__device__ __forceinline__ void tp_row_probs_logits(const float* sz, int l, int t, int T, int U, int max_u, int c0,
                                                    float (&e)[CPL], float (&s)[CPL], float (&pu)[CPL], float (&qu)[CPL]) {
    if (t < T) {
        float rz[CPL];
        tp_ld<CPL, FULL>(sz + l * max_u, c0, max_u, rz);
        const bool last = t == T - 1;
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            sigmoid_pair(rz[r], pu[r], qu[r]);
            e[r] = (c0 + r < U) ? pu[r] : 0.0f;
            s[r] = (c0 + r < U - 1 && !last) ? qu[r] : 0.0f;
        }
    } else {
#pragma unroll
        for (int r = 0; r < CPL; ++r) { e[r] = 1.0f; s[r] = 0.0f; pu[r] = 0.0f; qu[r] = 0.0f; }
    }
}

// Starts the bulk copies of a chunk's raw rows (lane 0 only) and returns after arming the barrier.
__device__ __forceinline__ void tp_issue_chunk(const FbArgs& a, int b, int t0, int rows, float* se, float* ss, uint32_t bar) {
    const size_t slab = (size_t)a.max_t * a.max_u;
    const uint32_t bytes = (uint32_t)rows * (uint32_t)a.max_u * 4u;
    if (a.logits) {  // raw-logit mode: one tensor
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(se), a.logits + (size_t)b * slab + (size_t)t0 * a.max_u, bytes, bar);
        return;
    }
    mbar_expect_tx(bar, 2u * bytes);
    bulk_g2s(smem_u32(se), a.log_emit + (size_t)b * slab + (size_t)t0 * a.max_u, bytes, bar);
    bulk_g2s(smem_u32(ss), a.log_shift + (size_t)b * slab + (size_t)t0 * a.max_u, bytes, bar);
}

// The same chunk in kTpLoadStages pieces of L / kTpLoadStages rows, one mbarrier each (bars[0..kTpLoadStages)), so that
// the first rows can be worked on while the later ones are still in flight (the whole grid issues its loads at once:
// a chunk's last byte arrives microseconds after its first).  Pieces beyond `rows` are not issued and never awaited.
constexpr int kTpLoadStages = 4;
template <int L>
__device__ __forceinline__ void tp_issue_chunk_staged(const FbArgs& a, int b, int t0, int rows, float* se, float* ss, uint64_t* bars) {
    static_assert(L % kTpLoadStages == 0, "chunk length must be a multiple of the load stages");
    constexpr int R = L / kTpLoadStages;
#pragma unroll
    for (int s = 0; s < kTpLoadStages; ++s) {
        const int r = min(R, rows - s * R);
        if (r > 0) tp_issue_chunk(a, b, t0 + s * R, r, se + s * R * a.max_u, ss + s * R * a.max_u, smem_u32(bars + s));
    }
}

__device__ __forceinline__ bool tp_lengths(const FbArgs& a, int b, int& T, int& U) {
    T = a.t_len ? a.t_len[b] : a.max_t;
    U = a.u_len ? a.u_len[b] : a.max_u;
    T = min(max(T, 0), a.max_t);
    U = min(max(U, 0), a.max_u);
    return !(T <= 0 || U <= 0 || U > T);
}

// =================================================================================================
// Kernel 1: chunk operators.
// =================================================================================================
template <int CPL, int L, bool LG = false, bool FULL = false>
__global__ void __launch_bounds__(32) tp_build_kernel(const TpParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const FbArgs& a = p.a;
    const int lane = threadIdx.x;
    const int b = blockIdx.x / p.C, c = blockIdx.x % p.C;
    tp_pdl_trigger();  // the combine kernel's CTAs may be launched as soon as every build CTA has started
    int T, U;
    if (!tp_lengths(a, b, T, U)) return;
    const int t0 = c * L;
    if (t0 >= T) return;
    const int max_u = FULL ? 32 * CPL : a.max_u;  // FULL: a compile-time constant
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float* se = reinterpret_cast<float*>(smem_raw + 128);
    float* ss = se + L * max_u;
    const int rows = min(L, a.max_t - t0);
    if (lane == 0) {
        for (int s = 0; s < kTpLoadStages; ++s) mbar_init(smem_u32(bar + s), 1);
        fence_mbar_init();
    }
    // This grid is itself launched as a programmatic dependent of whatever precedes it in the stream (normally the
    // previous call's re-run kernel, which triggers at its start): everything above overlaps that kernel's tail; the
    // inputs may only be read and the workspace only be written once it has completed.
    tp_pdl_wait();
    if (lane == 0) tp_issue_chunk_staged<L>(a, b, t0, rows, se, ss, bar);
    __syncwarp();
    const int c0 = lane * CPL;
    const int src = (lane + 31) & 31;  // left neighbour, lane 0 wraps to lane 31 whose last shift is always 0

    float Q[CPL][L + 1];
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        Q[r][0] = 1.0f;
#pragma unroll
        for (int d = 1; d <= L; ++d) Q[r][d] = 0.0f;
    }
#pragma unroll
    for (int l = 0; l < L; ++l) {
        if (l % (L / kTpLoadStages) == 0 && l < rows) tp_wait(smem_u32(bar + l / (L / kTpLoadStages)), 0, 1);
        float e[CPL], s[CPL];
        if constexpr (LG) {
            float pu[CPL], qu[CPL];
            tp_row_probs_logits<CPL, FULL>(se, l, t0 + l, T, U, max_u, c0, e, s, pu, qu);
        } else {
            tp_row_probs<CPL, FULL>(se, ss, l, t0 + l, T, U, max_u, c0, e, s);
        }
        // what enters this lane's first token from the left neighbour's last token, per diagonal
        float X[L];
#pragma unroll
        for (int d = 0; d <= l; ++d) X[d] = __shfl_sync(kFull, s[CPL - 1] * Q[CPL - 1][d], src);
#pragma unroll
        for (int r = CPL - 1; r >= 1; --r) {
            Q[r][l + 1] = s[r - 1] * Q[r - 1][l];
#pragma unroll
            for (int d = l; d >= 1; --d) Q[r][d] = fmaf(e[r], Q[r][d], s[r - 1] * Q[r - 1][d - 1]);
            Q[r][0] = e[r] * Q[r][0];
        }
        Q[0][l + 1] = X[l];
#pragma unroll
        for (int d = l; d >= 1; --d) Q[0][d] = fmaf(e[0], Q[0][d], X[d - 1]);
        Q[0][0] = e[0] * Q[0][0];
    }
    float* qg = p.Q + ((size_t)b * p.C + c) * (size_t)(L + 1) * p.UP + c0;
#pragma unroll
    for (int d = 0; d <= L; ++d) {
        float w[CPL];
#pragma unroll
        for (int r = 0; r < CPL; ++r) w[r] = Q[r][d];
        tp_store<CPL>(qg + (size_t)d * p.UP, w);
    }
}

// =================================================================================================
// Kernel 2: boundary vectors.  blockIdx.x = 2*b + dir (0 forward / alpha, 1 backward / beta); one thread per
// token.  Step k multiplies the current vector (shared memory, zero-padded by L on both sides) by the k-th
// operator of the sweep:  forward  y'(i) = sum_d Q(i,d) y(i-d),  backward  y'(j) = sum_d Q(j+d,d) y(j+d);
// every shared-memory access is a conflict-free scalar load.
//
// Block float: one power-of-two exponent per warp (32 tokens).  A row's maximum can sit hundreds of bits above
// the entries the other sweep will meet (beta piles up at token 0 long before alpha gets there), so one exponent
// per row is not enough.  A thread's window reaches into one neighbouring warp only (L <= 32): it accumulates the
// two warps' contributions separately, each in its source frame, and joins them with exact power-of-two factors.
// The output frame of step k is predicted from the stored maxima of step k-1 plus the shrink observed one step
// earlier (feedback, lag one), so no reduction sits on the step's dependency chain: one barrier per step.
// =================================================================================================
constexpr int kTpDead = -(1 << 20);  // frame of a group whose tokens are all zero

// The compute warps' sweep (DIR 0 forward, 1 backward): one straight-line block per step.
//
// Frames run two steps ahead of the data: while step k multiplies vector k, the warp reads the maxima of vector k
// (published before the barrier), extrapolates the frame of vector k+2 from the shrink it observed over the last step,
// and derives the two power-of-two factors step k+1 will join its sums with.  The step's own dependency chain is then
// window loads -> FMAs -> two multiplies by register factors -> store, maximum, barrier.
template <int NT, int L, int NS, int G, int DIR>
__device__ __forceinline__ void tp_combine_sweep(const TpParams& p, int b, int U, int Cb, unsigned* wmax, int* fsm,
                                                 float* vbuf, const float* ring) {
    constexpr int VB = L + NT + L;
    constexpr int stage_floats = (L + 1) * NT;
    // exponent groups of G tokens (a warp or half a warp): `warp`/`lane` below are the group and the position in it
    constexpr int kTpGroup = G;
    const int tid = threadIdx.x, lane = tid & (kTpGroup - 1), warp = tid / kTpGroup;
    const bool upper = (tid & 16) != 0;                     // upper half-warp
    float* vec = (DIR == 0 ? p.A : p.Bv) + (size_t)b * (p.C + 1) * (NT + 32);
    const int hot0 = DIR == 0 ? 0 : U - 1;
    const int wn = DIR == 0 ? warp - 1 : warp + 1;        // the neighbouring group the window reaches into
    const bool has_nb = wn >= 0 && wn < NT / kTpGroup;
    const int wnc = has_nb ? wn : warp;
    const int wn2 = DIR == 0 ? warp - 2 : warp + 2;       // two groups away: its mass can arrive within two steps
    const bool has_nb2 = (2 * L > G) && wn2 >= 0 && wn2 < NT / kTpGroup;  // (mass crosses at most 2L tokens in two steps)
    const int wn2c = has_nb2 ? wn2 : warp;
    const bool hot_w = (hot0 / kTpGroup) == warp, hot_n = has_nb && (hot0 / kTpGroup) == wn;
    float y = tid == hot0 ? 1.0f : 0.0f;
    // frames of vector 0 and 1 (no shrink known yet), by the same rule as in the loop
    int F0 = hot_w ? 0 : kTpDead;                           // frame of vector k   (uniform within the warp)
    int F1 = hot_w ? 0 : (hot_n ? -24 : kTpDead);           // frame of vector k+1
    const int Fn0 = hot_n ? 0 : kTpDead;
    float c_own = tp_pow2(F0 - F1), c_nb = tp_pow2(Fn0 - F1);   // factors of step 0
    int a_prev = kTpDead;                                   // absolute exponent of this warp's maximum one vector ago
    int shrink2 = 0;                                        // extrapolated shrink over two steps (exponent, <= 0)
    for (int i = tid; i < 2 * VB; i += NT) vbuf[i] = 0.0f;
    if (tid < 32) wmax[tid] = 0u;                           // [2][16]
    if (tid < 64) fsm[tid] = kTpDead;                       // [4][16]
    __syncthreads();  // S0
    vbuf[L + tid] = y;
    if (lane == 0) {
        wmax[warp] = hot_w ? 0x3f800000u : 0u;  // the unit vector's maximum
        fsm[warp] = F0;
        fsm[16 + warp] = F1;
    }
    // global rows of the boundary vectors, walked in sweep order
    const ptrdiff_t rstride = DIR == 0 ? (NT + 32) : -(NT + 32);
    float* row = vec + (size_t)(DIR == 0 ? 0 : Cb) * (NT + 32);
    row[tid] = y;
    if (lane == 0) reinterpret_cast<int*>(row + NT)[warp] = F0;
    const float* ybase = vbuf + L + tid;
    const float* qbase = ring + tid;
#pragma unroll 1
    for (int k = 0; k < Cb; ++k) {
        const int cur = k & 1;
        __syncthreads();  // B_k: vector k, its warp maxima and the frames up to k+1 are visible; stage k has landed
        const unsigned mw = wmax[cur * 16 + warp];
        const unsigned mn = wmax[cur * 16 + wnc];
        const unsigned mn2 = wmax[cur * 16 + wn2c];
        const int Fnk = fsm[(k & 3) * 16 + wnc];
        const int Fn2k = fsm[(k & 3) * 16 + wn2c];
        const int Fnk1 = fsm[((k + 1) & 3) * 16 + wnc];
        const float* q = qbase + (size_t)(k % NS) * stage_floats;
        const float* yv = ybase + cur * VB;
        float yw[L + 1], qw[L + 1];
#pragma unroll
        for (int d = 0; d <= L; ++d) {
            if (DIR == 0) {
                yw[d] = yv[-d];
                qw[d] = q[d * NT];
            } else {
                yw[d] = yv[d];
                qw[d] = (tid + d < NT) ? q[d * NT + d] : 0.0f;
            }
        }
        float s_own = 0.0f, s_own2 = 0.0f, s_nb = 0.0f, s_nb2 = 0.0f;
#pragma unroll
        for (int d = 0; d <= L; ++d) {
            if ((p.debug & 1) && d > 0) break;
            const float t = qw[d] * yw[d];
            const bool own = DIR == 0 ? (d <= lane) : (lane + d < kTpGroup);
            if (own) { if (d & 1) s_own2 += t; else s_own += t; }
            else { if (d & 1) s_nb2 += t; else s_nb += t; }
        }
        // vector k+1 in frame F1:  y' = S_own 2^(F0 - F1) + S_nb 2^(Fn0 - F1)
        y = (s_own + s_own2) * c_own + (s_nb + s_nb2) * c_nb;
        vbuf[(cur ^ 1) * VB + L + tid] = y;
        // half-warp maxima as two full-warp reductions of masked values (a reduction over a run-time sub-mask compiles to a loop)
        const unsigned ybits = __float_as_uint(fmaxf(y, 0.0f));
        unsigned wm;
        if constexpr (G == 32) {
            wm = __reduce_max_sync(kFull, ybits);
        } else {
            const unsigned wm_lo = __reduce_max_sync(kFull, upper ? 0u : ybits);
            const unsigned wm_hi = __reduce_max_sync(kFull, upper ? ybits : 0u);
            wm = upper ? wm_hi : wm_lo;
        }
        // ---- off the chain: frame of vector k+2 and the factors of step k+1 ----
        const int aw = mw ? F0 + (int)(mw >> 23) - 127 : kTpDead;                        // absolute exponent of this warp's maximum
        const int an = (has_nb && mn) ? Fnk + (int)(mn >> 23) - 127 : kTpDead;           // and of the neighbour's
        if (aw > kTpDead / 2 && a_prev > kTpDead / 2) shrink2 = 2 * max(min(aw - a_prev, 0), -100);
        a_prev = aw;
        const int an2 = (has_nb2 && mn2) ? Fn2k + (int)(mn2 >> 23) - 127 : kTpDead;      // and two groups away
        // frame of vector k+2: what enters from the neighbours has been shifted at least once per group crossed
        int F2 = max(max(aw, an - 24), an2 - 48);
        F2 = F2 > kTpDead / 2 ? F2 + shrink2 : kTpDead;
        const int Fn1 = has_nb ? Fnk1 : kTpDead;
        // a factor beyond 2^126 means the extrapolation was far too low: raise the frame so that the larger is 2^126
        const int top = max(F1, Fn1) - F2;
        if (F2 > kTpDead / 2 && top > 126) F2 += top - 126;
        c_own = tp_pow2(F1 - F2);
        c_nb = tp_pow2(Fn1 - F2);
        if (lane == 0) {
            wmax[(cur ^ 1) * 16 + warp] = wm;
            fsm[((k + 2) & 3) * 16 + warp] = F2;
        }
        row += rstride;
        if (!(p.debug & 2)) {
            row[tid] = y;
            if (lane == 0) reinterpret_cast<int*>(row + NT)[warp] = F1;
        }
        F0 = F1;
        F1 = F2;
    }
    // Z: forward = alpha_C(U-1) (beta_C is the unit vector there), backward = beta_0(0).  (F0 is the last vector's frame.)
    if (tid == (DIR == 0 ? U - 1 : 0)) {
        float* z = p.zlg + (size_t)b * 4 + DIR * 2;
        z[0] = y > 0.0f ? log2f(y) : -INFINITY;
        z[1] = (float)F0;
    }
}

// NT compute threads (one per token) + one producer warp.  The producer issues the TMA copies of the operators
// and waits for the next stage's mbarrier BEFORE it arrives at the step's CTA barrier, so the compute warps never
// touch an mbarrier (a try_wait costs ~90 cycles even when the data has long landed).
template <int NT, int L, int NS, int G>
__global__ void __launch_bounds__(NT + 32) tp_combine_kernel(const TpParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    static_assert(L <= G && (G == 16 || G == 32), "a window must not reach beyond the neighbouring exponent group");
    static_assert(NS >= 2 && NS <= 24, "ring size");
    constexpr int VB = L + NT + L;  // one padded vector
    constexpr int stage_floats = (L + 1) * NT;
    const FbArgs& a = p.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x >> 1, dir = blockIdx.x & 1;
    int T, U;
    if (!tp_lengths(a, b, T, U)) return;
    if (dir == 0 && tid == 0) p.status[b] = 0u;  // the fill kernel ORs into it
    const int Cb = (T + L - 1) / L;
    unsigned* wmax = reinterpret_cast<unsigned*>(smem_raw);                  // [2][16] group maxima (float bits)
    int* fsm = reinterpret_cast<int*>(smem_raw + 128);                       // [4][16] group frames of vectors k..k+3 (mod 4)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 384);            // [NS <= 24]
    float* vbuf = reinterpret_cast<float*>(smem_raw + 576);                  // [2][VB]
    float* ring = vbuf + 2 * VB + ((4 - (2 * VB) % 4) % 4);                  // 16-byte aligned
    const bool producer = tid >= NT;
    tp_pdl_trigger();  // the fill kernel's CTAs may be launched (they wait for this grid's completion before reading)

    if (producer) {
        const float* qb = p.Q + (size_t)b * p.C * stage_floats;
        auto issue = [&](int k) {
            const int ck = dir == 0 ? k : Cb - 1 - k;
            const int slot = k % NS;
            const uint32_t bar = smem_u32(bars + slot);
            mbar_expect_tx(bar, (uint32_t)stage_floats * 4u);
            bulk_g2s(smem_u32(ring + (size_t)slot * stage_floats), qb + (size_t)ck * stage_floats, (uint32_t)stage_floats * 4u, bar);
        };
        if (lane == 0) {
            for (int s = 0; s < NS; ++s) mbar_init(smem_u32(bars + s), 1);
            fence_mbar_init();
        }
        tp_pdl_wait();  // the build kernel has completed: its operators are visible
        if (lane == 0)
            for (int k = 0; k < min(NS, Cb); ++k) issue(k);
        __syncwarp();
        __syncthreads();  // S0 (pairs with the compute warps' set-up barrier)
        tp_wait(smem_u32(bars), 0u, 2);
        for (int k = 0; k < Cb; ++k) {
            __syncthreads();  // B_k: every compute warp is done with stage k-1
            if (lane == 0 && k >= 1 && k - 1 + NS < Cb) issue(k - 1 + NS);
            if (k + 1 < Cb && !(p.debug & 4)) tp_wait(smem_u32(bars + ((k + 1) % NS)), (unsigned)((k + 1) / NS) & 1u, 2);
        }
        return;
    }

    if (dir == 0) tp_combine_sweep<NT, L, NS, G, 0>(p, b, U, Cb, wmax, fsm, vbuf, ring);
    else tp_combine_sweep<NT, L, NS, G, 1>(p, b, U, Cb, wmax, fsm, vbuf, ring);
}

// =================================================================================================
// Kernel 3: chunk interiors and gradients.
// =================================================================================================
template <int CPL, int L, bool LG = false, bool FULL = false>
__global__ void __launch_bounds__(32) tp_fill_kernel(const TpParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const FbArgs& a = p.a;
    const int lane = threadIdx.x;
    const int b = blockIdx.x / p.C, c = blockIdx.x % p.C;
    const int max_u = FULL ? 32 * CPL : a.max_u, max_t = a.max_t, UP = FULL ? 32 * CPL : p.UP;
    const int c0 = lane * CPL;
    const int t0 = c * L;
    const size_t slab = (size_t)max_t * max_u;
    float* ge = (LG ? a.grad_logits : a.grad_emit) + (size_t)b * slab;   // raw-logit mode: the one gradient tensor
    float* gs = LG ? nullptr : a.grad_shift + (size_t)b * slab;
    const float zeros[CPL] = {};
    auto zero_rows = [&](int from, int to) {
        for (int t = from; t < to; ++t) {
            tp_st_cs<CPL, FULL>(ge + (size_t)t * max_u, c0, max_u, zeros);
            if constexpr (!LG) tp_st_cs<CPL, FULL>(gs + (size_t)t * max_u, c0, max_u, zeros);
        }
    };
    int T, U;
    const int rows_end = min(t0 + L, max_t);
    tp_pdl_trigger();  // the log-domain re-run kernel may be launched; it waits for this grid before reading status
    // This grid may itself be launched as a programmatic dependent of the combine kernel (tuning mask bit 2): everything
    // up to tp_pdl_wait() touches only the call's inputs and gradient outputs (the build kernel, a full dependency of
    // whatever preceded the call, has completed), never what the combine kernel writes.
    const bool feasible = tp_lengths(a, b, T, U);
    if (!feasible || t0 >= T) {
        zero_rows(t0, rows_end);
        if (!feasible && c == 0 && lane == 0) {
            tp_pdl_wait();  // the combine kernel writes nothing for such an utterance, but keep the order of the status writes
            a.log_likelihood[b] = -INFINITY;
            p.status[b] = 0u;
        }
        return;
    }
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float* se = reinterpret_cast<float*>(smem_raw + 128);
    float* ss = se + L * max_u;
    const int rows = min(L, max_t - t0);
    if (lane == 0) {
        for (int s = 0; s < kTpLoadStages; ++s) mbar_init(smem_u32(bar + s), 1);
        fence_mbar_init();
        tp_issue_chunk_staged<L>(a, b, t0, rows, se, ss, bar);
    }
    __syncwarp();
    // ---- the chunk's rows as probabilities, in place (each lane re-reads only what it wrote itself) ----
#pragma unroll
    for (int l = 0; l < L; ++l) {
        if (l % (L / kTpLoadStages) == 0 && l < rows) tp_wait(smem_u32(bar + l / (L / kTpLoadStages)), 0, 3);
        float e[CPL], s[CPL];
        if constexpr (LG) {  // the UNMASKED pair: the beta sweep chains the gradient through it; masks are applied on use
            float pu[CPL], qu[CPL];
            tp_row_probs_logits<CPL, FULL>(se, l, t0 + l, T, U, max_u, c0, e, s, pu, qu);
            tp_st<CPL, FULL>(se + l * max_u, c0, max_u, pu);
            tp_st<CPL, FULL>(ss + l * max_u, c0, max_u, qu);
        } else {
            tp_row_probs<CPL, FULL>(se, ss, l, t0 + l, T, U, max_u, c0, e, s);
            tp_st<CPL, FULL>(se + l * max_u, c0, max_u, e);
            tp_st<CPL, FULL>(ss + l * max_u, c0, max_u, s);
        }
    }
    tp_pdl_wait();     // the combine kernel has completed: boundary vectors, likelihoods and status are visible
    // the two sweeps' likelihoods: (log2 mantissa, exponent)
    const float* z = p.zlg + (size_t)b * 4;
    const float zf_lg = z[0], zf_ex = z[1], zb_lg = z[2], zb_ex = z[3];
    const float zdiff = (zf_ex - zb_ex) + (zf_lg - zb_lg);
    const bool z_ok = (zf_lg > -1e30f) && (zb_lg > -1e30f) && fabsf(zdiff) <= kTpZTol && !p.force_fallback;
    if (!z_ok) {
        // no mass reached the end (a true -inf or an underflow) or the sweeps disagree: the log-domain kernel decides
        if (c == 0 && lane == 0) p.status[b] = p.force_fallback ? (unsigned)kTpForced : (unsigned)kTpBadZ;
#ifdef SSNT_TP_TRACE
        if (c == 0 && lane == 0) printf("tp: b=%d bad Z: fwd %f + %f, bwd %f + %f, diff %g\n", b, zf_lg, zf_ex, zb_lg, zb_ex, zdiff);
#endif
        return;
    }
    if (c == 0 && lane == 0) {
        a.log_likelihood[b] = (float)(((double)zf_lg + (double)zf_ex) * kLn2);
        // status[b] was zeroed by the combine kernel; a bad row below ORs into it
    }
    // boundary vectors and their per-lane exponents
    const float* arow = p.A + ((size_t)b * (p.C + 1) + c) * (UP + 32);
    const float* brow = p.Bv + ((size_t)b * (p.C + 1) + c + 1) * (UP + 32);
    float av[CPL], bv[CPL];
    tp_load<CPL>(arow + c0, av);
    tp_load<CPL>(brow + c0, bv);
    // one exponent per G tokens (16 or 32, the combine kernel's choice); this lane's CPL tokens lie in group lane*CPL/G.
    // Within a group the entries the other sweep meets can sit ~100 bits below the group's maximum (steep fronts at
    // U = 256), so each lane first re-normalises its own CPL mantissas (exact power-of-two scaling).
    int ea = reinterpret_cast<const int*>(arow + UP)[(lane * CPL) / p.G];
    int eb = reinterpret_cast<const int*>(brow + UP)[(lane * CPL) / p.G];
    {
        float ma = av[0], mb = bv[0];
#pragma unroll
        for (int r = 1; r < CPL; ++r) { ma = fmaxf(ma, av[r]); mb = fmaxf(mb, bv[r]); }
        const int sha = ma > 0.0f ? (int)((__float_as_uint(ma) >> 23) & 0xffu) - 127 : 0;
        const int shb = mb > 0.0f ? (int)((__float_as_uint(mb) >> 23) & 0xffu) - 127 : 0;
        const float fa0 = tp_pow2(-sha), fb0 = tp_pow2(-shb);
#pragma unroll
        for (int r = 0; r < CPL; ++r) { av[r] *= fa0; bv[r] *= fb0; }
        ea = (ma > 0.0f && ea > kTpDead / 2) ? ea + sha : kTpDead;
        eb = (mb > 0.0f && eb > kTpDead / 2) ? eb + shb : kTpDead;
    }
    // Frames held fixed over the chunk, one per lane: F_l = max(ex_l, F_{l-1} - dec) for alpha (mass arrives from
    // the left), F_l = max(ex_l, F_{l+1} - dec) for beta (from the right): a lane the front has not reached takes
    // its neighbour's frame lowered by dec, so that what enters it within L rows neither overflows nor flushes.
    constexpr int kDec0 = 96 / ((L + CPL - 1) / CPL);
    constexpr int kDec = kDec0 < 48 ? kDec0 : 48;
    int fa = ea + kDec * lane;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int other = __shfl_up_sync(kFull, fa, o);
        if (lane >= o) fa = max(fa, other);
    }
    fa -= kDec * lane;
    int fb = eb - kDec * lane;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int other = __shfl_down_sync(kFull, fb, o);
        if (lane + o < 32) fb = max(fb, other);
    }
    fb += kDec * lane;
    // (shuffles executed by all lanes, then selected: a shuffle under a lane-dependent condition is undefined)
    const int fa_left = __shfl_up_sync(kFull, fa, 1), fb_right = __shfl_down_sync(kFull, fb, 1);
    const float ka = lane == 0 ? 0.0f : tp_pow2(fa_left - fa);     // applied to what enters from lane-1
    const float kb = lane == 31 ? 0.0f : tp_pow2(fb_right - fb);   // applied to what enters from lane+1
    {
        const float sa0 = tp_pow2_neg(ea - fa), sb0 = tp_pow2_neg(eb - fb);
#pragma unroll
        for (int r = 0; r < CPL; ++r) { av[r] *= sa0; bv[r] *= sb0; }
    }
    // occupancy = alpha * (e|s) * beta / Z = (a * 2^xa) * (p * 2^xb) with xa + xb = fa + fb - log2 Z, split evenly so
    // that neither factor leaves the fp32 range before the product is formed
    float sa, sb;
    {
        const float xi = (float)(fa + fb) - zf_ex;     // integers: exact
        const float half = floorf(0.5f * xi);
        sa = ex2(fminf(fmaxf((xi - half) - zf_lg, -126.0f), 126.0f));
        sb = ex2(fminf(fmaxf(half, -126.0f), 126.0f));
    }

    // ---- alpha forward: rows 0..L-1 of the chunk kept in registers (scaled by sa) ----
    float ar[L][CPL];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        float e[CPL], s[CPL];
        tp_ld<CPL, FULL>(se + l * max_u, c0, max_u, e);
        tp_ld<CPL, FULL>(ss + l * max_u, c0, max_u, s);
        if constexpr (LG) {  // sigmoid(+-z) as stored above; the lattice's masks are applied here
            const int t = t0 + l;
            const bool last = t == T - 1;
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                e[r] = t < T ? ((c0 + r < U) ? e[r] : 0.0f) : 1.0f;
                s[r] = (t < T && c0 + r < U - 1 && !last) ? s[r] : 0.0f;
            }
        }
#pragma unroll
        for (int r = 0; r < CPL; ++r) ar[l][r] = av[r] * sa;
        if (l < L - 1) {
            const float in = __shfl_up_sync(kFull, s[CPL - 1] * av[CPL - 1], 1) * ka;
#pragma unroll
            for (int r = CPL - 1; r >= 1; --r) av[r] = fmaf(e[r], av[r], s[r - 1] * av[r - 1]);
            av[0] = fmaf(e[0], av[0], in);
        }
    }
    __syncwarp();
    // ---- beta backward with the gradients fused ----
    float rsum[L];  // per-lane share of every row's occupancy sum; the sums are formed after the sweep
#pragma unroll
    for (int l = L - 1; l >= 0; --l) {
        const int t = t0 + l;
        float e[CPL], s[CPL], pu[CPL], qu[CPL];
        if constexpr (LG) {  // sigmoid(+-z) as stored by the alpha sweep; the lattice's masks are re-applied here
            tp_ld<CPL, FULL>(se + l * max_u, c0, max_u, pu);
            tp_ld<CPL, FULL>(ss + l * max_u, c0, max_u, qu);
            const bool last = t == T - 1;
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                e[r] = t < T ? ((c0 + r < U) ? pu[r] : 0.0f) : 1.0f;
                s[r] = (t < T && c0 + r < U - 1 && !last) ? qu[r] : 0.0f;
            }
        } else {
            tp_ld<CPL, FULL>(se + l * max_u, c0, max_u, e);
            tp_ld<CPL, FULL>(ss + l * max_u, c0, max_u, s);
        }
        const float bin = __shfl_down_sync(kFull, bv[0], 1) * kb;
        float g1[CPL], g2[CPL];
        float rowsum = 0.0f;
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            const float p1 = e[r] * bv[r];
            const float p2 = s[r] * (r + 1 < CPL ? bv[r + 1] : bin);
            g1[r] = ar[l][r] * (p1 * sb);
            g2[r] = ar[l][r] * (p2 * sb);
            rowsum += g1[r] + g2[r];
            bv[r] = p1 + p2;
        }
        if (t < T) {
            if constexpr (LG) {  // dLL/dz = occupancy(emit) sigmoid(-z) - occupancy(shift) sigmoid(z)
                float gz[CPL];
#pragma unroll
                for (int r = 0; r < CPL; ++r) gz[r] = g1[r] * qu[r] - g2[r] * pu[r];
                tp_st_cs<CPL, FULL>(ge + (size_t)t * max_u, c0, max_u, gz);
            } else {
                tp_st_cs<CPL, FULL>(ge + (size_t)t * max_u, c0, max_u, g1);
                tp_st_cs<CPL, FULL>(gs + (size_t)t * max_u, c0, max_u, g2);
            }
            rsum[l] = rowsum;
        } else {
            rsum[l] = lane == 0 ? 1.0f : 0.0f;  // a frame beyond the utterance: nothing to check
            if (t < max_t) {
                tp_st_cs<CPL, FULL>(ge + (size_t)t * max_u, c0, max_u, zeros);
                if constexpr (!LG) tp_st_cs<CPL, FULL>(gs + (size_t)t * max_u, c0, max_u, zeros);
            }
        }
    }
    // every frame's occupancies must sum to 1
    tp_sum_many<L>(rsum, lane);
    const float dev = fabsf(rsum[0] - 1.0f);
    const bool bad = !(dev <= kTpRowTol);  // also catches NaN
#ifdef SSNT_TP_TRACE
    if (bad) printf("tp: b=%d chunk %d lane %d: a row sum is %g (fa %d fb %d sa %g sb %g)\n", b, c, lane, rsum[0], fa, fb, sa, sb);
#endif
    if (__any_sync(kFull, bad) && lane == 0) atomicOr(p.status + b, (unsigned)kTpBadRow);
}

}  // namespace lattice
}  // namespace ssnt
