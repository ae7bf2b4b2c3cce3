// Warp-serial lattice forward-backward for large batches (kernel kind 8): throughput mode.
//
// The time-parallel kernels (fb_tp.cuh, kind 6) buy latency with memory traffic — operators written and read back,
// inputs read twice, ~37 bytes per cell against 16 algorithmic — which is the right trade while there are fewer
// utterances than SMs.  With hundreds of utterances per GPU there is nothing to buy: every SM has several independent
// recursions to interleave, and the bound is HBM bandwidth.  Here one warp owns one utterance end to end
// (SURVEY.md §8 a-FB; Emit/Shift semantics of src/lib.rs:187-225):
//
//   ws_forward_kernel   alpha(t+1) = M_t alpha(t) row by row in registers (CPL = U/32 tokens per lane, one shuffle per
//                       row), rows streamed through a small TMA ring; stores alpha only every L-th row (a checkpoint,
//                       ~0.5 byte per cell) and the forward likelihood.
//   ws_backward_kernel  chunks of L rows from the last to the first: re-runs alpha inside the chunk from its checkpoint
//                       (registers), then beta backward with the gradients fused, beta carried from chunk to chunk in
//                       registers; the next chunk's rows arrive by TMA meanwhile.
//
// Memory traffic: inputs twice (8 + 8 B per cell), gradients once (8 B), checkpoints ~1 B: 25 bytes per cell.
// Numerics: probabilities with one power-of-two exponent ("frame") per lane, rescaled every few rows by exact powers
// of two; a lane's frame is never more than kWsGuard below its upstream neighbour's largest exponent, so what enters
// from there cannot overflow.  Same safety net as the other block-float kernels: every frame's occupancies must sum to
// 1 and the two sweeps' likelihoods must agree, else the utterance is flagged and re-run by the log-domain kernel.
#pragma once
#include "fb_tp.cuh"

namespace ssnt {
namespace lattice {

constexpr int kWsGuard = 32;
constexpr int kWsStageBytes = 8192;   // forward ring: rows per stage = kWsStageBytes / (8 * max_u)
constexpr int kWsFwdStagesMax = 3;     // forward ring stages: 3 when an SM holds many utterances, else 2 (WsParams::NS)
#ifndef SSNT_WS_BWD_STAGES
#define SSNT_WS_BWD_STAGES 1
#endif
constexpr int kWsBwdStages = SSNT_WS_BWD_STAGES;  // chunks in shared memory per warp in the backward kernel: one, so that
                                                  // twice as many warps fit an SM (B=4096 U=256: 232 vs 213 G cells/s; B=1024: 201 vs 166)

struct WsParams {
    FbArgs a;
    float* A;          // [B][C+1][UP+32]  alpha checkpoints: UP mantissas + 32 lane frames (int)
    float* zlg;        // [B][2]           (log2 mantissa, frame as float) of the forward likelihood
    unsigned* status;  // [B]
    int C;             // checkpoints per utterance = ceil(max_t / L)
    int UP;            // padded token count = 32 * CPL
    int R;             // rows per forward ring stage
    int NS;            // forward ring stages (<= kWsFwdStagesMax)
    int force_fallback;
};

__device__ __forceinline__ int ws_exponent(float m) { return (int)((__float_as_uint(m) >> 23) & 0xffu) - 127; }

// Renormalises one lane's values (exactly) and re-derives the frame and the factor applied to what enters from the
// upstream lane.  DIR 0: mass comes from lane-1 (alpha), DIR 1: from lane+1 (beta).
template <int CPL, int DIR>
__device__ __forceinline__ void ws_renorm(float (&v)[CPL], int& F, float& kin, int lane) {
    float m = v[0];
#pragma unroll
    for (int r = 1; r < CPL; ++r) m = fmaxf(m, v[r]);
    const bool alive = m > 0.0f;
    const int sh = alive ? ws_exponent(m) : 0;
    const int A = alive ? F + sh : kTpDead;
    int An = DIR == 0 ? __shfl_up_sync(kFull, A, 1) : __shfl_down_sync(kFull, A, 1);
    const bool edge = DIR == 0 ? lane == 0 : lane == 31;
    if (edge) An = kTpDead;
    int Fn = max(A, An - kWsGuard);
    if (Fn <= kTpDead / 2) Fn = F;  // nothing alive here nor upstream: keep the frame
    // one exact power-of-two factor takes the mantissas from frame F to frame Fn (values more than 126 bits below the
    // new frame flush to zero: they are that far below what is about to enter)
    const float f = tp_pow2(F - Fn);
#pragma unroll
    for (int r = 0; r < CPL; ++r) v[r] *= f;
    F = Fn;
    const int Fu = DIR == 0 ? __shfl_up_sync(kFull, F, 1) : __shfl_down_sync(kFull, F, 1);
    kin = edge ? 0.0f : tp_pow2(Fu - F);
}

// =================================================================================================
// Forward: alpha checkpoints and the forward likelihood.  One warp (one CTA) per utterance.
// =================================================================================================
template <int CPL, int L>
__global__ void __launch_bounds__(32) ws_forward_kernel(const WsParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NS = p.NS;
    constexpr int RN = CPL < 4 ? CPL : 4;  // rows between renormalisations: mass must not cross a whole lane in between
    const FbArgs& a = p.a;
    const int lane = threadIdx.x, b = blockIdx.x;
    tp_pdl_trigger();
    int T, U;
    if (!tp_lengths(a, b, T, U)) return;
    const int max_u = a.max_u, UP = p.UP, R = p.R;
    const int c0 = lane * CPL;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    float* ring = reinterpret_cast<float*>(smem_raw + 128);
    const int stage_floats = 2 * R * max_u;
    const size_t slab = (size_t)a.max_t * max_u;
    const int nst = (T + R - 1) / R;
    auto issue = [&](int j) {
        const int t0 = j * R, rows = min(R, a.max_t - t0);
        const uint32_t bar = smem_u32(bars + j % NS);
        float* se = ring + (size_t)(j % NS) * stage_floats;
        const uint32_t bytes = (uint32_t)rows * (uint32_t)max_u * 4u;
        mbar_expect_tx(bar, 2u * bytes);
        bulk_g2s(smem_u32(se), a.log_emit + (size_t)b * slab + (size_t)t0 * max_u, bytes, bar);
        bulk_g2s(smem_u32(se + R * max_u), a.log_shift + (size_t)b * slab + (size_t)t0 * max_u, bytes, bar);
    };
    if (lane == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(smem_u32(bars + s), 1);
        fence_mbar_init();
        for (int j = 0; j < min(NS, nst); ++j) issue(j);
    }
    __syncwarp();

    float v[CPL];
#pragma unroll
    for (int r = 0; r < CPL; ++r) v[r] = (c0 + r == 0) ? 1.0f : 0.0f;
    int F = 0;
    float kin = lane == 0 ? 0.0f : 1.0f;
    float* ck = p.A + (size_t)b * (p.C + 1) * (UP + 32);
    for (int j = 0; j < nst; ++j) {
        mbar_wait_warp(smem_u32(bars + j % NS), (unsigned)(j / NS) & 1u);
        const float* se = ring + (size_t)(j % NS) * stage_floats;
        const float* ss = se + R * max_u;
        // groups of RN rows: all conversions (EX2) of a group are issued before its RN dependent row updates, the
        // renormalisation closes the group.  Rows t >= T are identity rows, so the last group may run past T.
        for (int q0 = 0; q0 < R && j * R + q0 < T; q0 += RN) {
            const int tg = j * R + q0;
            if (tg % L == 0) {  // checkpoint: alpha(t) before frame t is applied  (L is a multiple of RN)
                float* row = ck + (size_t)(tg / L) * (UP + 32);
                tp_store<CPL>(row + c0, v);
                reinterpret_cast<int*>(row + UP)[lane] = F;
            }
            float e[RN][CPL], s[RN][CPL];
#pragma unroll
            for (int q = 0; q < RN; ++q) tp_row_probs<CPL>(se, ss, q0 + q, tg + q, T, U, max_u, c0, e[q], s[q]);
#pragma unroll
            for (int q = 0; q < RN; ++q) {
                const float in = __shfl_up_sync(kFull, s[q][CPL - 1] * v[CPL - 1], 1) * kin;
#pragma unroll
                for (int r = CPL - 1; r >= 1; --r) v[r] = fmaf(e[q][r], v[r], s[q][r - 1] * v[r - 1]);
                v[0] = fmaf(e[q][0], v[0], in);
            }
            ws_renorm<CPL, 0>(v, F, kin, lane);
        }
        __syncwarp();
        if (lane == 0 && j + NS < nst) issue(j + NS);
    }
    // forward likelihood: alpha_T(U-1) (the last frame emits, shifts masked)
    const int zt = U - 1;
    if (zt / CPL == lane) {
        float yz = v[0];
#pragma unroll
        for (int r = 1; r < CPL; ++r) yz = (zt % CPL == r) ? v[r] : yz;
        float* z = p.zlg + (size_t)b * 2;
        z[0] = yz > 0.0f ? log2f(yz) : -INFINITY;
        z[1] = (float)F;
    }
}

// =================================================================================================
// Backward: chunk by chunk from the end; alpha re-run from the checkpoint, beta with the gradients fused.
// =================================================================================================
template <int CPL, int L, bool FULL = false>
__global__ void __launch_bounds__(32) ws_backward_kernel(const WsParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const FbArgs& a = p.a;
    const int lane = threadIdx.x, b = blockIdx.x;
    const int max_u = FULL ? 32 * CPL : a.max_u, max_t = a.max_t, UP = FULL ? 32 * CPL : p.UP;  // FULL: every lane's cells exist
    const int c0 = lane * CPL;
    const size_t slab = (size_t)max_t * max_u;
    float* ge = a.grad_emit + (size_t)b * slab;
    float* gs = a.grad_shift + (size_t)b * slab;
    const float zeros[CPL] = {};
    auto zero_rows = [&](int from, int to) {
        for (int t = from; t < to; ++t) {
            tp_st_cs<CPL, FULL>(ge + (size_t)t * max_u, c0, max_u, zeros);
            tp_st_cs<CPL, FULL>(gs + (size_t)t * max_u, c0, max_u, zeros);
        }
    };
    tp_pdl_trigger();  // the log-domain re-run kernel may be launched; it waits for this grid before reading status
    int T, U;
    const bool valid = tp_lengths(a, b, T, U);
    constexpr int NSB = kWsBwdStages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // [NSB]
    float* ring = reinterpret_cast<float*>(smem_raw + 128);
    const int stage_floats = 2 * L * max_u;
    const int Cb = valid ? (T + L - 1) / L : 0;
    auto issue = [&](int c, int slot) {
        const int t0 = c * L;
        tp_issue_chunk(a, b, t0, min(L, max_t - t0), ring + (size_t)slot * stage_floats,
                       ring + (size_t)slot * stage_floats + L * max_u, smem_u32(bars + slot));
    };
    if (valid && lane == 0) {  // the raw rows do not depend on the forward kernel: start their copies now
        for (int i = 0; i < NSB; ++i) mbar_init(smem_u32(bars + i), 1);
        fence_mbar_init();
        for (int i = 0; i < NSB && Cb - 1 - i >= 0; ++i) issue(Cb - 1 - i, i);
    }
    __syncwarp();
    tp_pdl_wait();  // the forward kernel has completed: checkpoints and likelihoods are visible
    if (!valid) {
        zero_rows(0, max_t);
        if (lane == 0) {
            a.log_likelihood[b] = -INFINITY;
            p.status[b] = 0u;
        }
        return;
    }
    const float zf_lg = p.zlg[(size_t)b * 2], zf_ex = p.zlg[(size_t)b * 2 + 1];
    if (!(zf_lg > -1e30f) || p.force_fallback) {
        // no mass reached the end (a true -inf or an underflow): the log-domain kernel decides
        if (lane == 0) p.status[b] = p.force_fallback ? (unsigned)kTpForced : (unsigned)kTpBadZ;
        for (int i = 0; i < NSB && Cb - 1 - i >= 0; ++i) mbar_wait_warp(smem_u32(bars + i), 0u);  // no bulk copy left in flight
        return;
    }
    if (lane == 0) a.log_likelihood[b] = (float)(((double)zf_lg + (double)zf_ex) * kLn2);
    zero_rows(Cb * L < max_t ? Cb * L : max_t, max_t);

    // beta at the virtual terminal frame T: the unit vector at token U-1
    float bv[CPL];
#pragma unroll
    for (int r = 0; r < CPL; ++r) bv[r] = (c0 + r == U - 1) ? 1.0f : 0.0f;
    int eb = ((U - 1) / CPL == lane) ? 0 : kTpDead;
    const float* ck = p.A + (size_t)b * (p.C + 1) * (UP + 32);
    float worst = 0.0f;
    for (int c = Cb - 1; c >= 0; --c) {
        const int k = Cb - 1 - c, slot = k % NSB;
        const int t0 = c * L;
        // the chunk's checkpoint (issued before the wait for the rows)
        float av[CPL];
        const float* arow = ck + (size_t)c * (UP + 32);
        tp_load<CPL>(arow + c0, av);
        int ea = reinterpret_cast<const int*>(arow + UP)[lane];
        mbar_wait_warp(smem_u32(bars + slot), (unsigned)(k / NSB) & 1u);
        float* se = ring + (size_t)slot * stage_floats;
        float* ss = se + L * max_u;
        // per-lane renormalisation (exact), then the frames held fixed over the chunk, one per lane:
        // F_l = max(ex_l, F_{l-1} - dec) for alpha (mass arrives from the left), F_l = max(ex_l, F_{l+1} - dec) for beta
        {
            float ma = av[0], mb = bv[0];
#pragma unroll
            for (int r = 1; r < CPL; ++r) { ma = fmaxf(ma, av[r]); mb = fmaxf(mb, bv[r]); }
            const int sha = ma > 0.0f ? ws_exponent(ma) : 0, shb = mb > 0.0f ? ws_exponent(mb) : 0;
            const float fa0 = tp_pow2(-sha), fb0 = tp_pow2(-shb);
#pragma unroll
            for (int r = 0; r < CPL; ++r) { av[r] *= fa0; bv[r] *= fb0; }
            ea = (ma > 0.0f && ea > kTpDead / 2) ? ea + sha : kTpDead;
            eb = (mb > 0.0f && eb > kTpDead / 2) ? eb + shb : kTpDead;
        }
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
        const int fa_left = __shfl_up_sync(kFull, fa, 1), fb_right = __shfl_down_sync(kFull, fb, 1);
        const float ka = lane == 0 ? 0.0f : tp_pow2(fa_left - fa);     // applied to what enters from lane-1
        const float kb = lane == 31 ? 0.0f : tp_pow2(fb_right - fb);   // applied to what enters from lane+1
        {
            const float sa0 = tp_pow2_neg(ea - fa), sb0 = tp_pow2_neg(eb - fb);
#pragma unroll
            for (int r = 0; r < CPL; ++r) { av[r] *= sa0; bv[r] *= sb0; }
        }
        // occupancy = alpha * (e|s) * beta / Z = (a * 2^x1) * (p * 2^x2), x1 + x2 = fa + fb - log2 Z, split evenly
        float sa, sb;
        {
            const float xi = fmaxf((float)fa + (float)fb - zf_ex, -1000.0f);
            const float half = floorf(0.5f * xi);
            sa = ex2(fminf(fmaxf((xi - half) - zf_lg, -126.0f), 126.0f));
            sb = ex2(fminf(fmaxf(half, -126.0f), 126.0f));
        }
        // ---- alpha forward over the chunk, rows kept in registers (scaled by sa); probabilities written back ----
        float ar[L][CPL];
#pragma unroll
        for (int l = 0; l < L; ++l) {
            float e[CPL], s[CPL];
            tp_row_probs<CPL, FULL>(se, ss, l, t0 + l, T, U, max_u, c0, e, s);
            tp_st<CPL, FULL>(se + l * max_u, c0, max_u, e);  // raw rows overwritten in place by the probabilities
            tp_st<CPL, FULL>(ss + l * max_u, c0, max_u, s);  // (each lane re-reads only what it wrote itself)
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
        float csum = 0.0f;
        int crows = 0;
#pragma unroll
        for (int l = L - 1; l >= 0; --l) {
            const int t = t0 + l;
            float e[CPL], s[CPL];
            tp_ld<CPL, FULL>(se + l * max_u, c0, max_u, e);
            tp_ld<CPL, FULL>(ss + l * max_u, c0, max_u, s);
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
                tp_st_cs<CPL, FULL>(ge + (size_t)t * max_u, c0, max_u, g1);
                tp_st_cs<CPL, FULL>(gs + (size_t)t * max_u, c0, max_u, g2);
                // every frame's occupancies sum to 1: accumulated per lane, checked once per chunk below
                csum += rowsum;
                ++crows;
            } else if (t < max_t) {
                tp_st_cs<CPL, FULL>(ge + (size_t)t * max_u, c0, max_u, zeros);
                tp_st_cs<CPL, FULL>(gs + (size_t)t * max_u, c0, max_u, zeros);
            }
        }
        // the chunk's occupancies sum to its number of frames (one reduction per chunk instead of one per frame)
        csum = warp_sum(csum);
        worst = (fabsf(csum - (float)crows) <= kTpRowTol * (float)crows) ? worst : 1.0f;  // also catches NaN
        eb = fb;  // beta(t0) now sits in frame fb
        // the slot is free: fetch the chunk after next (generic-proxy writes above must not overtake the bulk copy)
        __syncwarp();
        if (lane == 0 && c - NSB >= 0) {
            fence_proxy_async();
            issue(c - NSB, slot);
        }
    }
    // backward likelihood beta_0(0) against the forward one
    float zb_lg = -INFINITY, zb_ex = 0.0f;
    if (lane == 0) {
        zb_lg = bv[0] > 0.0f ? log2f(bv[0]) : -INFINITY;
        zb_ex = (float)eb;
    }
    zb_lg = __shfl_sync(kFull, zb_lg, 0);
    zb_ex = __shfl_sync(kFull, zb_ex, 0);
    const float zdiff = (zf_ex - zb_ex) + (zf_lg - zb_lg);
    unsigned st = 0u;
    if (!(zb_lg > -1e30f) || !(fabsf(zdiff) <= kTpZTol)) st |= (unsigned)kTpBadZ;
    if (worst != 0.0f) st |= (unsigned)kTpBadRow;
    if (lane == 0) p.status[b] = st;
}

}  // namespace lattice
}  // namespace ssnt
