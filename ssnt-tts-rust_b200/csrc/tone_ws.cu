// Tone-latent marginalised lattice, warp-serial block-float kernels for large batches (SURVEY.md §8 a-TL): the
// throughput path, organised like fb_ws.cuh (kind 8 of the plain lattice).  One warp owns one utterance end to end;
// a lane owns CPL = max_u/32 consecutive tokens with their K tone classes (W = CPL*K values):
//
//   alpha'(u,k) = alpha(u,k) e(u,k) + tone(u,k) X(u-1),      X(u) = sum_k alpha(u,k) s(u,k)
//   beta (u,k)  = e(u,k) beta'(u,k) + s(u,k) Y'(u+1),        Y(u) = sum_k tone(u,k) beta(u,k)
//
//   tone_ws_forward_kernel   alpha row by row in registers (one shuffle per row: X of the lane's last token), rows
//                            arriving through a cp.async ring; alpha is stored only every L-th row (checkpoint) plus
//                            the forward likelihood.
//   tone_ws_backward_kernel  chunks of L rows from the last to the first: alpha re-run inside the chunk from its
//                            checkpoint (registers), then beta backward with the gradients of log_emit / log_shift fused
//                            and the gradient of log_tone accumulated in registers over the whole sweep (one warp owns
//                            the utterance: deterministic, no atomics).
//
// Every lane copies and reads back only its own tokens' values (16-byte cp.async, lane-interleaved in shared memory:
// conflict-free 128-bit accesses, no cross-lane visibility to wait for).  Memory traffic: inputs twice, gradients once,
// checkpoints 1/L of a row: ~(12 + 8/L)*K bytes per cell against 16*K algorithmic.
// Numerics and safety net as in fb_ws.cuh: one power-of-two exponent per lane, exact rescaling; every frame's
// occupancies must sum to 1 and the two sweeps' likelihoods must agree, else the utterance is flagged and re-run by the
// log-domain kernel (tone_fb_kernels.cu).
#include "fb_ws.cuh"

namespace ssnt {
using namespace lattice;

namespace {

thread_local int tls_tone_force = -1;
thread_local int tls_tone_last = -1;

template <int CPL, int K>
struct ToneWsCfg {
    static constexpr int W = CPL * K;                                   // values per lane and row
    static constexpr int Q = W / 4;                                     // 16-byte pieces per lane and row
    static constexpr int L = 64 / W < 2 ? 2 : (64 / W > 16 ? 16 : 64 / W);   // rows per chunk (L*W alpha registers in the backward kernel)
    static constexpr int RN = CPL < 4 ? (CPL < L ? CPL : L) : (4 < L ? 4 : L);  // rows between renormalisations
    static constexpr int NSF = 4;                                       // forward ring: stages of RN rows
    static constexpr int NSB = 2;                                       // backward ring: chunks
    static_assert(W % 4 == 0 && W <= 32 && L % RN == 0, "unsupported tone lattice shape");
};

struct ToneWsParams {
    ToneFbArgs a;
    float* A;          // [B][C+1][32*W + 32]  alpha checkpoints, lane-interleaved pieces + 32 lane frames (int)
    float* zlg;        // [B][2]               (log2 mantissa, frame as float) of the forward likelihood
    unsigned* status;  // [B]
    int C;             // checkpoints per utterance = ceil(max_t / L)
    int force_fallback;
};

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// lane-interleaved piece addressing: piece q of this lane inside a row of 32*Q pieces
template <int Q>
__device__ __forceinline__ void ldi(const float* row, int lane, float (&v)[4 * Q]) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(row + (q * 32 + lane) * 4);
        v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
    }
}
template <int Q>
__device__ __forceinline__ void sti(float* row, int lane, const float (&v)[4 * Q]) {
#pragma unroll
    for (int q = 0; q < Q; ++q)
        *reinterpret_cast<float4*>(row + (q * 32 + lane) * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
// this lane's W contiguous floats of a global row, streaming store
template <int Q>
__device__ __forceinline__ void stg_cs(float* p, const float (&v)[4 * Q]) {
#pragma unroll
    for (int q = 0; q < Q; ++q)
        __stcs(reinterpret_cast<float4*>(p + 4 * q), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
}

// One row of probabilities from the raw log-probs (lane-interleaved in shared memory).  Frames t >= T act as the
// identity (e = 1, s = 0); tokens >= U and the prohibited shifts (last token, last frame) are 0.
template <int CPL, int K>
__device__ __forceinline__ void tone_row_probs(const float* re_row, const float* rs_row, int lane, int t, int T, int U, int c0,
                                               float (&e)[CPL * K], float (&s)[CPL * K]) {
    constexpr int W = CPL * K, Q = W / 4;
    if (t < T) {
        float re[W], rs[W];
        ldi<Q>(re_row, lane, re);
        ldi<Q>(rs_row, lane, rs);
        const bool last = t == T - 1;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int u = c0 + w / K;
            e[w] = (u < U) ? ex2(to_log2(re[w])) : 0.0f;
            s[w] = (u < U - 1 && !last) ? ex2(to_log2(rs[w])) : 0.0f;
        }
    } else {
#pragma unroll
        for (int w = 0; w < W; ++w) { e[w] = 1.0f; s[w] = 0.0f; }
    }
}

__device__ __forceinline__ bool tone_lengths(const ToneFbArgs& a, int b, int& T, int& U) {
    T = a.t_len ? a.t_len[b] : a.max_t;
    U = a.u_len ? a.u_len[b] : a.max_u;
    T = min(max(T, 0), a.max_t);
    U = min(max(U, 0), a.max_u);
    return !(T <= 0 || U <= 0 || U > T);
}

// =================================================================================================
// Forward: alpha checkpoints and the forward likelihood.  One warp (one CTA) per utterance.
// =================================================================================================
template <int CPL, int K>
__global__ void __launch_bounds__(32) tone_ws_forward_kernel(const ToneWsParams p) {
    using Cfg = ToneWsCfg<CPL, K>;
    constexpr int W = Cfg::W, Q = Cfg::Q, L = Cfg::L, RN = Cfg::RN, NS = Cfg::NSF;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const ToneFbArgs& a = p.a;
    const int lane = threadIdx.x, b = blockIdx.x;
    tp_pdl_trigger();
    int T, U;
    if (!tone_lengths(a, b, T, U)) return;
    const int c0 = lane * CPL;
    const size_t rowf = (size_t)a.max_u * K;               // floats of a global row
    const size_t slab = (size_t)a.max_t * rowf;
    const float* le = a.log_emit + (size_t)b * slab + (size_t)c0 * K;
    const float* ls = a.log_shift + (size_t)b * slab + (size_t)c0 * K;
    float* ring = reinterpret_cast<float*>(smem_raw);     // [NS][RN][2][32*W], lane-interleaved pieces
    constexpr int row_floats = 32 * W;
    constexpr int stage_floats = RN * 2 * row_floats;
    const int nst = (T + RN - 1) / RN;
    auto issue = [&](int j) {
        if (j < nst) {
            float* st = ring + (size_t)(j % NS) * stage_floats;
#pragma unroll
            for (int r = 0; r < RN; ++r) {
                const int t = j * RN + r;
                if (t < a.max_t) {
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        cp_async16(smem_u32(st + (r * 2 + 0) * row_floats + (q * 32 + lane) * 4), le + (size_t)t * rowf + 4 * q);
                        cp_async16(smem_u32(st + (r * 2 + 1) * row_floats + (q * 32 + lane) * 4), ls + (size_t)t * rowf + 4 * q);
                    }
                }
            }
        }
        cp_async_commit();  // one group per call, possibly empty: the wait below counts groups
    };
#pragma unroll
    for (int j = 0; j < NS; ++j) issue(j);

    // tone priors of this lane's tokens (probabilities; 0 beyond U)
    float tn[W];
    {
        const float* lt = a.log_tone + (size_t)b * rowf + (size_t)c0 * K;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(lt + 4 * q);
            tn[4 * q] = w.x; tn[4 * q + 1] = w.y; tn[4 * q + 2] = w.z; tn[4 * q + 3] = w.w;
        }
#pragma unroll
        for (int w = 0; w < W; ++w) tn[w] = (c0 + w / K < U) ? ex2(to_log2(tn[w])) : 0.0f;
    }
    float v[W];
#pragma unroll
    for (int w = 0; w < W; ++w) v[w] = (c0 == 0 && w < K) ? tn[w] : 0.0f;   // alpha(0,0,k) = tone(0,k)
    int F = 0;
    float kin = lane == 0 ? 0.0f : 1.0f;
    float* ck = p.A + (size_t)b * (p.C + 1) * (row_floats + 32);
    for (int j = 0; j < nst; ++j) {
        cp_async_wait<NS - 1>();  // stage j has landed (this lane's own pieces: nothing to synchronise)
        const float* st = ring + (size_t)(j % NS) * stage_floats;
        const int tg = j * RN;
        if (tg % L == 0) {  // checkpoint: alpha(t) before frame t is applied
            float* row = ck + (size_t)(tg / L) * (row_floats + 32);
            sti<Q>(row, lane, v);
            reinterpret_cast<int*>(row + row_floats)[lane] = F;
        }
        float e[RN][W], s[RN][W];
#pragma unroll
        for (int r = 0; r < RN; ++r)
            tone_row_probs<CPL, K>(st + (r * 2 + 0) * row_floats, st + (r * 2 + 1) * row_floats, lane, tg + r, T, U, c0, e[r], s[r]);
#pragma unroll
        for (int r = 0; r < RN; ++r) {
            float X[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                float x = v[c * K] * s[r][c * K];
#pragma unroll
                for (int k = 1; k < K; ++k) x = fmaf(v[c * K + k], s[r][c * K + k], x);
                X[c] = x;
            }
            const float in = __shfl_up_sync(kFull, X[CPL - 1], 1) * kin;
#pragma unroll
            for (int c = CPL - 1; c >= 0; --c) {
                const float xin = c > 0 ? X[c - 1] : in;
#pragma unroll
                for (int k = 0; k < K; ++k) v[c * K + k] = fmaf(e[r][c * K + k], v[c * K + k], tn[c * K + k] * xin);
            }
        }
        ws_renorm<W, 0>(v, F, kin, lane);
        issue(j + NS);  // the slot just consumed
    }
    cp_async_wait<0>();
    // forward likelihood: sum_k alpha_T(U-1,k) (the last frame emits, shifts masked)
    const int zt = U - 1;
    if (zt / CPL == lane) {
        float yz = 0.0f;
#pragma unroll
        for (int w = 0; w < W; ++w) yz += (w / K == zt % CPL) ? v[w] : 0.0f;
        float* z = p.zlg + (size_t)b * 2;
        z[0] = yz > 0.0f ? log2f(yz) : -INFINITY;
        z[1] = (float)F;
    }
}

// =================================================================================================
// Backward: chunk by chunk from the end; alpha re-run from the checkpoint, beta with the gradients fused.
// =================================================================================================
template <int CPL, int K>
__global__ void __launch_bounds__(32) tone_ws_backward_kernel(const ToneWsParams p) {
    using Cfg = ToneWsCfg<CPL, K>;
    constexpr int W = Cfg::W, Q = Cfg::Q, L = Cfg::L, NSB = Cfg::NSB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const ToneFbArgs& a = p.a;
    const int lane = threadIdx.x, b = blockIdx.x;
    const int max_t = a.max_t;
    const int c0 = lane * CPL;
    const size_t rowf = (size_t)a.max_u * K;
    const size_t slab = (size_t)max_t * rowf;
    const float* le = a.log_emit + (size_t)b * slab + (size_t)c0 * K;
    const float* ls = a.log_shift + (size_t)b * slab + (size_t)c0 * K;
    float* ge = a.grad_emit + (size_t)b * slab + (size_t)c0 * K;
    float* gs = a.grad_shift + (size_t)b * slab + (size_t)c0 * K;
    float* gtone = a.grad_tone + (size_t)b * rowf + (size_t)c0 * K;
    const float zeros[W] = {};
    auto zero_rows = [&](int from, int to) {
        for (int t = from; t < to; ++t) {
            stg_cs<Q>(ge + (size_t)t * rowf, zeros);
            stg_cs<Q>(gs + (size_t)t * rowf, zeros);
        }
    };
    tp_pdl_trigger();  // the log-domain re-run kernel may be launched; it waits for this grid before reading status
    int T, U;
    const bool valid = tone_lengths(a, b, T, U);
    float* ring = reinterpret_cast<float*>(smem_raw);     // [NSB][L][2][32*W]
    constexpr int row_floats = 32 * W;
    constexpr int stage_floats = L * 2 * row_floats;
    const int Cb = valid ? (T + L - 1) / L : 0;
    auto issue = [&](int c, int slot) {   // chunk c of the raw rows into ring slot `slot`
        if (c >= 0) {
            float* st = ring + (size_t)slot * stage_floats;
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const int t = c * L + l;
                if (t < max_t) {
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        cp_async16(smem_u32(st + (l * 2 + 0) * row_floats + (q * 32 + lane) * 4), le + (size_t)t * rowf + 4 * q);
                        cp_async16(smem_u32(st + (l * 2 + 1) * row_floats + (q * 32 + lane) * 4), ls + (size_t)t * rowf + 4 * q);
                    }
                }
            }
        }
        cp_async_commit();
    };
    // the raw rows do not depend on the forward kernel: start their copies now
#pragma unroll
    for (int i = 0; i < NSB; ++i) issue(Cb - 1 - i, i);
    float tn[W];
    {
        const float* lt = a.log_tone + (size_t)b * rowf + (size_t)c0 * K;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(lt + 4 * q);
            tn[4 * q] = w.x; tn[4 * q + 1] = w.y; tn[4 * q + 2] = w.z; tn[4 * q + 3] = w.w;
        }
#pragma unroll
        for (int w = 0; w < W; ++w) tn[w] = (valid && c0 + w / K < U) ? ex2(to_log2(tn[w])) : 0.0f;
    }
    tp_pdl_wait();  // the forward kernel has completed: checkpoints and likelihoods are visible
    if (!valid) {
        cp_async_wait<0>();
        zero_rows(0, max_t);
        stg_cs<Q>(gtone, zeros);
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
        cp_async_wait<0>();
        return;
    }
    if (lane == 0) a.log_likelihood[b] = (float)(((double)zf_lg + (double)zf_ex) * kLn2);
    zero_rows(Cb * L < max_t ? Cb * L : max_t, max_t);

    // beta at the virtual terminal frame T: 1 for every tone of token U-1
    float bv[W], gt[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        bv[w] = (c0 + w / K == U - 1) ? 1.0f : 0.0f;
        gt[w] = 0.0f;
    }
    int eb = ((U - 1) / CPL == lane) ? 0 : kTpDead;
    const float* ck = p.A + (size_t)b * (p.C + 1) * (row_floats + 32);
    float worst = 0.0f;
    for (int c = Cb - 1; c >= 0; --c) {
        const int kk = Cb - 1 - c, slot = kk % NSB;
        const int t0 = c * L;
        // the chunk's checkpoint (issued before the wait for the rows)
        float av[W];
        const float* arow = ck + (size_t)c * (row_floats + 32);
        ldi<Q>(arow, lane, av);
        int ea = reinterpret_cast<const int*>(arow + row_floats)[lane];
        cp_async_wait<NSB - 1>();
        float* st = ring + (size_t)slot * stage_floats;
        // per-lane renormalisation (exact), then the frames held fixed over the chunk, one per lane:
        // F_l = max(ex_l, F_{l-1} - dec) for alpha (mass arrives from the left), F_l = max(ex_l, F_{l+1} - dec) for beta
        {
            float ma = av[0], mb = bv[0];
#pragma unroll
            for (int w = 1; w < W; ++w) { ma = fmaxf(ma, av[w]); mb = fmaxf(mb, bv[w]); }
            const int sha = ma > 0.0f ? ws_exponent(ma) : 0, shb = mb > 0.0f ? ws_exponent(mb) : 0;
            const float fa0 = tp_pow2(-sha), fb0 = tp_pow2(-shb);
#pragma unroll
            for (int w = 0; w < W; ++w) { av[w] *= fa0; bv[w] *= fb0; }
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
            for (int w = 0; w < W; ++w) { av[w] *= sa0; bv[w] *= sb0; }
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
        float ar[L][W];
        float inx[L];   // what entered this lane's first token from the left at each row (X of the left lane's last token), scaled by sa
#pragma unroll
        for (int l = 0; l < L; ++l) {
            float e[W], s[W];
            float* re_row = st + (l * 2 + 0) * row_floats;
            float* rs_row = st + (l * 2 + 1) * row_floats;
            tone_row_probs<CPL, K>(re_row, rs_row, lane, t0 + l, T, U, c0, e, s);
            sti<Q>(re_row, lane, e);  // raw rows overwritten in place by the probabilities (own pieces only)
            sti<Q>(rs_row, lane, s);
#pragma unroll
            for (int w = 0; w < W; ++w) ar[l][w] = av[w] * sa;
            float X[CPL];
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                float x = av[cc * K] * s[cc * K];
#pragma unroll
                for (int k = 1; k < K; ++k) x = fmaf(av[cc * K + k], s[cc * K + k], x);
                X[cc] = x;
            }
            const float in = __shfl_up_sync(kFull, X[CPL - 1], 1) * ka;
            inx[l] = in * sa;
            if (l < L - 1) {
#pragma unroll
                for (int cc = CPL - 1; cc >= 0; --cc) {
                    const float xin = cc > 0 ? X[cc - 1] : in;
#pragma unroll
                    for (int k = 0; k < K; ++k) av[cc * K + k] = fmaf(e[cc * K + k], av[cc * K + k], tn[cc * K + k] * xin);
                }
            }
        }
        // ---- beta backward with the gradients fused ----
        float csum = 0.0f;
        int crows = 0;
#pragma unroll
        for (int l = L - 1; l >= 0; --l) {
            const int t = t0 + l;
            float e[W], s[W];
            ldi<Q>(st + (l * 2 + 0) * row_floats, lane, e);
            ldi<Q>(st + (l * 2 + 1) * row_floats, lane, s);
            // Y(u) = sum_k tone(u,k) beta'(u,k) per token; the lane's last token takes Y of the right lane's first
            float tb[W], Y[CPL];
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                float yy = 0.0f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    tb[cc * K + k] = tn[cc * K + k] * bv[cc * K + k];
                    yy += tb[cc * K + k];
                }
                Y[cc] = yy;
            }
            const float bin = __shfl_down_sync(kFull, Y[0], 1) * kb;
            float g1[W], g2[W];
            float rowsum = 0.0f;
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const float yn = (cc + 1 < CPL ? Y[cc + 1] : bin) * sb;
                // X(t, u-1) scaled by sa: what enters token u at this frame (for the tone gradient)
                float xs;
                if (cc == 0) {
                    xs = inx[l];
                } else {
                    xs = ar[l][(cc - 1) * K] * s[(cc - 1) * K];
#pragma unroll
                    for (int k = 1; k < K; ++k) xs = fmaf(ar[l][(cc - 1) * K + k], s[(cc - 1) * K + k], xs);
                }
                xs *= sb;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int w = cc * K + k;
                    const float p1 = e[w] * bv[w];
                    g1[w] = ar[l][w] * (p1 * sb);
                    g2[w] = (ar[l][w] * s[w]) * yn;
                    rowsum += g1[w] + g2[w];
                    gt[w] = fmaf(xs, tb[w], gt[w]);   // identity rows (t >= T) have s = 0: xs = 0
                    bv[w] = fmaf(s[w], (cc + 1 < CPL ? Y[cc + 1] : bin), p1);
                }
            }
            if (t < T) {
                stg_cs<Q>(ge + (size_t)t * rowf, g1);
                stg_cs<Q>(gs + (size_t)t * rowf, g2);
                csum += rowsum;   // every frame's occupancies sum to 1: accumulated per lane, checked once per chunk
                ++crows;
            } else if (t < max_t) {
                stg_cs<Q>(ge + (size_t)t * rowf, zeros);
                stg_cs<Q>(gs + (size_t)t * rowf, zeros);
            }
        }
        csum = warp_sum(csum);
        worst = (fabsf(csum - (float)crows) <= kTpRowTol * (float)crows) ? worst : 1.0f;  // also catches NaN
        eb = fb;  // beta(t0) now sits in frame fb
        issue(c - NSB, slot);  // the slot is free (own pieces only: program order is enough)
    }
    cp_async_wait<0>();
    // backward likelihood sum_k tone(0,k) beta_0(0,k) against the forward one; token 0 draws its tone at the start
    float zb = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) zb += tn[k] * bv[k];
    float zb_lg = -INFINITY, zb_ex = 0.0f;
    if (lane == 0) {
        zb_lg = zb > 0.0f ? log2f(zb) : -INFINITY;
        zb_ex = (float)eb;
        const float sc = ex2(fminf(fmaxf(((float)eb - zf_ex) - zf_lg, -126.0f), 126.0f));
#pragma unroll
        for (int k = 0; k < K; ++k) gt[k] += (tn[k] * bv[k]) * sc;
    }
    stg_cs<Q>(gtone, gt);
    zb_lg = __shfl_sync(kFull, zb_lg, 0);
    zb_ex = __shfl_sync(kFull, zb_ex, 0);
    const float zdiff = (zf_ex - zb_ex) + (zf_lg - zb_lg);
    unsigned stw = 0u;
    if (!(zb_lg > -1e30f) || !(fabsf(zdiff) <= kTpZTol)) stw |= (unsigned)kTpBadZ;
    if (worst != 0.0f) stw |= (unsigned)kTpBadRow;
    if (lane == 0) p.status[b] = stw;
}

template <int CPL, int K>
void launch_tone_ws_t(const ToneWsParams& p, cudaStream_t stream) {
    using Cfg = ToneWsCfg<CPL, K>;
    const size_t fwd_smem = (size_t)Cfg::NSF * Cfg::RN * 2 * 32 * Cfg::W * sizeof(float);
    const size_t bwd_smem = (size_t)Cfg::NSB * Cfg::L * 2 * 32 * Cfg::W * sizeof(float);
    static bool configured_[64] = {};  // per device
    bool& configured = configured_[device_ordinal()];
    if (!configured) {
        SSNT_CUDA(cudaFuncSetAttribute(tone_ws_forward_kernel<CPL, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem));
        SSNT_CUDA(cudaFuncSetAttribute(tone_ws_backward_kernel<CPL, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem));
        configured = true;
    }
    tone_ws_forward_kernel<CPL, K><<<(unsigned)p.a.batch_size, 32, fwd_smem, stream>>>(p);
    SSNT_CUDA(cudaGetLastError());
    cudaLaunchAttribute pdl[1];
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.stream = stream;
    cfg.attrs = pdl;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)p.a.batch_size);
    cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = bwd_smem;
    SSNT_CUDA(cudaLaunchKernelEx(&cfg, tone_ws_backward_kernel<CPL, K>, p));
}

template <int CPL, int K>
constexpr int tone_ws_L() { return ToneWsCfg<CPL, K>::L; }

// chunk length of a supported shape, 0 otherwise
int tone_ws_chunk(int max_u, int K) {
    if (max_u % 32 != 0) return 0;
    const int cpl = max_u / 32;
    switch (cpl * 16 + K) {
        case 1 * 16 + 4: return tone_ws_L<1, 4>();
        case 1 * 16 + 8: return tone_ws_L<1, 8>();
        case 2 * 16 + 2: return tone_ws_L<2, 2>();
        case 2 * 16 + 4: return tone_ws_L<2, 4>();
        case 2 * 16 + 8: return tone_ws_L<2, 8>();
        case 4 * 16 + 2: return tone_ws_L<4, 2>();
        case 4 * 16 + 4: return tone_ws_L<4, 4>();
        case 4 * 16 + 8: return tone_ws_L<4, 8>();
        case 8 * 16 + 2: return tone_ws_L<8, 2>();
        case 8 * 16 + 4: return tone_ws_L<8, 4>();
        default: return 0;
    }
}

}  // namespace

void tone_force_kernel_kind(int kind) { tls_tone_force = kind; }
int tone_forced_kernel_kind() { return tls_tone_force; }
int tone_last_kernel_kind() { return tls_tone_last; }
void tone_note_kernel_kind(int kind) { tls_tone_last = kind; }

bool tone_ws_supported(const ToneFbArgs& a) {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    return tone_ws_chunk(a.max_u, a.tone_class_size) > 0 && al(a.log_emit) && al(a.log_shift) && al(a.log_tone) &&
           al(a.grad_emit) && al(a.grad_shift) && al(a.grad_tone);
}

size_t tone_ws_workspace_bytes(int B, int max_t, int max_u, int K) {
    const int L = tone_ws_chunk(max_u, K);
    if (L == 0 || B <= 0 || max_t <= 0) return 0;
    const size_t C = ((size_t)max_t + L - 1) / L;
    const size_t row = (size_t)max_u * K + 32;
    return ((size_t)B * (C + 1) * row + (size_t)B * 2) * sizeof(float) + (((size_t)B * sizeof(unsigned) + 255) & ~(size_t)255) + 512;
}

// Runs the warp-serial kernels; returns the device pointer of the [B] status words (non-zero = the utterance must be
// re-run in the log domain).
unsigned* launch_tone_ws(const ToneFbArgs& a, void* ws, int force_fallback, cudaStream_t stream) {
    ToneWsParams p;
    p.a = a;
    const int K = a.tone_class_size;
    const int L = tone_ws_chunk(a.max_u, K);
    SSNT_ASSERT(L > 0, "tone warp-serial kernels: unsupported shape");
    p.C = (a.max_t + L - 1) / L;
    const size_t row = (size_t)a.max_u * K + 32;
    p.A = (float*)ws;
    p.zlg = p.A + (size_t)a.batch_size * (p.C + 1) * row;
    p.status = (unsigned*)(p.zlg + (size_t)a.batch_size * 2);
    p.force_fallback = force_fallback;
    const int cpl = a.max_u / 32;
    switch (cpl * 16 + K) {
        case 1 * 16 + 4: launch_tone_ws_t<1, 4>(p, stream); break;
        case 1 * 16 + 8: launch_tone_ws_t<1, 8>(p, stream); break;
        case 2 * 16 + 2: launch_tone_ws_t<2, 2>(p, stream); break;
        case 2 * 16 + 4: launch_tone_ws_t<2, 4>(p, stream); break;
        case 2 * 16 + 8: launch_tone_ws_t<2, 8>(p, stream); break;
        case 4 * 16 + 2: launch_tone_ws_t<4, 2>(p, stream); break;
        case 4 * 16 + 4: launch_tone_ws_t<4, 4>(p, stream); break;
        case 4 * 16 + 8: launch_tone_ws_t<4, 8>(p, stream); break;
        case 8 * 16 + 2: launch_tone_ws_t<8, 2>(p, stream); break;
        default: launch_tone_ws_t<8, 4>(p, stream); break;
    }
    return p.status;
}

}  // namespace ssnt
