// Time-parallel lattice forward-backward with register sweeps (kernel kind 8).
//
// Same factorisation as fb_tp.cuh (kind 6) — alpha(t+1) = M_t alpha(t), beta(t) = M_t^T beta(t+1) with the
// bidiagonal M_t of the Emit/Shift lattice (SURVEY.md §8 a-FB; semantics of src/lib.rs:187-225) — but the frames are
// composed in GROUPS OF FOUR:  P_g = M_{4g+3} M_{4g+2} M_{4g+1} M_{4g}  has five diagonals, Q_g(d, i) = P_g(i, i-d).
// With five diagonals one warp holds a whole boundary vector in registers (CPL = U/32 tokens per lane, CPL >= 4) and a
// step needs only the neighbouring lane's last (first) four tokens: four shuffles and 5·CPL FMAs per four frames, no
// shared-memory vector, no CTA barrier.  Three kernels, chained with programmatic dependent launch:
//
//   tp4_build_kernel  one warp per (utterance, 16-frame chunk): TMA-loads the chunk's raw rows, converts them (EX2,
//                     length masks) and writes the four group operators of the chunk.
//   tp4_sweep_kernel  one warp per (utterance, direction): alpha_{g+1} = P_g alpha_g forward, beta_g = P_g^T beta_{g+1}
//                     backward, operators streamed through a TMA ring; writes every boundary vector with one
//                     power-of-two exponent ("frame") per lane.
//   tp4_fill_kernel   one warp per (utterance, 16-frame chunk): for each of its four groups re-runs the group's rows
//                     from alpha_g and beta_{g+1} (alpha forward into registers, beta backward with the gradients
//                     fused) and writes grad_emit / grad_shift with streaming stores.
//
// Frames.  A lane's values are mantissas times 2^F, F one integer per lane.  The frame used for vector k+1 is decided
// from vector k-1 (two steps of lag keep the decision off the dependency chain of the data): the lane's own largest
// exponent, but never more than kTp4Guard below those of the two upstream lanes — everything that can reach the lane
// within the lag comes from there and a step grows a value by at most 2^4, so the mantissas stay below 2^(guard+8);
// behind a steep front they grow by 2^40 per step, which a frame that only followed the lane itself would not survive.
// All rescaling is by exact powers of two.
// As with the other block-float kernels every frame's occupancies must sum to 1 and the two sweeps' likelihoods must
// agree, else the utterance is flagged and re-run by the log-domain kernel (fb_log_warp.cuh).
#pragma once
#include "fb_tp.cuh"

namespace ssnt {
namespace lattice {

constexpr int kTp4L = 4;        // frames per group
constexpr int kTp4Rows = 16;    // frames per build / fill task (four groups)
constexpr int kTp4GPS = 4;      // groups per ring stage of the sweep kernel
constexpr int kTp4Guard = 64;   // a lane's frame is at most this far below its upstream neighbour's largest exponent

template <int CPL>
struct Tp4Dims {
    static constexpr int UP = 32 * CPL;
    static constexpr int UPQ = UP + 4;                    // one diagonal: UP entries + 4 zeros (shifted reads of the backward sweep)
    static constexpr int group_floats = (kTp4L + 1) * UPQ;
    static constexpr int stage_floats = kTp4GPS * group_floats;
    static constexpr int NS = CPL <= 4 ? 8 : 6;           // ring stages: 84 KB (U <= 128), 125 KB (U <= 256)
};

// unbiased exponent of a positive float (denormals count as 2^-127)
__device__ __forceinline__ int tp4_exponent(float m) { return (int)((__float_as_uint(m) >> 23) & 0xffu) - 127; }

// row[D + k], k = 0..CPL-1, with the widest aligned vector loads (row is 16-byte aligned)
template <int CPL, int D>
__device__ __forceinline__ void tp4_load_shifted(const float* row, float (&v)[CPL]) {
    constexpr int h1 = D & 1;
    constexpr int h2 = (((D + h1) & 3) == 2) ? 2 : 0;
    constexpr int body0 = h1 + h2;
    constexpr int nb = (CPL - body0) / 4;
    constexpr int t0 = body0 + 4 * nb;
    constexpr int t2 = (CPL - t0 >= 2) ? 2 : 0;
    constexpr int t1 = CPL - t0 - t2;
    if constexpr (h1 != 0) v[0] = row[D];
    if constexpr (h2 != 0) {
        const float2 w = *reinterpret_cast<const float2*>(row + D + h1);
        v[h1] = w.x; v[h1 + 1] = w.y;
    }
#pragma unroll
    for (int q = 0; q < nb; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(row + D + body0 + 4 * q);
        v[body0 + 4 * q] = w.x; v[body0 + 4 * q + 1] = w.y; v[body0 + 4 * q + 2] = w.z; v[body0 + 4 * q + 3] = w.w;
    }
    if constexpr (t2 != 0) {
        const float2 w = *reinterpret_cast<const float2*>(row + D + t0);
        v[t0] = w.x; v[t0 + 1] = w.y;
    }
    if constexpr (t1 != 0) v[CPL - 1] = row[D + CPL - 1];
}

// =================================================================================================
// Kernel 1: group operators.  Q [B][NG][5][UPQ], NG = ceil(max_t / 4).
// =================================================================================================
template <int CPL>
__global__ void __launch_bounds__(32) tp4_build_kernel(const TpParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int L = kTp4L, R = kTp4Rows, UPQ = Tp4Dims<CPL>::UPQ;
    const FbArgs& a = p.a;
    const int lane = threadIdx.x;
    const int nchunk = (a.max_t + R - 1) / R;
    const int b = blockIdx.x / nchunk, c = blockIdx.x % nchunk;
    tp_pdl_trigger();
    int T, U;
    if (!tp_lengths(a, b, T, U)) return;
    const int t0 = c * R;
    if (t0 >= T) return;
    const int max_u = a.max_u;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float* se = reinterpret_cast<float*>(smem_raw + 128);
    float* ss = se + R * max_u;
    if (lane == 0) {
        mbar_init(smem_u32(bar), 1);
        fence_mbar_init();
        tp_issue_chunk(a, b, t0, min(R, a.max_t - t0), se, ss, smem_u32(bar));
    }
    __syncwarp();
    const int c0 = lane * CPL;
    const int src = (lane + 31) & 31;  // left neighbour; lane 0 wraps to lane 31, whose last shift is always 0
    tp_wait(smem_u32(bar), 0, 1);
#pragma unroll
    for (int gi = 0; gi < R / L; ++gi) {
        const int g = c * (R / L) + gi;
        if (g * L < T) {
            float Q[CPL][L + 1];
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                Q[r][0] = 1.0f;
#pragma unroll
                for (int d = 1; d <= L; ++d) Q[r][d] = 0.0f;
            }
#pragma unroll
            for (int l = 0; l < L; ++l) {
                float e[CPL], s[CPL];
                tp_row_probs<CPL>(se, ss, gi * L + l, g * L + l, T, U, max_u, c0, e, s);
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
            float* qg = p.Q + ((size_t)b * p.C + g) * (size_t)Tp4Dims<CPL>::group_floats + c0;
#pragma unroll
            for (int d = 0; d <= L; ++d) {
                float w[CPL];
#pragma unroll
                for (int r = 0; r < CPL; ++r) w[r] = Q[r][d];
                tp_store<CPL>(qg + (size_t)d * UPQ, w);
                if (lane == 31) *reinterpret_cast<float4*>(qg + (size_t)d * UPQ + CPL) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

// =================================================================================================
// Kernel 2: boundary vectors, one warp per (utterance, direction).  A, Bv [B][NG+1][UP+32]: UP mantissas + 32 frames.
// =================================================================================================
template <int CPL, int DIR>
__device__ __forceinline__ void tp4_load_q(const float* grp, int c0, float (&q)[kTp4L + 1][CPL]) {
    constexpr int UPQ = Tp4Dims<CPL>::UPQ;
    if constexpr (DIR == 0) {
#pragma unroll
        for (int d = 0; d <= kTp4L; ++d) tp_load<CPL>(grp + d * UPQ + c0, q[d]);
    } else {
        tp4_load_shifted<CPL, 0>(grp + 0 * UPQ + c0, q[0]);
        tp4_load_shifted<CPL, 1>(grp + 1 * UPQ + c0, q[1]);
        tp4_load_shifted<CPL, 2>(grp + 2 * UPQ + c0, q[2]);
        tp4_load_shifted<CPL, 3>(grp + 3 * UPQ + c0, q[3]);
        tp4_load_shifted<CPL, 4>(grp + 4 * UPQ + c0, q[4]);
    }
}

template <int CPL, int DIR>
__device__ __forceinline__ void tp4_sweep(const TpParams& p, int b, int U, int Gb, int lane, uint64_t* bars, const float* ring) {
    constexpr int L = kTp4L, UP = Tp4Dims<CPL>::UP, GPS = kTp4GPS, NS = Tp4Dims<CPL>::NS;
    constexpr int group_floats = Tp4Dims<CPL>::group_floats, stage_floats = Tp4Dims<CPL>::stage_floats;
    const int c0 = lane * CPL;
    const int nst = (Gb + GPS - 1) / GPS;
    const float* qb = p.Q + (size_t)b * p.C * group_floats;
    auto stage_span = [&](int j, int& lo, int& n) {
        if (DIR == 0) { lo = j * GPS; n = min(GPS, Gb - lo); }
        else { const int hi = Gb - j * GPS; lo = max(hi - GPS, 0); n = hi - lo; }
    };
    auto issue = [&](int j) {
        int lo, n;
        stage_span(j, lo, n);
        const int slot = j % NS;
        const uint32_t bar = smem_u32(bars + slot);
        const uint32_t bytes = (uint32_t)n * (uint32_t)group_floats * 4u;
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(ring + (size_t)slot * stage_floats), qb + (size_t)lo * group_floats, bytes, bar);
    };
    // location of step k's operator in the ring
    auto group_ptr = [&](int k) -> const float* {
        const int j = k / GPS, i = k % GPS;
        int lo, n;
        stage_span(j, lo, n);
        return ring + (size_t)(j % NS) * stage_floats + (size_t)(DIR == 0 ? i : n - 1 - i) * group_floats;
    };
    tp_pdl_wait();  // the build kernel has completed: its operators are visible
    if (lane == 0)
        for (int j = 0; j < min(NS, nst); ++j) issue(j);
    __syncwarp();

    float* vec = (DIR == 0 ? p.A : p.Bv) + (size_t)b * (p.C + 1) * (UP + 32);
    const int hot = DIR == 0 ? 0 : U - 1;
    float y[CPL];
#pragma unroll
    for (int r = 0; r < CPL; ++r) y[r] = (c0 + r == hot) ? 1.0f : 0.0f;
    int F = 0, Fn = 0;         // frames of the current vector and of the next one
    float c = 1.0f;            // 2^-(Fn - F)
    float kin = 0.0f;          // 2^(F_neighbour - F): applied to what enters from the upstream lane
    {
        const bool edge = DIR == 0 ? lane == 0 : lane == 31;
        kin = edge ? 0.0f : 1.0f;
    }
    const bool edge = DIR == 0 ? lane == 0 : lane == 31;
    {
        float* row = vec + (size_t)(DIR == 0 ? 0 : Gb) * (UP + 32);
        tp_store<CPL>(row + c0, y);
        reinterpret_cast<int*>(row + UP)[lane] = 0;
    }
    float q[2][L + 1][CPL];
    tp_wait(smem_u32(bars), 0u, 2);
    tp4_load_q<CPL, DIR>(group_ptr(0), c0, q[0]);

#pragma unroll 1
    for (int j = 0; j < nst; ++j) {
#pragma unroll
        for (int i = 0; i < GPS; ++i) {
            const int k = j * GPS + i;
            if (k < Gb) {
                // ---- prefetch the next step's operator (registers) ----
                if (k + 1 < Gb) {
                    if (i == GPS - 1) tp_wait(smem_u32(bars + ((j + 1) % NS)), (unsigned)((j + 1) / NS) & 1u, 2);
                    tp4_load_q<CPL, DIR>(group_ptr(k + 1), c0, q[(i + 1) & 1]);
                }
                // ---- off the chain: frame of vector k+1, from vector k-1 (the current y) ----
                float m = y[0];
#pragma unroll
                for (int r = 1; r < CPL; ++r) m = fmaxf(m, y[r]);
                const bool alive = m > 0.0f;
                const int A = alive ? F + tp4_exponent(m) : kTpDead;
                // own largest exponent, but at most kTp4Guard below the two upstream lanes': whatever can arrive within
                // the two steps of lag comes from there, and one step grows a value by at most 2^4
                int A1 = DIR == 0 ? __shfl_up_sync(kFull, A, 1) : __shfl_down_sync(kFull, A, 1);
                int A2 = DIR == 0 ? __shfl_up_sync(kFull, A, 2) : __shfl_down_sync(kFull, A, 2);
                if (DIR == 0 ? lane < 1 : lane > 30) A1 = kTpDead;
                if (DIR == 0 ? lane < 2 : lane > 29) A2 = kTpDead;
                int F2 = max(A, max(A1, A2) - kTp4Guard);
                if (F2 <= kTpDead / 2) F2 = Fn;  // nothing alive within reach: keep the frame
                const int dF = min(max(F2 - Fn, -126), 126);   // the rescaling factor must be a normal float
                F2 = Fn + dF;
                const float c_next = tp_pow2(-dF);
                const int Fnb = DIR == 0 ? __shfl_up_sync(kFull, Fn, 1) : __shfl_down_sync(kFull, Fn, 1);
                const float kin_next = edge ? 0.0f : tp_pow2(Fnb - Fn);
                // ---- the step ----
                float xin[L];
#pragma unroll
                for (int jj = 0; jj < L; ++jj)
                    xin[jj] = (DIR == 0 ? __shfl_up_sync(kFull, y[CPL - L + jj], 1) : __shfl_down_sync(kFull, y[jj], 1)) * kin;
                const float (&qq)[L + 1][CPL] = q[i & 1];
                float out[CPL];
#pragma unroll
                for (int r = 0; r < CPL; ++r) {
                    float acc = qq[0][r] * y[r];
#pragma unroll
                    for (int d = 1; d <= L; ++d) {
                        const int idx = DIR == 0 ? r - d : r + d;
                        if (DIR == 0 ? idx >= 0 : idx < CPL) acc = fmaf(qq[d][r], y[idx], acc);
                    }
#pragma unroll
                    for (int d = 1; d <= L; ++d) {
                        const int idx = DIR == 0 ? r - d : r + d;
                        if (DIR == 0 ? idx < 0 : idx >= CPL) acc = fmaf(qq[d][r], xin[DIR == 0 ? L + idx : idx - CPL], acc);
                    }
                    out[r] = acc * c;
                }
#pragma unroll
                for (int r = 0; r < CPL; ++r) y[r] = out[r];
                // vector k (the result) has frame Fn
                {
                    float* row = vec + (size_t)(DIR == 0 ? k + 1 : Gb - 1 - k) * (UP + 32);
                    tp_store<CPL>(row + c0, y);
                    reinterpret_cast<int*>(row + UP)[lane] = Fn;
                }
                F = Fn; Fn = F2; c = c_next; kin = kin_next;
            }
        }
        // stage j is consumed (its last operator went through the FMAs above): refill its slot
        __syncwarp();
        if (lane == 0 && j + NS < nst) issue(j + NS);
    }
    // Z: forward = alpha_G(U-1) (beta_G is the unit vector there), backward = beta_0(0); F is the last vector's frame
    const int zt = DIR == 0 ? U - 1 : 0;
    if (zt / CPL == lane) {
        float yz = y[0];
#pragma unroll
        for (int r = 1; r < CPL; ++r) yz = (zt % CPL == r) ? y[r] : yz;
        float* z = p.zlg + (size_t)b * 4 + DIR * 2;
        z[0] = yz > 0.0f ? log2f(yz) : -INFINITY;
        z[1] = (float)F;
    }
}

template <int CPL>
__global__ void __launch_bounds__(32) tp4_sweep_kernel(const TpParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const FbArgs& a = p.a;
    const int lane = threadIdx.x;
    const int b = blockIdx.x >> 1, dir = blockIdx.x & 1;
    tp_pdl_trigger();  // the fill kernel's CTAs may be launched (they wait for this grid's completion before reading)
    int T, U;
    if (!tp_lengths(a, b, T, U)) return;
    if (dir == 0 && lane == 0) p.status[b] = 0u;  // the fill kernel ORs into it
    const int Gb = (T + kTp4L - 1) / kTp4L;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // [NS]
    const float* ring = reinterpret_cast<const float*>(smem_raw + 128);
    if (lane == 0) {
        for (int s = 0; s < Tp4Dims<CPL>::NS; ++s) mbar_init(smem_u32(bars + s), 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (dir == 0) tp4_sweep<CPL, 0>(p, b, U, Gb, lane, bars, ring);
    else tp4_sweep<CPL, 1>(p, b, U, Gb, lane, bars, ring);
}

// =================================================================================================
// Kernel 3: group interiors and gradients.
// =================================================================================================
template <int CPL>
__global__ void __launch_bounds__(32) tp4_fill_kernel(const TpParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int L = kTp4L, R = kTp4Rows, UP = Tp4Dims<CPL>::UP, NGI = R / L;
    const FbArgs& a = p.a;
    const int lane = threadIdx.x;
    const int nchunk = (a.max_t + R - 1) / R;
    const int b = blockIdx.x / nchunk, c = blockIdx.x % nchunk;
    const int max_u = a.max_u, max_t = a.max_t;
    const int c0 = lane * CPL;
    const int t0 = c * R;
    const size_t slab = (size_t)max_t * max_u;
    float* ge = a.grad_emit + (size_t)b * slab;
    float* gs = a.grad_shift + (size_t)b * slab;
    const float zeros[CPL] = {};
    auto zero_rows = [&](int from, int to) {
        for (int t = from; t < to; ++t) {
            store_cells_cs<CPL>(ge + (size_t)t * max_u, c0, max_u, zeros);
            store_cells_cs<CPL>(gs + (size_t)t * max_u, c0, max_u, zeros);
        }
    };
    int T, U;
    const int rows_end = min(t0 + R, max_t);
    tp_pdl_trigger();  // the log-domain re-run kernel may be launched; it waits for this grid before reading status
    const bool valid = tp_lengths(a, b, T, U);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float* se = reinterpret_cast<float*>(smem_raw + 128);
    float* ss = se + R * max_u;
    if (valid && t0 < T && lane == 0) {  // the raw rows do not depend on the predecessor kernels: start their copy now
        mbar_init(smem_u32(bar), 1);
        fence_mbar_init();
        tp_issue_chunk(a, b, t0, min(R, max_t - t0), se, ss, smem_u32(bar));
    }
    __syncwarp();
    tp_pdl_wait();     // the sweep kernel has completed: boundary vectors, likelihoods and status are visible
    if (!valid) {
        zero_rows(t0, rows_end);
        if (c == 0 && lane == 0) {
            a.log_likelihood[b] = -INFINITY;
            p.status[b] = 0u;
        }
        return;
    }
    const float* z = p.zlg + (size_t)b * 4;
    const float zf_lg = z[0], zf_ex = z[1], zb_lg = z[2], zb_ex = z[3];
    const float zdiff = (zf_ex - zb_ex) + (zf_lg - zb_lg);
    const bool z_ok = (zf_lg > -1e30f) && (zb_lg > -1e30f) && fabsf(zdiff) <= kTpZTol && !p.force_fallback;
    if (!z_ok) {
        // no mass reached the end (a true -inf or an underflow) or the sweeps disagree: the log-domain kernel decides
        if (c == 0 && lane == 0) p.status[b] = p.force_fallback ? (unsigned)kTpForced : (unsigned)kTpBadZ;
#ifdef SSNT_TP_TRACE
        if (c == 0 && lane == 0) printf("tp4: b=%d bad Z: fwd %f + %f, bwd %f + %f, diff %g\n", b, zf_lg, zf_ex, zb_lg, zb_ex, zdiff);
#endif
        if (t0 < T) tp_wait(smem_u32(bar), 0, 3);  // do not leave the CTA with a bulk copy in flight
        return;
    }
    if (t0 >= T) {
        zero_rows(t0, rows_end);
        return;
    }
    if (c == 0 && lane == 0) a.log_likelihood[b] = (float)(((double)zf_lg + (double)zf_ex) * kLn2);
    // ---- boundary vectors of the chunk's groups (registers), before waiting for the rows ----
    float av[NGI][CPL], bv[NGI][CPL];
    int ea[NGI], eb[NGI];
#pragma unroll
    for (int gi = 0; gi < NGI; ++gi) {
        const int g = c * NGI + gi;
        if (g * L < T) {
            const float* arow = p.A + ((size_t)b * (p.C + 1) + g) * (UP + 32);
            const float* brow = p.Bv + ((size_t)b * (p.C + 1) + g + 1) * (UP + 32);
            tp_load<CPL>(arow + c0, av[gi]);
            tp_load<CPL>(brow + c0, bv[gi]);
            ea[gi] = reinterpret_cast<const int*>(arow + UP)[lane];
            eb[gi] = reinterpret_cast<const int*>(brow + UP)[lane];
        } else {
#pragma unroll
            for (int r = 0; r < CPL; ++r) { av[gi][r] = 0.0f; bv[gi][r] = 0.0f; }
            ea[gi] = eb[gi] = kTpDead;
        }
    }
    tp_wait(smem_u32(bar), 0, 3);
    float worst = 0.0f;
#pragma unroll
    for (int gi = 0; gi < NGI; ++gi) {
        const int g = c * NGI + gi;
        const int tg = g * L;
        if (tg >= T) {
            zero_rows(min(tg, rows_end), min(tg + L, rows_end));
            continue;
        }
        float (&A)[CPL] = av[gi];
        float (&Bt)[CPL] = bv[gi];
        // per-lane renormalisation (exact), then the frames held fixed over the group: the lane's own exponent, but at
        // most kTp4Guard below the neighbour's the mass comes from (what enters within four frames cannot overflow)
        int xa, xb;
        {
            float ma = A[0], mb = Bt[0];
#pragma unroll
            for (int r = 1; r < CPL; ++r) { ma = fmaxf(ma, A[r]); mb = fmaxf(mb, Bt[r]); }
            const int sha = ma > 0.0f ? tp4_exponent(ma) : 0, shb = mb > 0.0f ? tp4_exponent(mb) : 0;
            const float fa0 = tp_pow2(-sha), fb0 = tp_pow2(-shb);
#pragma unroll
            for (int r = 0; r < CPL; ++r) { A[r] *= fa0; Bt[r] *= fb0; }
            xa = ma > 0.0f ? ea[gi] + sha : kTpDead;
            xb = mb > 0.0f ? eb[gi] + shb : kTpDead;
        }
        int xa_l = __shfl_up_sync(kFull, xa, 1), xb_r = __shfl_down_sync(kFull, xb, 1);
        if (lane == 0) xa_l = kTpDead;
        if (lane == 31) xb_r = kTpDead;
        const int fa = max(xa, xa_l - kTp4Guard), fb = max(xb, xb_r - kTp4Guard);
        const int fa_l = __shfl_up_sync(kFull, fa, 1), fb_r = __shfl_down_sync(kFull, fb, 1);
        const float ka = lane == 0 ? 0.0f : tp_pow2(fa_l - fa);     // applied to what enters from lane-1
        const float kb = lane == 31 ? 0.0f : tp_pow2(fb_r - fb);    // applied to what enters from lane+1
        {
            const float sa0 = tp_pow2_neg(xa - fa), sb0 = tp_pow2_neg(xb - fb);
#pragma unroll
            for (int r = 0; r < CPL; ++r) { A[r] *= sa0; Bt[r] *= sb0; }
        }
        // occupancy = alpha * (e|s) * beta / Z = (a * 2^x1) * (p * 2^x2), x1 + x2 = fa + fb - log2 Z, split evenly
        float sa, sb;
        {
            const float xi = fmaxf((float)fa + (float)fb - zf_ex, -1000.0f);
            const float half = floorf(0.5f * xi);
            sa = ex2(fminf(fmaxf((xi - half) - zf_lg, -126.0f), 126.0f));
            sb = ex2(fminf(fmaxf(half, -126.0f), 126.0f));
        }
        // ---- alpha forward: the group's rows kept in registers (scaled by sa); probabilities written back ----
        float ar[L][CPL];
#pragma unroll
        for (int l = 0; l < L; ++l) {
            float e[CPL], s[CPL];
            tp_row_probs<CPL>(se, ss, gi * L + l, tg + l, T, U, max_u, c0, e, s);
            store_cells<CPL>(se + (gi * L + l) * max_u, c0, max_u, e);  // raw rows overwritten in place by the probabilities
            store_cells<CPL>(ss + (gi * L + l) * max_u, c0, max_u, s);  // (each lane re-reads only what it wrote itself)
#pragma unroll
            for (int r = 0; r < CPL; ++r) ar[l][r] = A[r] * sa;
            if (l < L - 1) {
                const float in = __shfl_up_sync(kFull, s[CPL - 1] * A[CPL - 1], 1) * ka;
#pragma unroll
                for (int r = CPL - 1; r >= 1; --r) A[r] = fmaf(e[r], A[r], s[r - 1] * A[r - 1]);
                A[0] = fmaf(e[0], A[0], in);
            }
        }
        __syncwarp();
        // ---- beta backward with the gradients fused ----
        float gsum = 0.0f;
        int nrows = 0;
#pragma unroll
        for (int l = L - 1; l >= 0; --l) {
            const int t = tg + l;
            float e[CPL], s[CPL];
            load_cells<CPL>(se + (gi * L + l) * max_u, c0, max_u, 0.0f, e);
            load_cells<CPL>(ss + (gi * L + l) * max_u, c0, max_u, 0.0f, s);
            const float bin = __shfl_down_sync(kFull, Bt[0], 1) * kb;
            float g1[CPL], g2[CPL];
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                const float p1 = e[r] * Bt[r];
                const float p2 = s[r] * (r + 1 < CPL ? Bt[r + 1] : bin);
                g1[r] = ar[l][r] * (p1 * sb);
                g2[r] = ar[l][r] * (p2 * sb);
                Bt[r] = p1 + p2;
            }
            if (t < T) {
                store_cells_cs<CPL>(ge + (size_t)t * max_u, c0, max_u, g1);
                store_cells_cs<CPL>(gs + (size_t)t * max_u, c0, max_u, g2);
#pragma unroll
                for (int r = 0; r < CPL; ++r) gsum += g1[r] + g2[r];
                ++nrows;
            } else if (t < max_t) {
                store_cells_cs<CPL>(ge + (size_t)t * max_u, c0, max_u, zeros);
                store_cells_cs<CPL>(gs + (size_t)t * max_u, c0, max_u, zeros);
            }
        }
        // every frame's occupancies sum to 1: checked per group (one reduction instead of four)
        gsum = warp_sum(gsum);
        const float dev = fabsf(gsum - (float)nrows);
        worst = (dev <= kTpRowTol * (float)nrows) ? worst : 1.0f;  // also catches NaN
#ifdef SSNT_TP_TRACE
        if (!(dev <= kTpRowTol * (float)nrows) && lane == 0) printf("tp4: b=%d g=%d group sum %g of %d (fa %d fb %d sa %g sb %g)\n", b, g, gsum, nrows, fa, fb, sa, sb);
#endif
    }
    if (worst != 0.0f && lane == 0) atomicOr(p.status + b, (unsigned)kTpBadRow);
}

}  // namespace lattice
}  // namespace ssnt
