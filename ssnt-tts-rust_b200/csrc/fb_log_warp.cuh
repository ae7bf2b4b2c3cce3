// Log-domain lattice sweep of one utterance by one cluster of two CTAs (one warp each does the
// recursion): the numerically unconditional path.  Used (a) as kernel kind 1 (fb_log_warp_kernel)
// and (b) as the in-kernel fallback of the block-floating-point kernel for utterances whose
// dynamic range it cannot hold.
//
// Structure.  Rank 0 sweeps alpha forward from frame 0, rank 1 sweeps beta backward from the
// virtual terminal frame T, concurrently.  Phase 1 stores the first half of each sweep to the
// scratch rows; after one cluster barrier both ranks compute the log-likelihood from the meeting
// row (identical arithmetic → identical bits) and phase 2 emits the gradients of the rows it
// walks, reading the partner's stored half.  Rows are prefetched by TMA bulk copies into a
// shared-memory ring (NS stages x 8 rows).
//
// Numerics.  log2 domain, fp32, with one integer-valued offset PER LANE (true value = stored +
// lane offset).  Every row each lane re-centres on its own maximum of two rows earlier by biasing
// that row's inputs (lae is shift-invariant), so the re-centring never sits on the recursion's
// dependency chain and stored magnitudes stay at the in-lane spread (a few tens) instead of
// |log-likelihood| — fp32 rounding then acts at ~1e-6 instead of ~1e-4.  Cross-lane exchanges add
// the (exact) difference of the two lanes' offsets.  -inf is the finite sentinel kNeg.
#pragma once
#include "lattice_common.cuh"

namespace ssnt {
namespace lattice {

namespace cg = cooperative_groups;

struct LogParams {
    FbArgs a;
    float* scratch;  // [B][max_t + 1][SU]; row t = alpha(t) for t < m, beta(t) for t >= m
    int SU;          // scratch row stride in floats = round_up4(max_u) + 32 (32 lane offsets)
    int NS;          // pipeline stages
    unsigned* counter;
    const unsigned* only = nullptr;  // optional [B]: run only the utterances whose word is non-zero (re-run of what a
                                     // block-float kernel flagged); the others keep their results
    unsigned* fallbacks = nullptr;   // counts the utterances re-run through `only`
};

// Returns through global memory: ll[b], gradients of rows [0, T), zeros for rows [T, max_t).
// Must be called by all 32 lanes of ONE warp in each of the two CTAs of a cluster; executes
// exactly one cluster barrier.  `bars`/`ring`: 128 B of mbarriers + NS * stage_floats floats.
template <int CPL>
__device__ void log_lattice_cta(const LogParams& p, int b, unsigned rank, int lane, int T, int U,
                                uint64_t* bars, float* ring, cg::cluster_group& cluster) {
    const FbArgs& a = p.a;
    const int max_t = a.max_t, max_u = a.max_u, SU = p.SU, NS = p.NS;
    const size_t slab = (size_t)max_t * max_u;
    // raw-logit mode (FbArgs::logits): one input tensor z and one gradient tensor; log2 sigmoid(+-z) formed per row
    const bool logits = a.logits != nullptr;
    const float* le = (logits ? a.logits : a.log_emit) + (size_t)b * slab;
    const float* ls = logits ? nullptr : a.log_shift + (size_t)b * slab;
    float* ge = (logits ? a.grad_logits : a.grad_emit) + (size_t)b * slab;
    float* gs = logits ? nullptr : a.grad_shift + (size_t)b * slab;
    float* scr = p.scratch + (size_t)b * (max_t + 1) * SU;
    const int c0 = lane * CPL;
    const int UP = SU - 32;  // where the lane offsets start in a scratch row

    const int stage_floats = kG * (2 * max_u + SU);
    const int off_e = 0, off_s = kG * max_u, off_x = 2 * kG * max_u;
    if (lane == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(smem_u32(bars + s), 1);
        fence_mbar_init();
    }
    __syncwarp();

    const int m = (T + 1) >> 1;  // alpha phase 1: rows 0..m-1; beta phase 1: rows T..m
    unsigned kg = 0;             // global stage counter (slot = kg % NS, parity = (kg / NS) & 1)

    auto issue = [&](int k, int n, int t0, int dir, bool with_x, int xoff, unsigned kbase) {
        const int j0 = k * kG;
        const int cnt = min(kG, n - j0);
        const int r0 = dir > 0 ? t0 + j0 : t0 - j0 - cnt + 1;
        const unsigned kk = kbase + (unsigned)k;
        const int slot = (int)(kk % (unsigned)NS);
        const uint32_t bar = smem_u32(bars + slot);
        float* dst = ring + (size_t)slot * stage_floats;
        const uint32_t bytes_e = (uint32_t)cnt * (uint32_t)max_u * 4u;
        const uint32_t bytes_x = with_x ? (uint32_t)cnt * (uint32_t)SU * 4u : 0u;
        mbar_expect_tx(bar, (logits ? 1u : 2u) * bytes_e + bytes_x);
        bulk_g2s(smem_u32(dst + off_e), le + (size_t)r0 * max_u, bytes_e, bar);
        if (!logits) bulk_g2s(smem_u32(dst + off_s), ls + (size_t)r0 * max_u, bytes_e, bar);
        if (with_x) bulk_g2s(smem_u32(dst + off_x), scr + (size_t)(r0 + xoff) * SU, bytes_x, bar);
    };

    // ---- state --------------------------------------------------------------------------------
    float v[CPL];      // alpha~ (rank 0) / beta~ (rank 1) of the current row, this lane's frame
    float off = 0.0f;  // lane offset (integer-valued): true = v + off
    // Lazy re-centring: the shift applied at row t is  c_t = rint(max v_{t-1}) - c_{t-1}  (deadbeat:
    // the frame follows the lane maximum of two rows ago), all of it known before row t starts.
    float m_old = 0.0f, m_new = 0.0f, c_prev = 0.0f;
    float dnb = 0.0f;  // beta: off(lane+1) - off(lane) of the CURRENT frame

    auto convert = [&](float (&x)[CPL], bool all_masked) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) x[i] = (c0 + i < U && !all_masked) ? to_log2(x[i]) : kNeg;
    };
    // One row's log2 emit / shift scores from the ring stage.  Raw-logit mode: z' = z log2 e,
    // log2 sigmoid(z) = min(z', 0) - log2(1 + 2^-|z'|), log2 sigmoid(-z) = that - z'; Zr keeps z' for the gradient.
    auto load_row = [&](const float* st, int idx, int t, float (&E)[CPL], float (&Sh)[CPL], float (&Zr)[CPL]) {
        load_cells<CPL>(st + off_e + idx * max_u, c0, max_u, 0.0f, E);
        if (logits) {
            const bool last = t == T - 1;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const float zz = fminf(fmaxf(E[i] * kLog2e, kNeg), -kNeg);
                const float sp = lg2(1.0f + ex2(-fabsf(zz)));
                Zr[i] = zz;
                E[i] = (c0 + i < U) ? fmaxf(fminf(zz, 0.0f) - sp, kNeg) : kNeg;
                Sh[i] = (c0 + i < U && !last) ? fmaxf(fminf(-zz, 0.0f) - sp, kNeg) : kNeg;
            }
        } else {
            load_cells<CPL>(st + off_s + idx * max_u, c0, max_u, 0.0f, Sh);
            convert(E, false);
            convert(Sh, t == T - 1);  // the last frame must emit
        }
    };
    auto lane_max = [&]() {
        float mx = v[0];
#pragma unroll
        for (int i = 1; i < CPL; ++i) mx = fmaxf(mx, v[i]);
        return mx > kNegTest ? rintf(mx) : 0.0f;
    };
    // alpha step with lazy re-centring: new frame = off + c_now.
    auto alpha_step = [&](const float (&E)[CPL], const float (&Sh)[CPL]) {
        const float c = m_old - c_prev;
        off += c;
        const float d = __shfl_up_sync(kFull, off, 1) - off;  // neighbour frame - my frame (new frames)
        float y[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) y[i] = v[i] + (Sh[i] - c);
        float yin = __shfl_up_sync(kFull, y[CPL - 1], 1) + d;
        if (lane == 0) yin = kNeg;
#pragma unroll
        for (int i = CPL - 1; i >= 1; --i) v[i] = lae2(v[i] + (E[i] - c), y[i - 1]);
        v[0] = lae2(v[0] + (E[0] - c), yin);
        c_prev = c;
        m_old = m_new;
        m_new = lane_max();
    };
    // beta step: v = beta~(t+1) → beta~(t); x, y are the unbiased E + beta(t+1,u), S + beta(t+1,u+1)
    // in the OLD frame; the new frame is off + c_now.
    auto beta_finish = [&](const float (&x)[CPL], const float (&y)[CPL]) {
        const float c = m_old - c_prev;
#pragma unroll
        for (int i = 0; i < CPL; ++i) v[i] = lae2(x[i] - c, y[i] - c);
        off += c;
        dnb = __shfl_down_sync(kFull, off, 1) - off;
        c_prev = c;
        m_old = m_new;
        m_new = lane_max();
    };
    auto store_state_row = [&](int t) {
        float* row = scr + (size_t)t * SU;
        store_cells<CPL>(row, c0, max_u, v);
        row[UP + lane] = off;
    };

    // =========================== phase 1 ===========================
    if (rank == 0) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) v[i] = (c0 + i == 0) ? 0.0f : kNeg;
        store_state_row(0);
    } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) v[i] = (c0 + i == U - 1) ? 0.0f : kNeg;
        store_state_row(T);  // virtual terminal row beta(T, .)
    }
    {
        const int n = rank == 0 ? (m - 1) : (T - m);
        const int t0 = rank == 0 ? 0 : T - 1;
        const int dir = rank == 0 ? 1 : -1;
        const int nst = (n + kG - 1) / kG;
        if (lane == 0)
            for (int k = 0; k < min(NS, nst); ++k) issue(k, n, t0, dir, false, 0, kg);
        for (int k = 0; k < nst; ++k) {
            const unsigned kk = kg + (unsigned)k;
            const int slot = (int)(kk % (unsigned)NS);
            mbar_wait_warp(smem_u32(bars + slot), (kk / (unsigned)NS) & 1u);  // warp-uniform: stays converged
            const float* st = ring + (size_t)slot * stage_floats;
            const int j0 = k * kG;
            const int cnt = min(kG, n - j0);
#pragma unroll
            for (int q = 0; q < kG; ++q) {
                if (q < cnt) {
                    const int t = t0 + dir * (j0 + q);
                    const int idx = dir > 0 ? q : cnt - 1 - q;
                    float E[CPL], Sh[CPL], Zr[CPL];
                    load_row(st, idx, t, E, Sh, Zr);
                    if (rank == 0) {
                        alpha_step(E, Sh);
                        store_state_row(t + 1);
                    } else {
                        float bs = __shfl_down_sync(kFull, v[0], 1) + dnb;
                        if (lane == 31) bs = kNeg;
                        float x[CPL], y[CPL];
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            const float nb = (i + 1 < CPL) ? v[i + 1] : bs;
                            x[i] = E[i] + v[i];
                            y[i] = Sh[i] + nb;
                        }
                        beta_finish(x, y);
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
    // rank 0: rows t = m-1 .. T-1, scratch row t+1 = beta(t+1); the first row only yields LL.
    // rank 1: rows t = m-1 .. 0,   scratch row t   = alpha(t);  every row emits gradients.
    float llt = 0.0f, ref = 0.0f;  // LL2 = ref + llt
    bool dead = false;
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
            mbar_wait_warp(smem_u32(bars + slot), (kk / (unsigned)NS) & 1u);  // warp-uniform: stays converged
            const float* st = ring + (size_t)slot * stage_floats;
            const int j0 = k * kG;
            const int cnt = min(kG, n - j0);
#pragma unroll
            for (int q = 0; q < kG; ++q) {
                if (q < cnt) {
                    const int t = t0 + dir * (j0 + q);
                    const int idx = dir > 0 ? q : cnt - 1 - q;
                    float E[CPL], Sh[CPL], X[CPL], Zr[CPL];
                    load_row(st, idx, t, E, Sh, Zr);
                    const float* xrow = st + off_x + idx * SU;
                    load_cells<CPL>(xrow, c0, max_u, kNeg, X);
                    const float xo = xrow[UP + lane];  // partner's lane offset of that row
                    const bool first = (j0 + q) == 0;  // t == m-1: the meeting row
                    // beta~(t+1,u) in frame fb, and its right neighbour brought into the same frame
                    float bn[CPL], x[CPL], y[CPL];
                    float fb, fa, bsh;
                    if (rank == 0) {
#pragma unroll
                        for (int i = 0; i < CPL; ++i) bn[i] = X[i];
                        fb = xo;
                        fa = off;
                        bsh = __shfl_down_sync(kFull, X[0], 1) + (__shfl_down_sync(kFull, xo, 1) - xo);
                    } else {
#pragma unroll
                        for (int i = 0; i < CPL; ++i) bn[i] = v[i];
                        fb = off;
                        fa = xo;
                        bsh = __shfl_down_sync(kFull, v[0], 1) + dnb;
                    }
                    if (lane == 31) bsh = kNeg;
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        const float nb = (i + 1 < CPL) ? bn[i + 1] : bsh;
                        x[i] = E[i] + bn[i];
                        y[i] = Sh[i] + nb;
                    }
                    const float frame = fa + fb;  // exact: integer-valued floats
                    if (first) {
                        // LL = LSE_u( alpha(m-1,u) + beta(m-1,u) ); same operands in both CTAs.
                        float term[CPL];
                        float mx = kNeg;
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            const float av = rank == 0 ? v[i] : X[i];
                            term[i] = fmaxf(av + lae2(x[i], y[i]), kNeg);
                            mx = fmaxf(mx, term[i]);
                        }
                        const float gmx = warp_max(mx > kNegTest ? mx + frame : kNeg);
                        dead = !(gmx > kNegTest);
                        ref = dead ? 0.0f : rintf(gmx);
                        const float rel = frame - ref;
                        float sum = 0.0f;
#pragma unroll
                        for (int i = 0; i < CPL; ++i) sum += ex2(term[i] + rel);
                        sum = warp_sum(sum);
                        llt = lg2(sum);
                        if (rank == 0 && lane == 0) {
                            const double ll2 = (double)llt + (double)ref;
                            a.log_likelihood[b] = dead ? -INFINITY : (float)(ll2 * kLn2);
                        }
                    }
                    if (rank == 1 || !first) {
                        const float kt = (frame - ref) - llt;
                        float g1[CPL], g2[CPL];
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            const float av = rank == 0 ? v[i] : X[i];
                            g1[i] = dead ? 0.0f : ex2((av + x[i]) + kt);
                            g2[i] = dead ? 0.0f : ex2((av + y[i]) + kt);
                        }
                        if (logits) {  // dLL/dz = occupancy(emit) sigmoid(-z) - occupancy(shift) sigmoid(z)
#pragma unroll
                            for (int i = 0; i < CPL; ++i) {
                                const float sp = lg2(1.0f + ex2(-fabsf(Zr[i])));
                                g1[i] = g1[i] * ex2(fminf(-Zr[i], 0.0f) - sp) - g2[i] * ex2(fminf(Zr[i], 0.0f) - sp);
                            }
                            store_cells_cs<CPL>(ge + (size_t)t * max_u, c0, max_u, g1);
                        } else {
                            store_cells_cs<CPL>(ge + (size_t)t * max_u, c0, max_u, g1);
                            store_cells_cs<CPL>(gs + (size_t)t * max_u, c0, max_u, g2);
                        }
                    }
                    if (rank == 0) {
                        if (t < T - 1) alpha_step(E, Sh);
                    } else {
                        beta_finish(x, y);
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
        const float zeros[CPL] = {};
        float* g = rank == 0 ? ge : gs;
        if (g) for (int t = T; t < max_t; ++t) store_cells_cs<CPL>(g + (size_t)t * max_u, c0, max_u, zeros);
    }
}

}  // namespace lattice
}  // namespace ssnt
