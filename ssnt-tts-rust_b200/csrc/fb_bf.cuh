// Block-floating-point lattice kernel — the hot path (kernel kind 2).
//
// Why.  In the log domain every lattice cell costs an EX2 + LG2 pair on the recursion's dependency
// chain and ~170 instructions per row in one warp (profiles/r01a_*).  Here the recursion runs on
// PROBABILITIES: alpha'(u) = alpha(u)·e(u) + alpha(u-1)·s(u-1) is one FMUL + one FFMA per cell, and
// the only cross-lane step is one shuffle per row whose latency hides behind the other cells of the
// lane.  Range is handled by block floating point: each lane (CPL consecutive tokens) carries one
// shared power-of-two exponent, re-normalised every 8 rows by exact power-of-two scaling, so the
// arithmetic error is fp32's relative 6e-8 per operation (measured ~1e-6 on the gradients, two
// orders better than an fp32 log-domain recursion).  Utterances whose dynamic range this cannot
// hold (all mass lost, non-finite values, or the three independent likelihood estimates
// disagreeing) are re-run by the same cluster in the log domain (fb_log_warp.cuh) — no host
// round trip, no CPU path.
//
// Organisation.  One cluster of two CTAs per utterance (rank 0: alpha from frame 0, rank 1: beta
// from the virtual terminal frame T; they meet in the middle exactly as in fb_log_warp.cuh).  Each
// CTA is warp-specialised, 5 warps:
//   warp 0   recursion ("chain"): the only serial work, ~15 instructions per row.
//   warp 1-3 helpers: (prep) convert the TMA-landed log-prob rows to probabilities in place, with
//            the length masks; (post, phase 2) combine the chain's state row with the partner's
//            stored row into gradients and write them with 128-bit streaming stores; the first
//            post row also produces the log-likelihood.
//   warp 4   producer: one lane issues the TMA bulk copies (log_emit / log_shift rows and, in
//            phase 2, the partner's scratch rows) NS stages x 8 rows ahead.
// Hand-offs are mbarriers per ring slot: raw_full (TMA → prep), prep_full (prep → chain),
// state_full (chain → post), slot_free (last reader → producer).
#pragma once
#include "fb_log_warp.cuh"

namespace ssnt {
namespace lattice {

constexpr int kTarget = 24;       // lane maximum is scaled to ~2^kTarget at every re-normalisation
constexpr int kSlack = 32;        // a lane's frame may sit this far below its feeding neighbour's edge
constexpr int kNoMass = -100000;  // exponent key of an all-zero lane
constexpr int kBfThreads = 160;
constexpr int kHelpers = 3;
constexpr int kLookahead = 2;     // stages the helpers' prep runs ahead of their post

struct BfParams {
    FbArgs a;
    float* scratch;    // [B][max_t + 1][SU]: per row CPL·32 values + 32 lane exponents (int bits)
    unsigned* status;  // [B] nonzero → utterance was re-run in the log domain
    int SU, NS;
    int force_fallback;
    unsigned* counter;
};

enum BfStatus : unsigned { kBfNoMass = 1u, kBfNonFinite = 2u, kBfInconsistent = 4u, kBfForced = 8u };

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ int ilogb_pos(float x) {  // floor(log2 x) for normal x > 0
    return ((__float_as_int(x) >> 23) & 0xff) - 127;
}
__device__ __forceinline__ float pow2i(int k) {  // 2^k, k in [-126, 127]
    return __int_as_float((k + 127) << 23);
}
// x * 2^k for |k| up to 252, exact unless the result leaves the fp32 range.
__device__ __forceinline__ float scale_pow2(float x, int k) {
    k = max(-252, min(252, k));
    const int k1 = k >> 1;
    return (x * pow2i(k1)) * pow2i(k - k1);
}

// Re-normalisation of one lane at a stage boundary.  DIR=+1: alpha (fed by lane-1's last cell),
// DIR=-1: beta (fed by lane+1's first cell).  Returns the factor g = 2^(ex_feeder - ex_mine) that
// brings the feeder's edge value into this lane's frame (0 for the lane without a feeder).
template <int CPL, int DIR>
__device__ __forceinline__ float renorm(float (&v)[CPL], int& ex, int lane) {
    float mx = v[0];
#pragma unroll
    for (int i = 1; i < CPL; ++i) mx = fmaxf(mx, v[i]);
    const int own = mx > 0.0f ? ex + ilogb_pos(mx) - kTarget : kNoMass;
    const float edge = DIR > 0 ? v[CPL - 1] : v[0];
    const int amag = edge > 0.0f ? ex + ilogb_pos(edge) : kNoMass;
    int nb = DIR > 0 ? __shfl_up_sync(kFull, amag, 1) : __shfl_down_sync(kFull, amag, 1);
    if ((DIR > 0 && lane == 0) || (DIR < 0 && lane == 31)) nb = kNoMass;
    int nw = max(own, nb - kTarget - kSlack);
    if (nw <= kNoMass / 2) nw = ex;  // nothing here and nothing arriving: keep the frame
    const int shift = ex - nw;
#pragma unroll
    for (int i = 0; i < CPL; ++i) v[i] = scale_pow2(v[i], shift);
    ex = nw;
    const int fe = DIR > 0 ? __shfl_up_sync(kFull, ex, 1) : __shfl_down_sync(kFull, ex, 1);
    float g = pow2i(max(-126, min(126, fe - ex)));
    if ((DIR > 0 && lane == 0) || (DIR < 0 && lane == 31)) g = 0.0f;
    return g;
}

template <int CPL>
__device__ void bf_lattice_cta(const BfParams& p, int b, unsigned rank, int T, int U,
                               unsigned char* smem_raw, cg::cluster_group& cluster) {
    const FbArgs& a = p.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int max_t = a.max_t, max_u = a.max_u, SU = p.SU, NS = p.NS;
    const int UP = SU - 32;
    const size_t slab = (size_t)max_t * max_u;
    const float* le = a.log_emit + (size_t)b * slab;
    const float* ls = a.log_shift + (size_t)b * slab;
    float* ge = a.grad_emit + (size_t)b * slab;
    float* gs = a.grad_shift + (size_t)b * slab;
    float* scr = p.scratch + (size_t)b * (max_t + 1) * SU;
    const int c0 = lane * CPL;

    // ---- shared memory carve-up -------------------------------------------------------------------
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // 4 x NS barriers (NS <= 8)
    uint64_t* raw_full = bars;
    uint64_t* prep_full = bars + 8;
    uint64_t* state_full = bars + 16;
    uint64_t* slot_free = bars + 24;
    float* llinfo = reinterpret_cast<float*>(smem_raw + 256);  // [0] M (int bits) [1] 1/sum [2] dead
    float* ring = reinterpret_cast<float*>(smem_raw + 384);
    // per slot: e[8][max_u] | s[8][max_u] | x[8][SU] | state[8][max_u] | state_exp[32]
    const int off_e = 0, off_s = kG * max_u, off_x = 2 * kG * max_u, off_v = off_x + kG * SU,
              off_ve = off_v + kG * max_u;
    const int stage_floats = off_ve + 32;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(smem_u32(raw_full + s), 1);
            mbar_init(smem_u32(prep_full + s), kHelpers);
            mbar_init(smem_u32(state_full + s), 1);
            mbar_init(smem_u32(slot_free + s), kHelpers);
        }
        fence_mbar_init();
        if (rank == 0) p.status[b] = p.force_fallback ? (unsigned)kBfForced : 0u;
    }
    __syncthreads();

    const int m = (T + 1) >> 1;
    const int dir = rank == 0 ? 1 : -1;

    // Sweep geometry of a phase: n rows, row(j) = t0 + dir*j; stage k holds j in [8k, 8k+cnt).
    struct Phase { int n, t0, nst, xoff; bool with_x; };
    Phase ph[2];
    ph[0].n = rank == 0 ? (m - 1) : (T - m);
    ph[0].t0 = rank == 0 ? 0 : T - 1;
    ph[0].with_x = false; ph[0].xoff = 0;
    ph[1].n = rank == 0 ? (T - m + 1) : m;
    ph[1].t0 = m - 1;
    ph[1].with_x = true; ph[1].xoff = rank == 0 ? 1 : 0;
    ph[0].nst = (ph[0].n + kG - 1) / kG;
    ph[1].nst = (ph[1].n + kG - 1) / kG;

    auto slot_of = [&](unsigned kk) { return (int)(kk % (unsigned)NS); };
    auto use_of = [&](unsigned kk) { return kk / (unsigned)NS; };
    auto slot_ptr = [&](int slot) { return ring + (size_t)slot * stage_floats; };

    // =================================================================================================
    if (warp == 4) {
        // ------------------------------- producer -------------------------------
        unsigned kg = 0;
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == 1) {
                cluster.sync();
                fence_proxy_async();
            }
            const Phase& P = ph[phase];
            if (lane == 0) {
                for (int k = 0; k < P.nst; ++k) {
                    const unsigned kk = kg + (unsigned)k;
                    const int slot = slot_of(kk);
                    const unsigned use = use_of(kk);
                    if (use > 0) mbar_wait(smem_u32(slot_free + slot), (use - 1) & 1u);
                    fence_proxy_async();
                    const int j0 = k * kG;
                    const int cnt = min(kG, P.n - j0);
                    const int r0 = dir > 0 ? P.t0 + j0 : P.t0 - j0 - cnt + 1;
                    const uint32_t bar = smem_u32(raw_full + slot);
                    float* dst = slot_ptr(slot);
                    const uint32_t bytes_e = (uint32_t)cnt * (uint32_t)max_u * 4u;
                    const uint32_t bytes_x = P.with_x ? (uint32_t)cnt * (uint32_t)SU * 4u : 0u;
                    mbar_expect_tx(bar, 2u * bytes_e + bytes_x);
                    bulk_g2s(smem_u32(dst + off_e), le + (size_t)r0 * max_u, bytes_e, bar);
                    bulk_g2s(smem_u32(dst + off_s), ls + (size_t)r0 * max_u, bytes_e, bar);
                    if (P.with_x) bulk_g2s(smem_u32(dst + off_x), scr + (size_t)(r0 + P.xoff) * SU, bytes_x, bar);
                }
            }
            __syncwarp();
            kg += (unsigned)P.nst;
        }
    } else if (warp == 0) {
        // ------------------------------- chain -------------------------------
        float v[CPL];
        int ex = 0;
#pragma unroll
        for (int i = 0; i < CPL; ++i) v[i] = 0.0f;
        if (rank == 0) {
            if (lane == 0) v[0] = 1.0f;  // alpha(0,0) = 1
        } else {
#pragma unroll
            for (int i = 0; i < CPL; ++i)
                if (c0 + i == U - 1) v[i] = 1.0f;  // virtual terminal row beta(T, U-1) = 1
        }
        auto store_scratch = [&](int row) {
            float* r = scr + (size_t)row * SU;
            store_cells<CPL>(r, c0, max_u, v);
            reinterpret_cast<int*>(r)[UP + lane] = ex;
        };
        unsigned kg = 0;
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == 1) {
                // hand the last phase-1 state over, then meet the partner
                store_scratch(rank == 0 ? m - 1 : m);
                __threadfence();
                fence_proxy_async();
                cluster.sync();
            }
            const Phase& P = ph[phase];
            for (int k = 0; k < P.nst; ++k) {
                const unsigned kk = kg + (unsigned)k;
                const int slot = slot_of(kk);
                mbar_wait(smem_u32(prep_full + slot), use_of(kk) & 1u);
                float* sp = slot_ptr(slot);
                const int j0 = k * kG;
                const int cnt = min(kG, P.n - j0);
                float g;
                if (rank == 0) g = renorm<CPL, 1>(v, ex, lane);
                else g = renorm<CPL, -1>(v, ex, lane);
                if (phase == 1) reinterpret_cast<int*>(sp + off_ve)[lane] = ex;
#pragma unroll
                for (int q = 0; q < kG; ++q) {
                    if (q < cnt) {
                        const int t = P.t0 + dir * (j0 + q);
                        const int idx = dir > 0 ? q : cnt - 1 - q;
                        float E[CPL], S[CPL];
                        load_cells<CPL>(sp + off_e + idx * max_u, c0, max_u, 0.0f, E);
                        load_cells<CPL>(sp + off_s + idx * max_u, c0, max_u, 0.0f, S);
                        // the state BEFORE the step belongs to row t (alpha(t) / beta(t+1))
                        if (phase == 0) store_scratch(rank == 0 ? t : t + 1);
                        else store_cells<CPL>(sp + off_v + q * max_u, c0, max_u, v);
                        if (rank == 0) {
                            float bsh[CPL];
#pragma unroll
                            for (int i = 0; i < CPL; ++i) bsh[i] = v[i] * S[i];
                            const float in = __shfl_up_sync(kFull, bsh[CPL - 1], 1);
#pragma unroll
                            for (int i = CPL - 1; i >= 1; --i) v[i] = fmaf(v[i], E[i], bsh[i - 1]);
                            v[0] = fmaf(in, g, v[0] * E[0]);
                        } else {
                            const float in = __shfl_down_sync(kFull, v[0], 1) * g;
#pragma unroll
                            for (int i = 0; i < CPL; ++i) {
                                const float nb = (i + 1 < CPL) ? v[i + 1] : in;
                                v[i] = fmaf(E[i], v[i], S[i] * nb);
                            }
                        }
                        (void)t;
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(smem_u32(state_full + slot));
                    if (phase == 0) mbar_arrive_n(smem_u32(slot_free + slot), kHelpers);
                }
            }
            kg += (unsigned)P.nst;
        }
    } else {
        // ------------------------------- helpers -------------------------------
        const int h = warp - 1;
        unsigned kg = 0;
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == 1) cluster.sync();
            const Phase& P = ph[phase];
            float f_inv_sum = 0.0f;
            int f_M = 0;
            bool f_dead = false;
            for (int it = 0; it < P.nst + (phase == 1 ? kLookahead : 0); ++it) {
                // ---- prep(it): log-probs → probabilities, in place ----
                if (it < P.nst) {
                    const int k = it;
                    const unsigned kk = kg + (unsigned)k;
                    const int slot = slot_of(kk);
                    mbar_wait(smem_u32(raw_full + slot), use_of(kk) & 1u);
                    float* sp = slot_ptr(slot);
                    const int j0 = k * kG;
                    const int cnt = min(kG, P.n - j0);
                    for (int q = (h + 3 - (k % 3)) % 3; q < cnt; q += kHelpers) {
                        const int t = P.t0 + dir * (j0 + q);
                        const int idx = dir > 0 ? q : cnt - 1 - q;
                        float E[CPL], S[CPL];
                        float* er = sp + off_e + idx * max_u;
                        float* sr = sp + off_s + idx * max_u;
                        load_cells<CPL>(er, c0, max_u, 0.0f, E);
                        load_cells<CPL>(sr, c0, max_u, 0.0f, S);
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            E[i] = (c0 + i < U) ? ex2(E[i] * kLog2e) : 0.0f;
                            S[i] = (c0 + i < U - 1 && t != T - 1) ? ex2(S[i] * kLog2e) : 0.0f;
                        }
                        store_cells<CPL>(er, c0, max_u, E);
                        store_cells<CPL>(sr, c0, max_u, S);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(prep_full + slot));
                }
                // ---- post(it - lookahead): gradients (phase 2 only) ----
                if (phase == 1 && it >= kLookahead) {
                    const int k = it - kLookahead;
                    const unsigned kk = kg + (unsigned)k;
                    const int slot = slot_of(kk);
                    mbar_wait(smem_u32(state_full + slot), use_of(kk) & 1u);
                    float* sp = slot_ptr(slot);
                    const int j0 = k * kG;
                    const int cnt = min(kG, P.n - j0);
                    const int ex_state = reinterpret_cast<const int*>(sp + off_ve)[lane];
                    if (k == 0 && h != 0) {  // wait for helper 0's log-likelihood
                        named_bar_sync(1, 32 * kHelpers);
                        f_M = __float_as_int(llinfo[0]);
                        f_inv_sum = llinfo[1];
                        f_dead = llinfo[2] != 0.0f;
                    }
                    for (int q = (h + 3 - (k % 3)) % 3; q < cnt; q += kHelpers) {
                        const int j = j0 + q;
                        const int t = P.t0 + dir * j;
                        const int idx = dir > 0 ? q : cnt - 1 - q;
                        float E[CPL], S[CPL], VA[CPL], VB[CPL];
                        load_cells<CPL>(sp + off_e + idx * max_u, c0, max_u, 0.0f, E);
                        load_cells<CPL>(sp + off_s + idx * max_u, c0, max_u, 0.0f, S);
                        const float* xrow = sp + off_x + idx * SU;
                        const int ex_x = reinterpret_cast<const int*>(xrow)[UP + lane];
                        int exA, exB;
                        if (rank == 0) {
                            load_cells<CPL>(sp + off_v + q * max_u, c0, max_u, 0.0f, VA);
                            load_cells<CPL>(xrow, c0, max_u, 0.0f, VB);
                            exA = ex_state; exB = ex_x;
                        } else {
                            load_cells<CPL>(xrow, c0, max_u, 0.0f, VA);
                            load_cells<CPL>(sp + off_v + q * max_u, c0, max_u, 0.0f, VB);
                            exA = ex_x; exB = ex_state;
                        }
                        // beta(t+1, u+1): in-lane neighbour, or lane+1's first cell re-framed
                        const int exBn = __shfl_down_sync(kFull, exB, 1);
                        float vbn_edge = scale_pow2(__shfl_down_sync(kFull, VB[0], 1), max(-252, min(252, exBn - exB)));
                        if (lane == 31) vbn_edge = 0.0f;
                        float pe[CPL], ps[CPL];  // e·beta(t+1,u), s·beta(t+1,u+1)
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            const float nb = (i + 1 < CPL) ? VB[i + 1] : vbn_edge;
                            pe[i] = E[i] * VB[i];
                            ps[i] = S[i] * nb;
                        }
                        const int EL = exA + exB;
                        if (j == 0) {
                            // log-likelihood from the meeting row: Z = sum_u alpha(m-1,u)·beta(m-1,u)
                            float w = 0.0f;
#pragma unroll
                            for (int i = 0; i < CPL; ++i) w += VA[i] * (pe[i] + ps[i]);
                            const bool finite = w == w && w < 3.0e38f;
                            int key = (finite && w > 0.0f) ? EL + ilogb_pos(w) : kNoMass;
                            int M = key;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) M = max(M, __shfl_xor_sync(kFull, M, o));
                            float part = (finite && w > 0.0f) ? scale_pow2(w, EL - M) : 0.0f;
                            const float sum = warp_sum(part);
                            const unsigned bad = __ballot_sync(kFull, !finite);
                            unsigned st = 0;
                            if (bad) st |= kBfNonFinite;
                            if (M <= kNoMass / 2 || !(sum > 0.0f)) st |= kBfNoMass;
                            f_M = M;
                            f_inv_sum = st ? 0.0f : 1.0f / sum;
                            f_dead = st != 0;
                            if (lane == 0) {
                                llinfo[0] = __int_as_float(M);
                                llinfo[1] = f_inv_sum;
                                llinfo[2] = f_dead ? 1.0f : 0.0f;
                                if (st) atomicOr(p.status + b, st);
                                if (rank == 0) {
                                    const double ll2 = (double)lg2(sum) + (double)M;
                                    a.log_likelihood[b] = st ? -INFINITY : (float)(ll2 * kLn2);
                                }
                            }
                            named_bar_sync(1, 32 * kHelpers);
                        }
                        if (rank == 1 || j > 0) {
                            // gamma = alpha·p·beta / Z, exponents split over two factors
                            const int kf = max(-252, min(252, EL - f_M));
                            const int k1 = kf >> 1;
                            const float fa = pow2i(max(-126, k1));
                            const float fb = pow2i(max(-126, kf - k1)) * f_inv_sum;
                            float g1[CPL], g2[CPL];
#pragma unroll
                            for (int i = 0; i < CPL; ++i) {
                                const float va = VA[i] * fa;
                                g1[i] = f_dead ? 0.0f : va * (pe[i] * fb);
                                g2[i] = f_dead ? 0.0f : va * (ps[i] * fb);
                            }
                            store_cells_cs<CPL>(ge + (size_t)t * max_u, c0, max_u, g1);
                            store_cells_cs<CPL>(gs + (size_t)t * max_u, c0, max_u, g2);
                            // consistency: occupancy of the terminal cell (alpha side) / of frame 0
                            // (beta side) must be 1 — these are independent likelihood estimates.
                            if (!f_dead) {
                                bool bad = false;
                                if (rank == 0 && t == T - 1) {
#pragma unroll
                                    for (int i = 0; i < CPL; ++i)
                                        if (c0 + i == U - 1) bad = !(fabsf(g1[i] - 1.0f) < 2e-4f);
                                }
                                if (rank == 1 && t == 0 && lane == 0) bad = !(fabsf(g1[0] + g2[0] - 1.0f) < 2e-4f);
                                if (bad) atomicOr(p.status + b, (unsigned)kBfInconsistent);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(slot_free + slot));
                }
            }
            kg += (unsigned)P.nst;
        }
    }
}

}  // namespace lattice
}  // namespace ssnt
