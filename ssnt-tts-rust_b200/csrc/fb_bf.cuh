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
// CTA is warp-specialised, 8 warps:
//   warp 0   recursion ("chain"): the only serial work.  It has SM sub-partition 0 to itself
//            (the producer that shares it sleeps between polls), preloads the 8 rows of a stage
//            into registers and runs them as straight-line FMUL/FFMA code with one shuffle per row.
//   warps 1-3, 5-7  helpers (two per remaining sub-partition): (prep) convert the TMA-landed
//            log-prob rows to probabilities in place, with the length masks; (post, phase 2)
//            combine the chain's state row with the partner's stored row into gradients and write
//            them with 128-bit streaming stores; the first post row also produces the
//            log-likelihood.
//   warp 4   producer: one lane issues the TMA bulk copies (log_emit / log_shift rows and, in
//            phase 2, the partner's scratch rows) NS stages x 8 rows ahead, and L2 prefetches
//            pf_rows ahead of those.
// Hand-offs are mbarriers per ring slot: raw_full (TMA → prep), prep_full (prep → chain),
// state_full (chain → post), slot_free (last reader → producer).
#pragma once
#include "fb_log_warp.cuh"

namespace ssnt {
namespace lattice {

constexpr int kTarget = 24;       // lane maximum is scaled to ~2^kTarget at every re-normalisation
constexpr int kSlack = 0;         // a lane's frame never sits below its feeding neighbour's edge: what arrives lands at <= 2^kTarget
                                  // when decided, leaving ~100 bits for its growth until the next decision
constexpr float kBfConsistency = 2e-5f;  // the three likelihood estimates must agree this well (typ. 1e-6)
constexpr int kNoMass = -100000;  // exponent key of an all-zero lane
constexpr int kBfHeaderBytes = 768;  // mbarriers (128) | flags (256) | llinfo | log-domain re-run barriers at 576 (128) | pad
constexpr int kBfThreads = 256;   // warp 0 chain | warp 4 producer | warps 1-3,5-7 helpers
constexpr int kHelpers = 6;
constexpr int kPairs = 3;        // helper pairs; pair p owns stages p, p+3, ...
constexpr int kHalf = 4;         // rows of a stage per helper warp
constexpr int kLookahead = 2;     // stages the helpers' prep runs ahead of their post

struct BfParams {
    FbArgs a;
    float* scratch;    // [B][max_t + 1][SU]: per row CPL·32 values + 32 lane exponents (int bits)
    unsigned* status;  // [B] nonzero → utterance was re-run in the log domain
    unsigned* fallbacks;  // cumulative count of such utterances (monitoring)
    int SU, NS;
    int pf_rows;       // L2 prefetch distance of the producer, in lattice rows
    int pf_sleep_ns;   // producer's poll interval while the ring is full
    int force_fallback;
    int debug_skip;    // profiling aid: 1 = helpers skip prep work, 2 = skip post work, 3 = both (results are wrong)
    unsigned* counter;
    long long* stats;  // optional [2B][8 warps][16] cycle counters (profiling aid), or null
};

enum BfStatus : unsigned { kBfNoMass = 1u, kBfNonFinite = 2u, kBfInconsistent = 4u, kBfForced = 8u };

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Arrival without release ordering: for hand-offs that only say "I am done READING this slot"
// (a release would first drain the warp's outstanding global stores).
__device__ __forceinline__ void mbar_arrive_relaxed_n(uint32_t bar, uint32_t n) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
// Pull a contiguous range of global memory into L2 ahead of the TMA loads that will read it.
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// Polling wait for a thread that has slack (the producer): sleeps between probes so that it does
// not take issue slots from the recursion warp on the same SM sub-partition.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, unsigned sleep_ns = 64) {
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        __nanosleep(sleep_ns);
    }
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

// Conversion of NR consecutive (in sweep order) rows of a stage from log-probabilities to
// probabilities, in place, with the length masks: tokens >= U have e = s = 0, the last token and
// the last frame cannot shift.  All loads first, then the EX2s, then the stores (ILP across rows).
template <int CPL, int NR, bool FULL = false>
__device__ __forceinline__ void prep_rows(float* e0, float* s0, int in_stride, const int t_first,
                                          const int dir, const int T, const bool (&me)[CPL],
                                          const bool (&ms)[CPL], const int c0, int max_u) {
    if (FULL) {  // max_u == 32*CPL: strides become immediates and the column bounds checks fold away
        max_u = 32 * CPL;
        in_stride = dir > 0 ? 32 * CPL : -32 * CPL;
    }
    float E[NR][CPL], S[NR][CPL];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        load_cells<CPL>(e0 + r * in_stride, c0, max_u, 0.0f, E[r]);
        load_cells<CPL>(s0 + r * in_stride, c0, max_u, 0.0f, S[r]);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const bool not_last = (t_first + dir * r) != T - 1;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            E[r][i] = me[i] ? ex2(E[r][i] * kLog2e) : 0.0f;
            S[r][i] = (ms[i] && not_last) ? ex2(S[r][i] * kLog2e) : 0.0f;
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        store_cells<CPL>(e0 + r * in_stride, c0, max_u, E[r]);
        store_cells<CPL>(s0 + r * in_stride, c0, max_u, S[r]);
    }
}

struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};

// ---- skewed recursion -----------------------------------------------------------------------------
// A warp issues in order, so a row written as "shuffle the edge cell, then use it" stalls for the
// whole shuffle latency every row (measured: 52 cycles/row for 14 instructions).  Here the cell
// that CONSUMES the neighbour lane's value lags one row behind the lane's other cells and uses a
// shuffle that was issued two rows earlier, so nothing waits:
//   RANK 0 (alpha): a[0] = alpha(c0)[s-1], a[i>=1] = alpha(c0+i)[s] before the step of row s;
//   RANK 1 (beta):  a[CPL-1] lags, the others lead (mirror image, sweep order instead of t).
// Step s: lag cell  <- lag*Ec + inA*Qc     (Ec, Qc: its emit prob / feeding shift prob x frame factor
//                                          of the PREVIOUS row, carried in registers)
//         row s of the state is now complete and is written out,
//         the other cells advance to row s+1, the new edge cell is shuffled (→ inB, used at s+2).
template <int CPL>
struct ChainState {
    float a[CPL];
    float inA, inB;  // pending shuffle results: raw neighbour values in the NEIGHBOUR's frame
    float Ec, Pc;    // carry of the lag cell: emit probability and feeding shift probability
    __device__ __forceinline__ void init(int rank, int lane, int U) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) a[i] = 0.0f;
        if (rank == 0) {
            if (lane == 0) a[0] = 1.0f;  // alpha(0,0) = 1
        } else {
#pragma unroll
            for (int i = 0; i < CPL; ++i)
                if (lane * CPL + i == U - 1) a[i] = 1.0f;  // virtual terminal row beta(T, U-1) = 1
        }
        // the state is one complete row: the next step must not advance the lag cell (identity
        // carry), and the step after it needs the feeder's edge cell of THIS row
        inA = 0.0f;
        inB = rank == 0 ? __shfl_up_sync(kFull, a[CPL - 1], 1) : __shfl_down_sync(kFull, a[0], 1);
        Ec = 1.0f; Pc = 0.0f;
    }
    // Brings the lag cell up to the row of the others (the state becomes one complete row) and
    // arms the identity carry so the next step does not advance it twice.
    __device__ __forceinline__ void flush(int rank, float g) {
        constexpr int L = CPL - 1;
        if (rank == 0) a[0] = fmaf(inA, Pc * g, a[0] * Ec);
        else a[L] = fmaf(inA, Pc * g, a[L] * Ec);
        Ec = 1.0f; Pc = 0.0f;
    }
};

// One step.  E,S: this lane's probabilities of row s; P: alpha only, s(row s, c0-1).
// out = the complete state row s (what the old code stored "before the step").
template <int CPL, int RANK>
__device__ __forceinline__ void skew_step(ChainState<CPL>& cs, const float (&E)[CPL], const float (&S)[CPL],
                                          const float P, const float g, float (&out)[CPL]) {
    static_assert(CPL >= 2, "the skewed recursion needs two cells per lane");
    constexpr int L = CPL - 1;
    const float Qc = cs.Pc * g;
    if (RANK == 0) {
        float last_new;
        if (CPL > 2) last_new = fmaf(cs.a[L], E[L], cs.a[L - 1] * S[L - 1]);
        // (inA, Qc, a0*Ec) and not (a0, Ec, inA*Qc): the product inA*Qc would be hoisted to right behind the
        // shuffle that produces inA and stall on it; this form is tied to the lag cell's previous value
        const float a0n = fmaf(cs.inA, Qc, cs.a[0] * cs.Ec);
        out[0] = a0n;
#pragma unroll
        for (int i = 1; i < CPL; ++i) out[i] = cs.a[i];
        if (CPL == 2) last_new = fmaf(cs.a[1], E[1], a0n * S[0]);
        const float sh = __shfl_up_sync(kFull, last_new, 1);
#pragma unroll
        for (int i = L - 1; i >= 2; --i) cs.a[i] = fmaf(cs.a[i], E[i], cs.a[i - 1] * S[i - 1]);
        if (CPL > 2) cs.a[1] = fmaf(cs.a[1], E[1], a0n * S[0]);
        cs.a[L] = last_new;
        cs.a[0] = a0n;
        cs.inA = cs.inB;
        cs.inB = sh;
        cs.Ec = E[0];
        cs.Pc = P;
    } else {
        float first_new;
        if (CPL > 2) first_new = fmaf(E[0], cs.a[0], S[0] * cs.a[1]);
        const float aLn = fmaf(cs.inA, Qc, cs.a[L] * cs.Ec);
        out[L] = aLn;
#pragma unroll
        for (int i = 0; i < L; ++i) out[i] = cs.a[i];
        if (CPL == 2) first_new = fmaf(E[0], cs.a[0], S[0] * aLn);
        const float sh = __shfl_down_sync(kFull, first_new, 1);
#pragma unroll
        for (int i = 1; i < L - 1; ++i) cs.a[i] = fmaf(E[i], cs.a[i], S[i] * cs.a[i + 1]);
        if (CPL > 2) cs.a[L - 1] = fmaf(E[L - 1], cs.a[L - 1], S[L - 1] * aLn);
        cs.a[0] = first_new;
        cs.a[L] = aLn;
        cs.inA = cs.inB;
        cs.inB = sh;
        cs.Ec = E[L];
        cs.Pc = S[L];
    }
}

// NR rows (one or two full stages, max_u == 32*CPL) as straight-line code.  The rows' probabilities
// are pulled into registers a chunk ahead.  e0/s0: first row IN SWEEP ORDER of the stage (rank 1 walks
// memory backwards); rows [8, 16) of a 16-row round come from e1/s1.  State rows go to st0 + q*stride
// (shared, stride max_u) or to the global scratch (stride +-SU, lane exponents behind the row).
// Hooks: h1 after row NR-3, h2 after row NR-2 (the two halves of the re-normalisation decision).
// FWD: both directions walk their stage rows forwards in memory (stages stored in sweep order).
template <int CPL, int RANK, bool TO_SMEM, int NR, typename H1, typename H2, bool FWD = false>
__device__ __forceinline__ void chain_round_skew(ChainState<CPL>& cs, const float g, const float* e0, const float* s0,
                                                 float* st0, const int ex, const int lane, H1 h1, H2 h2,
                                                 const float* e1 = nullptr, const float* s1 = nullptr,
                                                 float* st1 = nullptr) {
    constexpr int max_u = 32 * CPL, SU = max_u + 32;
    constexpr int istr = (RANK == 0 || FWD) ? max_u : -max_u;
    constexpr int sstr = TO_SMEM ? max_u : (RANK == 0 ? SU : -SU);
    constexpr int CH = 4;  // rows per chunk (CPL = 8: 2 x 4 x 17 = 136 registers of operands, fine at 255)
    constexpr int NC = NR / CH;
    const int c0 = lane * CPL;
    const int pc = c0 > 0 ? c0 - 1 : 0;
    float E[2][CH][CPL], S[2][CH][CPL], P[2][CH];
    auto load_chunk = [&](int c, int buf) {
#pragma unroll
        for (int r = 0; r < CH; ++r) {
            const int q = c * CH + r;
            const float* er = (q < 8 ? e0 : e1) + (q & 7) * istr;
            const float* sr = (q < 8 ? s0 : s1) + (q & 7) * istr;
            load_cells<CPL>(er, c0, max_u, 0.0f, E[buf][r]);
            load_cells<CPL>(sr, c0, max_u, 0.0f, S[buf][r]);
            if (RANK == 0) P[buf][r] = sr[pc];
            else P[buf][r] = 0.0f;
        }
    };
    load_chunk(0, 0);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (c + 1 < NC) load_chunk(c + 1, (c + 1) & 1);
#pragma unroll
        for (int r = 0; r < CH; ++r) {
            const int q = c * CH + r;
            float out[CPL];
            skew_step<CPL, RANK>(cs, E[c & 1][r], S[c & 1][r], P[c & 1][r], g, out);
            float* dst = (q < 8 ? st0 : st1) + (q & 7) * sstr;
            store_cells<CPL>(dst, c0, max_u, out);
            if (!TO_SMEM) reinterpret_cast<int*>(dst)[max_u + lane] = ex;
            if (q == NR - 3) h1();
            if (q == NR - 2) h2();
        }
    }
}

// Non-blocking probe of an mbarrier phase (the chain warp probes the NEXT round's barriers while it
// computes the current one, so a hand-off that is already complete costs no round trip).
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// Gradients of NRP consecutive (sweep order) rows of one stage, all loads first (ILP across rows).
// Row q of the stage: E,S probabilities, the chain's state row (own sweep) and the partner's
// scratch row x (other sweep).  gamma = alpha * p * beta / Z with the exponents split over two
// power-of-two factors.
struct PostCtx {
    float* sp;        // ring slot
    int off_e, off_s, off_x, off_v;
    int max_u, SU, UP;
    int dir, rank, lane, c0;
    int cnt;          // rows in the stage
    int t_base;       // frame of the stage's row 0 (sweep order)
    int T, U;
    int ex_state;     // this lane's exponent of the chain's state rows
    int f_M;          // likelihood: Z = sum * 2^M
    float f_inv_sum;
    bool f_dead;
    float* ge;
    float* gs;
    unsigned* status;
};

template <int CPL, int NRP, bool FULL = false, bool CHECK = true>
__device__ __forceinline__ void post_rows(const PostCtx& cx, const int q0) {
    PostCtx c = cx;
    if (FULL) {  // max_u == 32*CPL: strides become immediates and the column bounds checks fold away
        c.max_u = 32 * CPL;
        c.SU = 32 * CPL + 32;
        c.UP = 32 * CPL;
        c.off_e = 0; c.off_s = kG * 32 * CPL; c.off_x = 2 * kG * 32 * CPL;
        c.off_v = 2 * kG * 32 * CPL + kG * (32 * CPL + 32);
    }
    float E[NRP][CPL], S[NRP][CPL], VA[NRP][CPL], VB[NRP][CPL];
    int exA[NRP], exB[NRP];
#pragma unroll
    for (int r = 0; r < NRP; ++r) {
        const int q = q0 + r;
        const int idx = c.dir > 0 ? q : c.cnt - 1 - q;
        load_cells<CPL>(c.sp + c.off_e + idx * c.max_u, c.c0, c.max_u, 0.0f, E[r]);
        load_cells<CPL>(c.sp + c.off_s + idx * c.max_u, c.c0, c.max_u, 0.0f, S[r]);
        const float* xrow = c.sp + c.off_x + idx * c.SU;
        const int ex_x = reinterpret_cast<const int*>(xrow)[c.UP + c.lane];
        if (c.rank == 0) {
            load_cells<CPL>(c.sp + c.off_v + q * c.max_u, c.c0, c.max_u, 0.0f, VA[r]);
            load_cells<CPL>(xrow, c.c0, c.max_u, 0.0f, VB[r]);
            exA[r] = c.ex_state; exB[r] = ex_x;
        } else {
            load_cells<CPL>(xrow, c.c0, c.max_u, 0.0f, VA[r]);
            load_cells<CPL>(c.sp + c.off_v + q * c.max_u, c.c0, c.max_u, 0.0f, VB[r]);
            exA[r] = ex_x; exB[r] = c.ex_state;
        }
    }
    float edge[NRP];
    int exBn[NRP];
#pragma unroll
    for (int r = 0; r < NRP; ++r) {
        exBn[r] = __shfl_down_sync(kFull, exB[r], 1);
        edge[r] = __shfl_down_sync(kFull, VB[r][0], 1);
    }
#pragma unroll
    for (int r = 0; r < NRP; ++r) {
        const int t = c.t_base + c.dir * (q0 + r);
        // beta(t+1, u+1): in-lane neighbour, or lane+1's first cell re-framed
        float vbn_edge = scale_pow2(edge[r], exBn[r] - exB[r]);
        if (c.lane == 31) vbn_edge = 0.0f;
        const int kf = max(-252, min(252, exA[r] + exB[r] - c.f_M));
        const int k1 = kf >> 1;
        const float fa = pow2i(max(-126, k1));
        const float fb = pow2i(max(-126, kf - k1)) * c.f_inv_sum;
        float g1[CPL], g2[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const float nb = (i + 1 < CPL) ? VB[r][i + 1] : vbn_edge;
            const float va = VA[r][i] * fa;
            g1[i] = c.f_dead ? 0.0f : va * ((E[r][i] * VB[r][i]) * fb);
            g2[i] = c.f_dead ? 0.0f : va * ((S[r][i] * nb) * fb);
        }
        store_cells_cs<CPL>(c.ge + (size_t)t * c.max_u, c.c0, c.max_u, g1);
        store_cells_cs<CPL>(c.gs + (size_t)t * c.max_u, c.c0, c.max_u, g2);
        // consistency: occupancy of the terminal cell (alpha side) / of frame 0 (beta side) must
        // be 1 — these are independent likelihood estimates.
        if (CHECK && !c.f_dead && (t == c.T - 1 || t == 0)) {
            bool bad = false;
            if (c.rank == 0 && t == c.T - 1) {
#pragma unroll
                for (int i = 0; i < CPL; ++i)
                    if (c.c0 + i == c.U - 1) bad = !(fabsf(g1[i] - 1.0f) < kBfConsistency);
            }
            if (c.rank == 1 && t == 0 && c.lane == 0) bad = !(fabsf(g1[0] + g2[0] - 1.0f) < kBfConsistency);
            if (bad) atomicOr(c.status, (unsigned)kBfInconsistent);
        }
    }
}

// ---- shared-memory flags ---------------------------------------------------------------------------
// Hand-offs that involve the recursion warp are plain shared-memory words, not mbarriers: an
// mbarrier probe costs 90-150 cycles on an idle SM and several hundred while six helper warps poll
// the same unit (measured: ~1400 cycles of fixed cost per round), and it cannot be issued ahead of
// its use.  A flag is an ordinary LDS (29 cycles) that the recursion warp issues a round early.
__device__ __forceinline__ int flag_load(const int* f) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(f)) : "memory");
    return v;
}
__device__ __forceinline__ int2 flag_load2(const int* f) {
    int2 v;
    asm volatile("ld.volatile.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_u32(f)) : "memory");
    return v;
}
// Publishes `v` after everything the warp wrote before (call by one lane after __syncwarp()).
__device__ __forceinline__ void flag_publish(int* f, int v) {
    __threadfence_block();
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(f)), "r"(v) : "memory");
}

// Phase-1 duty of the helpers' "post" slot: copy NR state rows of a stage (values + the lane
// exponents of the round) from shared memory to the global scratch rows the partner CTA reads.
template <int CPL, int NR, bool FULL = false>
__device__ __forceinline__ void copy_out_rows(const float* st0, float* g0, const long long gstride, const int ex,
                                              const int c0, int max_u, int UP, const int lane) {
    if (FULL) { max_u = 32 * CPL; UP = 32 * CPL; }
    float V[NR][CPL];
#pragma unroll
    for (int r = 0; r < NR; ++r) load_cells<CPL>(st0 + r * max_u, c0, max_u, 0.0f, V[r]);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        float* dst = g0 + (long long)r * gstride;
        store_cells<CPL>(dst, c0, max_u, V[r]);
        reinterpret_cast<int*>(dst)[UP + lane] = ex;
    }
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
    // header: raw_full mbarriers [16] | prep_flag [16][2] | state_flag [16] | done [8] | llinfo | (log re-run barriers at 576)
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem_raw);           // TMA → prep helpers
    int* prep_flag = reinterpret_cast<int*>(smem_raw + 128);              // [slot][half] = use+1 once prepped
    int* state_flag = reinterpret_cast<int*>(smem_raw + 256);             // [slot] = use+1 once the state rows are written
    int* done = reinterpret_cast<int*>(smem_raw + 320);                   // [helper] = 1 + last global stage it finished
    float* llinfo = reinterpret_cast<float*>(smem_raw + 384);             // [0] M (int bits) [1] 1/sum [2] dead
    float* ring = reinterpret_cast<float*>(smem_raw + kBfHeaderBytes);
    // per slot: e[8][max_u] | s[8][max_u] | x[8][SU] | state[8][max_u] | state_exp[32]
    const int off_e = 0, off_s = kG * max_u, off_x = 2 * kG * max_u, off_v = off_x + kG * SU,
              off_ve = off_v + kG * max_u;
    const int stage_floats = off_ve + 32;

    if (tid < 64) reinterpret_cast<int*>(smem_raw + 128)[tid] = 0;  // all flags
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(smem_u32(raw_full + s), 1);
        fence_mbar_init();
        // (the status words are zeroed on the stream before the launch: a fault bit the partner CTA raises early
        // cannot be wiped by a late reset here)
        if (rank == 0 && p.force_fallback) atomicOr(p.status + b, (unsigned)kBfForced);
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
    // profiling aid: cycles spent blocked on each kind of hand-off, per warp
    long long st_wait[4] = {0, 0, 0, 0};
    const long long st_t0 = p.stats ? clock64() : 0;
    long long st_sync = 0, st_phase0 = 0, st_prep = 0, st_post = 0;
    auto timed_cluster_sync = [&]() {
        if (p.stats) {
            const long long t0 = clock64();
            st_phase0 = t0 - st_t0;
            cluster.sync();
            st_sync += clock64() - t0;
        } else {
            cluster.sync();
        }
    };
    long long* tl = (p.stats && blockIdx.x == 0 && (p.debug_skip & 128)) ? p.stats + (size_t)gridDim.x * 8 * 16 : nullptr;  // timeline of CTA 0
    auto tl_mark = [&](int role, int stage, int ev) {
        if (tl && lane == 0 && stage < 256) tl[(role * 256 + stage) * 4 + ev] = clock64() - st_t0;
    };

    // =================================================================================================
    if (warp == 4) {
        // ------------------------------- producer -------------------------------
        // L2 prefetch runs pf_rows ahead of the shared-memory ring, over the whole sweep (both
        // phases): rank 0 walks rows 0..T-1, rank 1 rows T-1..0, so the ring's TMA loads hit L2
        // instead of paying the HBM latency with only NS stages in flight.
        int pfJ = 0;  // sweep positions [0, pfJ) are already requested
        {
            const int upto = min(p.pf_rows, T);  // start-up: request the first pf_rows rows, 16 per lane
            const int j = lane * 16;
            if (j < upto) {
                const int n = min(16, upto - j);
                const int pr0 = dir > 0 ? j : T - j - n;
                prefetch_l2(le + (size_t)pr0 * max_u, (uint32_t)n * (uint32_t)max_u * 4u);
                prefetch_l2(ls + (size_t)pr0 * max_u, (uint32_t)n * (uint32_t)max_u * 4u);
            }
            pfJ = upto;
        }
        int Jdone = 0;  // sweep position after the last issued stage
        unsigned kg = 0;
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == 1) {
                timed_cluster_sync();
                fence_proxy_async();
            }
            const Phase& P = ph[phase];
            // Up to four stages are issued per round: lane l handles array (l % 3) of stage (l / 3) of the
            // batch, so one cp.async.bulk instruction starts up to 12 copies (the issue cost is per
            // instruction, ~150-300 cycles, not per copy).  Lanes 12-13 issue the L2 prefetches.  No
            // proxy fence per stage: a slot is only refilled after the helpers that read it last
            // published `done`, and a fence.proxy.async here would wait for the copies still in
            // flight, collapsing the ring to a single outstanding stage.
            // The batch must stay shallow relative to the ring: a helper pair frees the slot of stage y
            // (post) only after it prepared stage y+3, so a batch that waits for the slots of k0-NS ..
            // k0+NB-1-NS needs the copies of stage k0+NB+2-NS to have been issued by an EARLIER batch:
            // NS > NB + 2 + (NB - 1) (shallow rings: max_u > 128).  A deeper batch dead-locks.
            const int NB = NS >= 10 ? 4 : (NS >= 6 ? 2 : 1);
            const int jb = lane / 3, arr = lane - jb * 3;
            const int narr = P.with_x ? 3 : 2;
            for (int k0 = 0; k0 < P.nst; k0 += NB) {
                const int k = k0 + jb;
                const bool mine = lane < 3 * NB && k < P.nst && arr < narr;
                const unsigned kk = kg + (unsigned)k;
                const int slot = (int)(kk % (unsigned)NS);
                const int j0 = k * kG;
                const int cnt = min(kG, P.n - j0);
                const int r0 = dir > 0 ? P.t0 + j0 : P.t0 - j0 - cnt + 1;
                const uint32_t bar = smem_u32(raw_full + slot);
                const uint32_t bytes_e = (uint32_t)cnt * (uint32_t)max_u * 4u;
                const uint32_t bytes_x = P.with_x ? (uint32_t)cnt * (uint32_t)SU * 4u : 0u;
                if (lane == 0) tl_mark(1, (int)kk, 0);
                if (mine && arr == 0) {
                    if (kk >= (unsigned)NS) {
                        // previous occupant of the slot: global stage kk - NS, owned (both halves) by
                        // helper pair (its phase-local index % 3)
                        const int old = (int)kk - NS;
                        const int old_local = old < ph[0].nst ? old : old - ph[0].nst;
                        const int* d = done + 2 * (old_local % kPairs);
                        for (;;) {
                            const int2 v = flag_load2(d);
                            if (v.x > old && v.y > old) break;
                            __nanosleep((unsigned)p.pf_sleep_ns);
                        }
                    }
                    mbar_expect_tx(bar, 2u * bytes_e + bytes_x);
                }
                __syncwarp();
                if (mine) {
                    float* dst = ring + (size_t)slot * stage_floats;
                    const float* src = arr == 0 ? le + (size_t)r0 * max_u
                                     : arr == 1 ? ls + (size_t)r0 * max_u
                                                : scr + (size_t)(r0 + P.xoff) * SU;
                    const uint32_t d = smem_u32(dst + (arr == 0 ? off_e : arr == 1 ? off_s : off_x));
                    bulk_g2s(d, src, arr < 2 ? bytes_e : bytes_x, bar);
                }
                if (lane == 0) tl_mark(1, (int)kk, 1);
                Jdone += min(NB * kG, P.n - k0 * kG);
                // L2 prefetch: keep pf_rows requested ahead of the ring, 32 rows (2 x 16) per round
                if (pfJ < T && pfJ < Jdone + p.pf_rows) {
                    const int n = min(32, T - pfJ);
                    const int h16 = lane - 12;  // lanes 12,13: first / second 16-row block
                    if (h16 >= 0 && h16 < 2 && h16 * 16 < n) {
                        const int nn = min(16, n - h16 * 16);
                        const int pj = pfJ + h16 * 16;
                        const int pr0 = dir > 0 ? pj : T - pj - nn;
                        const uint32_t pbytes = (uint32_t)nn * (uint32_t)max_u * 4u;
                        prefetch_l2(le + (size_t)pr0 * max_u, pbytes);
                        prefetch_l2(ls + (size_t)pr0 * max_u, pbytes);
                    }
                    pfJ += n;
                }
            }
            __syncwarp();
            kg += (unsigned)P.nst;
        }
    } else if (warp == 0) {
        // ------------------------------- chain -------------------------------
        ChainState<CPL> cs;
        cs.init((int)rank, lane, U);
        int ex = 0;     // this lane's frame (shared exponent of its CPL cells)
        int nb_ex = 0;  // the feeding lane's frame
        const bool edge_lane = rank == 0 ? lane == 0 : lane == 31;  // the lane without a feeder
        const bool full_u = max_u == 32 * CPL;
        float g = edge_lane ? 0.0f : 1.0f;  // 2^(nb_ex - ex); all frames start at 0
        // Pipelined re-normalisation: the new frame of round k+1 is DECIDED during round k (two
        // shuffles whose results are not needed before the next round) and APPLIED at the start of
        // round k+1, so no shuffle latency sits between the rows.
        bool have_dec = false;
        int ex_dec = 0, nbex_dec = 0;
        auto apply_decision = [&]() {
            if (!have_dec) return;
            const int shift = ex - ex_dec;
#pragma unroll
            for (int i = 0; i < CPL; ++i) cs.a[i] = scale_pow2(cs.a[i], shift);
            // pending neighbour values follow the neighbour's re-framing
            const int fshift = nb_ex - nbex_dec;
            cs.inA = scale_pow2(cs.inA, fshift);
            cs.inB = scale_pow2(cs.inB, fshift);
            ex = ex_dec;
            nb_ex = nbex_dec;
            g = edge_lane ? 0.0f : pow2i(max(-126, min(126, nb_ex - ex)));
            have_dec = false;
        };
        // decision, part 1: magnitudes of this lane and (shuffle in flight) of the feeder's edge cell
        auto decide_1 = [&](int& own, int& nbmag) {
            float mx = cs.a[0];
#pragma unroll
            for (int i = 1; i < CPL; ++i) mx = fmaxf(mx, cs.a[i]);
            own = mx > 0.0f ? ex + ilogb_pos(mx) - kTarget : kNoMass;
            const float edge = rank == 0 ? cs.a[CPL - 1] : cs.a[0];
            const int amag = edge > 0.0f ? ex + ilogb_pos(edge) : kNoMass;
            nbmag = rank == 0 ? __shfl_up_sync(kFull, amag, 1) : __shfl_down_sync(kFull, amag, 1);
        };
        // decision, part 2: new frame, and (shuffle in flight) the feeder's new frame
        auto decide_2 = [&](int own, int nbmag) {
            if (edge_lane) nbmag = kNoMass;
            int nw = max(own, nbmag - kTarget - kSlack);
            if (nw <= kNoMass / 2) nw = ex;  // nothing here and nothing arriving: keep the frame
            ex_dec = nw;
            nbex_dec = rank == 0 ? __shfl_up_sync(kFull, nw, 1) : __shfl_down_sync(kFull, nw, 1);
            have_dec = true;
        };
        unsigned kg = 0;
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == 1) {
                // complete the last phase-1 row, hand it over, then meet the partner
                cs.flush((int)rank, g);
                float* r = scr + (size_t)(rank == 0 ? m - 1 : m) * SU;
                store_cells<CPL>(r, c0, max_u, cs.a);
                reinterpret_cast<int*>(r)[UP + lane] = ex;
                __threadfence();
                fence_proxy_async();
                timed_cluster_sync();
            }
            const int Pn = ph[phase].n, Pnst = ph[phase].nst;
            int slot = (int)(kg % (unsigned)NS);
            int use1 = (int)(kg / (unsigned)NS) + 1;  // value the slot's flags take once it is ready
            // flags of this round's slots, loaded one round ahead
            int2 fA = make_int2(0, 0), fB = make_int2(0, 0);
            for (int k = 0; k < Pnst;) {
                const bool two = full_u && Pn - k * kG >= 2 * kG && !(p.debug_skip & 2);  // two full stages per round
                const int slot2 = slot + 1 == NS ? 0 : slot + 1;
                const int use2 = slot + 1 == NS ? use1 + 1 : use1;
                tl_mark(0, (int)kg + k, 0);
                {
                    const long long t0 = p.stats ? clock64() : 0;
                    while (!__all_sync(kFull, fA.x >= use1 && fA.y >= use1)) fA = flag_load2(prep_flag + 2 * slot);
                    if (two)
                        while (!__all_sync(kFull, fB.x >= use2 && fB.y >= use2)) fB = flag_load2(prep_flag + 2 * slot2);
                    if (p.stats) st_wait[1] += clock64() - t0;
                }
                tl_mark(0, (int)kg + k, 1);
                const int adv = two ? 2 : 1;
                // request the flags of the next round's slots now; they are looked at a round later
                int slot_n = slot, use_n = use1;
                for (int z = 0; z < adv; ++z)
                    if (++slot_n == NS) { slot_n = 0; ++use_n; }
                {
                    const int s4 = slot_n + 1 == NS ? 0 : slot_n + 1;
                    fA = flag_load2(prep_flag + 2 * slot_n);
                    fB = flag_load2(prep_flag + 2 * s4);
                }
                float* spA = slot_ptr(slot);
                float* spB = slot_ptr(slot2);
                const int j0 = k * kG;
                const int cnt = min(kG, Pn - j0);
                const long long tr0 = p.stats ? clock64() : 0;
                apply_decision();
                reinterpret_cast<int*>(spA + off_ve)[lane] = ex;
                if (two) reinterpret_cast<int*>(spB + off_ve)[lane] = ex;
                int own = kNoMass, nbmag = kNoMass;
                auto d1 = [&]() { decide_1(own, nbmag); };
                auto d2 = [&]() { decide_2(own, nbmag); };
                const long long tr1 = p.stats ? clock64() : 0;
                st_post += tr1 - tr0;
                // first row of the stage in sweep order (rank 1 walks the slot's rows backwards);
                // the state rows go to the slot's shared state rows in both phases
                const int eo = dir > 0 ? 0 : (kG - 1) * max_u;
                if (full_u && cnt == kG) {
                    const float* eA = spA + off_e + eo;
                    const float* sA = spA + off_s + eo;
                    const float* eB = spB + off_e + eo;
                    const float* sB = spB + off_s + eo;
                    if (two) {
                        if (rank == 0) chain_round_skew<CPL, 0, true, 16>(cs, g, eA, sA, spA + off_v, ex, lane, d1, d2, eB, sB, spB + off_v);
                        else chain_round_skew<CPL, 1, true, 16>(cs, g, eA, sA, spA + off_v, ex, lane, d1, d2, eB, sB, spB + off_v);
                    } else {
                        if (rank == 0) chain_round_skew<CPL, 0, true, 8>(cs, g, eA, sA, spA + off_v, ex, lane, d1, d2);
                        else chain_round_skew<CPL, 1, true, 8>(cs, g, eA, sA, spA + off_v, ex, lane, d1, d2);
                    }
                } else {
                    // generic rows: short last stage of a phase, or max_u < 32*CPL
                    const int pc = c0 > 0 ? c0 - 1 : 0;
                    for (int q = 0; q < cnt; ++q) {
                        const int idx = dir > 0 ? q : cnt - 1 - q;
                        const float* er = spA + off_e + idx * max_u;
                        const float* sr = spA + off_s + idx * max_u;
                        float E[CPL], S[CPL], out[CPL];
                        load_cells<CPL>(er, c0, max_u, 0.0f, E);
                        load_cells<CPL>(sr, c0, max_u, 0.0f, S);
                        const float P = (rank == 0 && pc < max_u) ? sr[pc] : 0.0f;
                        if (rank == 0) skew_step<CPL, 0>(cs, E, S, P, g, out);
                        else skew_step<CPL, 1>(cs, E, S, P, g, out);
                        store_cells<CPL>(spA + off_v + q * max_u, c0, max_u, out);
                    }
                    decide_1(own, nbmag);
                    decide_2(own, nbmag);
                }
                const long long tr2 = p.stats ? clock64() : 0;
                st_prep += tr2 - tr1;
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(state_flag + slot)), "r"(use1) : "memory");
                    if (two)
                        asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(state_flag + slot2)), "r"(use2) : "memory");
                }
                tl_mark(0, (int)kg + k, 2);
                if (two) tl_mark(0, (int)kg + k + 1, 2);
                if (p.stats) st_wait[3] += clock64() - tr2;
                k += adv;
                slot = slot_n;
                use1 = use_n;
            }
            kg += (unsigned)Pnst;
        }
    } else {
        // ------------------------------- helpers -------------------------------
        // Six helper warps form three pairs; pair p owns stages k = p, p+3, p+6, ... of each phase and
        // each warp of the pair owns one half (4 rows) of the stage, processed together for ILP.
        // Per stage: prep (log-probs → probabilities, in place) ahead of the recursion, and behind it
        // "post": phase 1 copies the state rows to the global scratch, phase 2 emits the gradients.
        const int h = warp < 4 ? warp - 1 : warp - 2;  // warps 1,2,3,5,6,7 → 0..5
        const int pair = h >> 1, half = h & 1;
        const bool full_u = max_u == 32 * CPL;
        bool me[CPL], ms[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            me[i] = c0 + i < U;
            ms[i] = c0 + i < U - 1;
        }
        unsigned kg = 0;
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == 1) {
                __threadfence();       // this warp's scratch rows → visible to the partner CTA's TMA reads
                fence_proxy_async();
                timed_cluster_sync();
            }
            const Phase& P = ph[phase];
            float f_inv_sum = 0.0f;
            int f_M = 0;
            bool f_dead = false, have_ll = false;
            for (int it = pair; it < P.nst + kPairs; it += kPairs) {
                // ---- prep(it) ----
                if (it < P.nst) {
                    const int k = it;
                    const unsigned kk = kg + (unsigned)k;
                    const int slot = slot_of(kk);
                    if (half == 0) tl_mark(2, (int)kk, 0);
                    {
                        const long long t0 = p.stats ? clock64() : 0;
                        mbar_wait_warp(smem_u32(raw_full + slot), use_of(kk) & 1u);
                        if (p.stats) st_wait[0] += clock64() - t0;
                    }
                    if (half == 0) tl_mark(2, (int)kk, 1);
                    const long long tp0 = p.stats ? clock64() : 0;
                    float* sp = slot_ptr(slot);
                    const int j0 = k * kG;
                    const int cnt = min(kG, P.n - j0);
                    const int q0 = half * kHalf;
                    const int nr = max(0, min(kHalf, cnt - q0));
                    const int idx0 = dir > 0 ? q0 : cnt - 1 - q0;  // memory row of consumption row q0
                    float* e0 = sp + off_e + idx0 * max_u;
                    float* s0 = sp + off_s + idx0 * max_u;
                    const int t0r = P.t0 + dir * (j0 + q0);
                    if (p.debug_skip & 32) {
                    } else if (nr == kHalf && full_u) {
                        if (dir > 0) prep_rows<CPL, kHalf, true>(e0, s0, 0, t0r, 1, T, me, ms, c0, max_u);
                        else prep_rows<CPL, kHalf, true>(e0, s0, 0, t0r, -1, T, me, ms, c0, max_u);
                    } else if (nr == kHalf) prep_rows<CPL, kHalf>(e0, s0, dir * max_u, t0r, dir, T, me, ms, c0, max_u);
                    else
                        for (int r = 0; r < nr; ++r)
                            prep_rows<CPL, 1>(e0 + r * dir * max_u, s0 + r * dir * max_u, dir * max_u, t0r + dir * r,
                                              dir, T, me, ms, c0, max_u);
                    __syncwarp();
                    if (lane == 0) flag_publish(prep_flag + 2 * slot + half, (int)use_of(kk) + 1);
                    if (half == 0) tl_mark(2, (int)kk, 2);
                    if (p.stats) st_prep += clock64() - tp0;
                }
                // ---- post(it - kPairs) ----
                if (it >= kPairs) {
                    const int k = it - kPairs;
                    const unsigned kk = kg + (unsigned)k;
                    const int slot = slot_of(kk);
                    const int need = (int)use_of(kk) + 1;
                    if (half == 0) tl_mark(3, (int)kk, 0);
                    {
                        const long long t0 = p.stats ? clock64() : 0;
                        while (flag_load(state_flag + slot) < need) __nanosleep(40);
                        if (p.stats) st_wait[2] += clock64() - t0;
                    }
                    if (half == 0) tl_mark(3, (int)kk, 1);
                    const long long tq0 = p.stats ? clock64() : 0;
                    float* sp = slot_ptr(slot);
                    const int j0 = k * kG;
                    const int cnt = min(kG, P.n - j0);
                    const int q0 = half * kHalf;
                    const int nr = max(0, min(kHalf, cnt - q0));
                    const int ex_state = reinterpret_cast<const int*>(sp + off_ve)[lane];
                    if (phase == 0) {
                        // state row q = alpha(t) → scratch row t (rank 0), beta(t+1) → row t+1 (rank 1)
                        const int t0r = P.t0 + dir * (j0 + q0) + (rank == 0 ? 0 : 1);
                        float* g0 = scr + (size_t)t0r * SU;
                        const float* st0 = sp + off_v + q0 * max_u;
                        if (nr == kHalf && full_u) copy_out_rows<CPL, kHalf, true>(st0, g0, (long long)dir * SU, ex_state, c0, max_u, UP, lane);
                        else if (nr == kHalf) copy_out_rows<CPL, kHalf>(st0, g0, (long long)dir * SU, ex_state, c0, max_u, UP, lane);
                        else
                            for (int r = 0; r < nr; ++r)
                                copy_out_rows<CPL, 1>(st0 + r * max_u, g0 + (long long)r * dir * SU, 0, ex_state, c0, max_u, UP, lane);
                    } else {
                        const bool ll_owner = (k == 0 && half == 0);
                        int r_first = 0;
                        if (ll_owner) {
                            // log-likelihood from the meeting row j = 0: Z = sum_u alpha(m-1,u)·beta(m-1,u)
                            const int idx = dir > 0 ? 0 : cnt - 1;
                            float E[CPL], S[CPL], VA[CPL], VB[CPL];
                            load_cells<CPL>(sp + off_e + idx * max_u, c0, max_u, 0.0f, E);
                            load_cells<CPL>(sp + off_s + idx * max_u, c0, max_u, 0.0f, S);
                            const float* xrow = sp + off_x + idx * SU;
                            const int ex_x = reinterpret_cast<const int*>(xrow)[UP + lane];
                            int exA, exB;
                            if (rank == 0) {
                                load_cells<CPL>(sp + off_v, c0, max_u, 0.0f, VA);
                                load_cells<CPL>(xrow, c0, max_u, 0.0f, VB);
                                exA = ex_state; exB = ex_x;
                            } else {
                                load_cells<CPL>(xrow, c0, max_u, 0.0f, VA);
                                load_cells<CPL>(sp + off_v, c0, max_u, 0.0f, VB);
                                exA = ex_x; exB = ex_state;
                            }
                            const int exBn = __shfl_down_sync(kFull, exB, 1);
                            float vbn_edge = scale_pow2(__shfl_down_sync(kFull, VB[0], 1), exBn - exB);
                            if (lane == 31) vbn_edge = 0.0f;
                            const int EL = exA + exB;
                            float w = 0.0f;
#pragma unroll
                            for (int i = 0; i < CPL; ++i) {
                                const float nb = (i + 1 < CPL) ? VB[i + 1] : vbn_edge;
                                w += VA[i] * (E[i] * VB[i] + S[i] * nb);
                            }
                            const bool finite = w == w && w < 3.0e38f;
                            int M = (finite && w > 0.0f) ? EL + ilogb_pos(w) : kNoMass;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) M = max(M, __shfl_xor_sync(kFull, M, o));
                            const float part = (finite && w > 0.0f) ? scale_pow2(w, EL - M) : 0.0f;
                            const float sum = warp_sum(part);
                            const unsigned bad = __ballot_sync(kFull, !finite);
                            unsigned st = 0;
                            if (bad) st |= kBfNonFinite;
                            if (M <= kNoMass / 2 || !(sum > 0.0f)) st |= kBfNoMass;
                            f_M = M;
                            f_inv_sum = st ? 0.0f : 1.0f / sum;
                            f_dead = st != 0;
                            have_ll = true;
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
                            // rank 0's meeting row only yields the likelihood: its gradients belong to rank 1
                            if (rank == 0) r_first = 1;
                        } else if (!have_ll) {  // wait for the log-likelihood of the meeting row
                            named_bar_sync(1, 32 * kHelpers);
                            f_M = __float_as_int(llinfo[0]);
                            f_inv_sum = llinfo[1];
                            f_dead = llinfo[2] != 0.0f;
                            have_ll = true;
                        }
                        PostCtx pc;
                        pc.sp = sp; pc.off_e = off_e; pc.off_s = off_s; pc.off_x = off_x; pc.off_v = off_v;
                        pc.max_u = max_u; pc.SU = SU; pc.UP = UP;
                        pc.dir = dir; pc.rank = (int)rank; pc.lane = lane; pc.c0 = c0;
                        pc.cnt = cnt; pc.t_base = P.t0 + dir * j0; pc.T = T; pc.U = U;
                        pc.ex_state = ex_state; pc.f_M = f_M; pc.f_inv_sum = f_inv_sum; pc.f_dead = f_dead;
                        pc.ge = ge; pc.gs = gs; pc.status = p.status + b;
                        if (p.debug_skip & 64) {
                        } else if (nr == kHalf && r_first == 0 && full_u && k + 1 < P.nst && !(p.debug_skip & 4)) {
                            // hot path: full stage that does not hold the sweep's last row (no consistency check)
                            if (rank == 0) { pc.dir = 1; pc.rank = 0; post_rows<CPL, kHalf, true, false>(pc, q0); }
                            else { pc.dir = -1; pc.rank = 1; post_rows<CPL, kHalf, true, false>(pc, q0); }
                        } else if (nr == kHalf && r_first == 0 && !(p.debug_skip & 4)) post_rows<CPL, kHalf>(pc, q0);
                        else
                            for (int r = r_first; r < nr; ++r) post_rows<CPL, 1>(pc, q0 + r);
                    }
                    __syncwarp();
                    // the slot may be refilled once both halves are through with it (no release needed for
                    // the global stores: `done` only guards the slot's shared memory)
                    if (lane == 0) asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(done + h)), "r"((int)kk + 1) : "memory");
                    if (half == 0) tl_mark(3, (int)kk, 2);
                    if (p.stats) st_post += clock64() - tq0;
                }
            }
            if (phase == 1 && !have_ll) named_bar_sync(1, 32 * kHelpers);  // every helper meets the barrier once
            kg += (unsigned)P.nst;
        }
    }
    if (p.stats && lane == 0) {
        long long* o = p.stats + ((size_t)blockIdx.x * 8 + warp) * 16;
        o[0] = clock64() - st_t0;
        o[1] = st_wait[0]; o[2] = st_wait[1]; o[3] = st_wait[2]; o[4] = st_wait[3];
        o[5] = st_sync; o[6] = st_phase0; o[7] = st_prep * 1000000 + st_post / 1;
    }
}

}  // namespace lattice
}  // namespace ssnt
