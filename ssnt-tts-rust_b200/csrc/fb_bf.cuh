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
constexpr int kSlack = 32;        // a lane's frame may sit this far below its feeding neighbour's edge
constexpr float kBfConsistency = 2e-5f;  // the three likelihood estimates must agree this well (typ. 1e-6)
constexpr int kNoMass = -100000;  // exponent key of an all-zero lane
constexpr int kBfHeaderBytes = 768;  // mbarriers (512) | llinfo (64) | log-domain re-run barriers (128) | pad
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

// Conversion of NR consecutive (in sweep order) rows of a stage from log-probabilities to
// probabilities, in place, with the length masks: tokens >= U have e = s = 0, the last token and
// the last frame cannot shift.  All loads first, then the EX2s, then the stores (ILP across rows).
template <int CPL, int NR>
__device__ __forceinline__ void prep_rows(float* e0, float* s0, const int in_stride, const int t_first,
                                          const int dir, const int T, const bool (&me)[CPL],
                                          const bool (&ms)[CPL], const int c0, const int max_u) {
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

// NR consecutive rows of the recursion: the rows' probabilities are pulled into registers first
// (all shared-memory latencies overlap), then the rows run as straight-line code.  The state
// BEFORE each step is written out (scratch row in phase 1, shared state row in phase 2).
// RANK 0: alpha'(u) = alpha(u) e(u) + alpha(u-1) s(u-1);  RANK 1: beta'(u) = e(u) beta(u) + s(u) beta(u+1).
struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};
// H1 / H2 are called after the rows with local index D1 / D2 (-1: never): the recursion warp hangs
// the two halves of its re-normalisation decision there, so their shuffles overlap the last rows.
// FULL: max_u == 32*CPL, the common case (U = 32, 64, 128, 256): every stride is a compile-time
// constant, so the loads and stores use immediate offsets and need no bounds predicates.
template <int CPL, int RANK, bool TO_SMEM, int NR, bool FULL, int D1 = -1, int D2 = -1, typename H1 = NoHook,
          typename H2 = NoHook>
__device__ __forceinline__ void chain_rows(float (&v)[CPL], const float g, const float* e0, const float* s0,
                                           int in_stride, float* st0, long long st_stride, const int ex, int UP,
                                           const int lane, const int c0, int max_u, H1 h1 = NoHook(),
                                           H2 h2 = NoHook()) {
    if (FULL) {
        max_u = 32 * CPL;
        UP = 32 * CPL;
        in_stride = RANK == 0 ? 32 * CPL : -32 * CPL;
        st_stride = TO_SMEM ? 32 * CPL : (RANK == 0 ? 32 * CPL + 32 : -(32 * CPL + 32));
    }
    float E[NR][CPL], S[NR][CPL];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        load_cells<CPL>(e0 + r * in_stride, c0, max_u, 0.0f, E[r]);
        load_cells<CPL>(s0 + r * in_stride, c0, max_u, 0.0f, S[r]);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        float* dst = st0 + (long long)r * st_stride;
        store_cells<CPL>(dst, c0, max_u, v);
        if (!TO_SMEM) reinterpret_cast<int*>(dst)[UP + lane] = ex;
        if (RANK == 0) {
            float bsh[CPL];
#pragma unroll
            for (int i = 0; i < CPL; ++i) bsh[i] = v[i] * S[r][i];
            const float in = __shfl_up_sync(kFull, bsh[CPL - 1], 1);
#pragma unroll
            for (int i = CPL - 1; i >= 1; --i) v[i] = fmaf(v[i], E[r][i], bsh[i - 1]);
            v[0] = fmaf(in, g, v[0] * E[r][0]);
        } else {
            const float in = __shfl_down_sync(kFull, v[0], 1) * g;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const float nb = (i + 1 < CPL) ? v[i + 1] : in;
                v[i] = fmaf(E[r][i], v[i], S[r][i] * nb);
            }
        }
        if (r == D1) h1();
        if (r == D2) h2();
    }
}

// One full stage (8 rows) with the row loads software-pipelined: the probabilities of chunk c+1 are
// requested before chunk c is computed, so shared-memory latency (and queueing behind the helper
// warps' traffic) overlaps the arithmetic.  Hooks: d1 after row 5, d2 after row 6.
template <int CPL, int RANK, bool TO_SMEM, bool FULL, int DBG, typename H1, typename H2>
__device__ __forceinline__ void chain_stage(float (&v)[CPL], const float g, const float* e0, const float* s0,
                                            int in_stride, float* st0, long long st_stride, const int ex, int UP,
                                            const int lane, const int c0, int max_u, H1 h1, H2 h2) {
    if (FULL) {
        max_u = 32 * CPL;
        UP = 32 * CPL;
        in_stride = RANK == 0 ? 32 * CPL : -32 * CPL;
        st_stride = TO_SMEM ? 32 * CPL : (RANK == 0 ? 32 * CPL + 32 : -(32 * CPL + 32));
    }
    constexpr int NR = CPL <= 4 ? 4 : 2;  // rows per chunk (register budget)
    constexpr int NC = kG / NR;
    float E[2][NR][CPL], S[2][NR][CPL];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        if (DBG & 2) {
#pragma unroll
            for (int i = 0; i < CPL; ++i) { E[0][r][i] = 0.6f + 1e-3f * lane; S[0][r][i] = 0.4f; E[1][r][i] = 0.6f; S[1][r][i] = 0.4f + 1e-3f * lane; }
        } else {
            load_cells<CPL>(e0 + r * in_stride, c0, max_u, 0.0f, E[0][r]);
            load_cells<CPL>(s0 + r * in_stride, c0, max_u, 0.0f, S[0][r]);
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (c + 1 < NC && !(DBG & 2)) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                load_cells<CPL>(e0 + ((c + 1) * NR + r) * in_stride, c0, max_u, 0.0f, E[(c + 1) & 1][r]);
                load_cells<CPL>(s0 + ((c + 1) * NR + r) * in_stride, c0, max_u, 0.0f, S[(c + 1) & 1][r]);
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int q = c * NR + r;
            float* dst = st0 + (long long)q * st_stride;
            if (!(DBG & 1)) {
                store_cells<CPL>(dst, c0, max_u, v);
                if (!TO_SMEM) reinterpret_cast<int*>(dst)[UP + lane] = ex;
            }
            if (RANK == 0) {
                float bsh[CPL];
#pragma unroll
                for (int i = 0; i < CPL; ++i) bsh[i] = v[i] * S[c & 1][r][i];
                const float in = __shfl_up_sync(kFull, bsh[CPL - 1], 1);
#pragma unroll
                for (int i = CPL - 1; i >= 1; --i) v[i] = fmaf(v[i], E[c & 1][r][i], bsh[i - 1]);
                v[0] = fmaf(in, g, v[0] * E[c & 1][r][0]);
            } else {
                const float in = __shfl_down_sync(kFull, v[0], 1) * g;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const float nb = (i + 1 < CPL) ? v[i + 1] : in;
                    v[i] = fmaf(E[c & 1][r][i], v[i], S[c & 1][r][i] * nb);
                }
            }
            if (q == 5) h1();
            if (q == 6) h2();
        }
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
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // 4 x NS barriers (NS <= 16)
    uint64_t* raw_full = bars;
    uint64_t* prep_full = bars + 16;
    uint64_t* state_full = bars + 32;
    uint64_t* slot_free = bars + 48;
    float* llinfo = reinterpret_cast<float*>(smem_raw + 512);  // [0] M (int bits) [1] 1/sum [2] dead
    float* ring = reinterpret_cast<float*>(smem_raw + kBfHeaderBytes);
    // per slot: e[8][max_u] | s[8][max_u] | x[8][SU] | state[8][max_u] | state_exp[32]
    const int off_e = 0, off_s = kG * max_u, off_x = 2 * kG * max_u, off_v = off_x + kG * SU,
              off_ve = off_v + kG * max_u;
    const int stage_floats = off_ve + 32;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(smem_u32(raw_full + s), 1);
            mbar_init(smem_u32(prep_full + s), 2);
            mbar_init(smem_u32(state_full + s), 1);
            mbar_init(smem_u32(slot_free + s), 2);
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
    // profiling aid: cycles spent blocked on each kind of barrier, per warp
    long long st_wait[4] = {0, 0, 0, 0};
    const long long st_t0 = p.stats ? clock64() : 0;
    long long st_sync = 0, st_phase0 = 0, st_prep = 0, st_post = 0;
    long long st_min = 1 << 30, st_max = 0, st_n = 0, st_lt400 = 0, st_lt800 = 0;  // chain: per-stage row time histogram
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
    long long* tl = (p.stats && blockIdx.x == 0) ? p.stats + (size_t)gridDim.x * 8 * 16 : nullptr;  // timeline of CTA 0
    auto tl_mark = [&](int role, int stage, int ev) {
        if (tl && lane == 0 && stage < 256) tl[(role * 256 + stage) * 4 + ev] = clock64() - st_t0;
    };
    // Warp-uniform wait (see mbar_wait_warp): keeps the role's warp converged.
    auto timed_wait = [&](int kind, uint32_t bar, uint32_t parity) {
        const long long t0 = p.stats ? clock64() : 0;
        mbar_wait_warp(bar, parity);
        if (p.stats) st_wait[kind] += clock64() - t0;
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
            // Four stages are issued per round: lane l handles array (l % 3) of stage (l / 3) of the
            // batch, so one cp.async.bulk instruction starts up to 12 copies (the issue cost is per
            // instruction, ~150-300 cycles, not per copy).  Lanes 12-13 issue the L2 prefetches.  No
            // proxy fence per stage: a slot is only refilled after its readers arrived on slot_free,
            // and a fence.proxy.async here would wait for the copies still in flight, collapsing the
            // ring to a single outstanding stage.
            constexpr int NB = 4;
            const int jb = lane / 3, arr = lane - jb * 3;
            const int narr = P.with_x ? 3 : 2;
            for (int k0 = 0; k0 < P.nst; k0 += NB) {
                const int k = k0 + jb;
                const bool mine = lane < 3 * NB && k < P.nst && arr < narr;
                const unsigned kk = kg + (unsigned)k;
                const int slot = (int)(kk % (unsigned)NS);
                const unsigned use = kk / (unsigned)NS;
                const int j0 = k * kG;
                const int cnt = min(kG, P.n - j0);
                const int r0 = dir > 0 ? P.t0 + j0 : P.t0 - j0 - cnt + 1;
                const uint32_t bar = smem_u32(raw_full + slot);
                const uint32_t bytes_e = (uint32_t)cnt * (uint32_t)max_u * 4u;
                const uint32_t bytes_x = P.with_x ? (uint32_t)cnt * (uint32_t)SU * 4u : 0u;
                if (lane == 0) tl_mark(1, (int)kk, 0);
                if (mine && arr == 0) {
                    if (use > 0) mbar_wait_backoff(smem_u32(slot_free + slot), (use - 1) & 1u, (unsigned)p.pf_sleep_ns);
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
        const bool edge_lane = rank == 0 ? lane == 0 : lane == 31;  // the lane without a feeder
        const bool full_u = max_u == 32 * CPL;
        float g = edge_lane ? 0.0f : 1.0f;  // 2^(ex_feeder - ex_mine); all frames start at 0
        // Pipelined re-normalisation: the new frame of stage k+1 is DECIDED during stage k (two
        // shuffles whose results are not needed before the next stage) and APPLIED at the start of
        // stage k+1, so no shuffle latency sits between the rows.
        bool have_dec = false;
        int ex_dec = 0, nbex_dec = 0;
        auto store_scratch = [&](int row) {
            float* r = scr + (size_t)row * SU;
            store_cells<CPL>(r, c0, max_u, v);
            reinterpret_cast<int*>(r)[UP + lane] = ex;
        };
        auto apply_decision = [&]() {
            if (!have_dec) return;
            const int shift = ex - ex_dec;
#pragma unroll
            for (int i = 0; i < CPL; ++i) v[i] = scale_pow2(v[i], shift);
            ex = ex_dec;
            g = edge_lane ? 0.0f : pow2i(max(-126, min(126, nbex_dec - ex_dec)));
        };
        // decision, part 1: magnitudes of this lane and (shuffle in flight) of the feeder's edge cell
        auto decide_1 = [&](int& own, int& nbmag) {
            float mx = v[0];
#pragma unroll
            for (int i = 1; i < CPL; ++i) mx = fmaxf(mx, v[i]);
            own = mx > 0.0f ? ex + ilogb_pos(mx) - kTarget : kNoMass;
            const float edge = rank == 0 ? v[CPL - 1] : v[0];
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
                // hand the last phase-1 state over, then meet the partner
                store_scratch(rank == 0 ? m - 1 : m);
                __threadfence();
                fence_proxy_async();
                timed_cluster_sync();
            }
            const int Pn = ph[phase].n, Pt0 = ph[phase].t0, Pnst = ph[phase].nst;
            int slot = (int)(kg % (unsigned)NS);
            unsigned par = (kg / (unsigned)NS) & 1u;
            for (int k = 0; k < Pnst; ++k) {
                // Fast path: two full stages (16 rows) per barrier round trip.  Waits, frame update and
                // hand-off cost ~350 cycles per round, so amortising them over 16 rows instead of 8
                // matters more than anything else in this warp.
                if (full_u && k + 1 < Pnst && Pn - (k + 1) * kG >= kG) {
                    const int slot2 = slot + 1 == NS ? 0 : slot + 1;
                    const unsigned par2 = slot + 1 == NS ? par ^ 1u : par;
                    tl_mark(0, (int)kg + k, 0);
                    timed_wait(1, smem_u32(prep_full + slot), par);
                    timed_wait(1, smem_u32(prep_full + slot2), par2);
                    tl_mark(0, (int)kg + k, 1);
                    float* spA = ring + (size_t)slot * stage_floats;
                    float* spB = ring + (size_t)slot2 * stage_floats;
                    const int j0 = k * kG;
                    const long long tr0 = p.stats ? clock64() : 0;
                    apply_decision();
                    if (phase == 1) {
                        reinterpret_cast<int*>(spA + off_ve)[lane] = ex;
                        reinterpret_cast<int*>(spB + off_ve)[lane] = ex;
                    }
                    int own = kNoMass, nbmag = kNoMass;
                    auto d1 = [&]() { decide_1(own, nbmag); };
                    auto d2 = [&]() { decide_2(own, nbmag); };
                    const long long tr1 = p.stats ? clock64() : 0;
                    st_post += tr1 - tr0;
                    const int eo = dir > 0 ? 0 : (kG - 1) * max_u;
                    const int istr = dir * max_u;
                    float* st_gA = scr + (size_t)(Pt0 + dir * j0 + (rank == 0 ? 0 : 1)) * SU;
                    float* st_gB = st_gA + (long long)dir * kG * SU;
                    const long long sstr_g = (long long)dir * SU, sstr_s = (long long)max_u;
                    if (rank == 0) {
                        if (phase == 0) {
                            chain_stage<CPL, 0, false, true, 0>(v, g, spA + off_e + eo, spA + off_s + eo, istr, st_gA, sstr_g, ex, UP, lane, c0, max_u, NoHook(), NoHook());
                            chain_stage<CPL, 0, false, true, 0>(v, g, spB + off_e + eo, spB + off_s + eo, istr, st_gB, sstr_g, ex, UP, lane, c0, max_u, d1, d2);
                        } else {
                            chain_stage<CPL, 0, true, true, 0>(v, g, spA + off_e + eo, spA + off_s + eo, istr, spA + off_v, sstr_s, ex, UP, lane, c0, max_u, NoHook(), NoHook());
                            chain_stage<CPL, 0, true, true, 0>(v, g, spB + off_e + eo, spB + off_s + eo, istr, spB + off_v, sstr_s, ex, UP, lane, c0, max_u, d1, d2);
                        }
                    } else {
                        if (phase == 0) {
                            chain_stage<CPL, 1, false, true, 0>(v, g, spA + off_e + eo, spA + off_s + eo, istr, st_gA, sstr_g, ex, UP, lane, c0, max_u, NoHook(), NoHook());
                            chain_stage<CPL, 1, false, true, 0>(v, g, spB + off_e + eo, spB + off_s + eo, istr, st_gB, sstr_g, ex, UP, lane, c0, max_u, d1, d2);
                        } else {
                            chain_stage<CPL, 1, true, true, 0>(v, g, spA + off_e + eo, spA + off_s + eo, istr, spA + off_v, sstr_s, ex, UP, lane, c0, max_u, NoHook(), NoHook());
                            chain_stage<CPL, 1, true, true, 0>(v, g, spB + off_e + eo, spB + off_s + eo, istr, spB + off_v, sstr_s, ex, UP, lane, c0, max_u, d1, d2);
                        }
                    }
                    if (p.stats) {
                        const long long dt = clock64() - tr1;
                        st_prep += dt;
                        st_min = min(st_min, dt); st_max = max(st_max, dt); st_n += 2;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if (phase == 0) {
                            mbar_arrive_relaxed_n(smem_u32(state_full + slot), 1);
                            mbar_arrive_relaxed_n(smem_u32(slot_free + slot), 2);
                            mbar_arrive_relaxed_n(smem_u32(state_full + slot2), 1);
                            mbar_arrive_relaxed_n(smem_u32(slot_free + slot2), 2);
                        } else {
                            mbar_arrive(smem_u32(state_full + slot));
                            mbar_arrive(smem_u32(state_full + slot2));
                        }
                    }
                    tl_mark(0, (int)kg + k, 2);
                    tl_mark(0, (int)kg + k + 1, 2);
                    ++k;
                    for (int z = 0; z < 2; ++z)
                        if (++slot == NS) { slot = 0; par ^= 1u; }
                    continue;
                }
                tl_mark(0, (int)kg + k, 0);
                timed_wait(1, smem_u32(prep_full + slot), par);
                tl_mark(0, (int)kg + k, 1);
                float* sp = ring + (size_t)slot * stage_floats;
                const int j0 = k * kG;
                const int cnt = min(kG, Pn - j0);
                const long long tr0 = p.stats ? clock64() : 0;
                apply_decision();
                if (phase == 1) reinterpret_cast<int*>(sp + off_ve)[lane] = ex;
                int own = kNoMass, nbmag = kNoMass;
                auto d1 = [&]() { decide_1(own, nbmag); };
                auto d2 = [&]() { decide_2(own, nbmag); };
                const long long tr1 = p.stats ? clock64() : 0;
                st_post += tr1 - tr0;  // chain warp: cycles in renorm
                if (p.stats) st_lt400 = min((long long)__popc(__activemask()), st_n == 0 ? 32LL : st_lt400);  // divergence probe
                if (cnt == kG) {
                    const float* e0 = sp + off_e + (dir > 0 ? 0 : (kG - 1) * max_u);
                    const float* s0 = sp + off_s + (dir > 0 ? 0 : (kG - 1) * max_u);
                    const int istr = dir * max_u;
                    // decision from the state after row 5 (part 1) / row 6 (part 2): 2 rows stale when applied.
                    // phase 1 state rows go to scratch (alpha(t) → row t, beta(t+1) → row t+1), phase 2 to shared.
                    float* st_g = scr + (size_t)(Pt0 + dir * j0 + (rank == 0 ? 0 : 1)) * SU;  // global scratch
                    float* st_s = sp + off_v;                                                   // shared state rows
                    const long long sstr_g = (long long)dir * SU, sstr_s = (long long)max_u;
#define SSNT_STAGE_G(RANK_, FULL_) chain_stage<CPL, RANK_, false, FULL_, 0>(v, g, e0, s0, istr, st_g, sstr_g, ex, UP, lane, c0, max_u, d1, d2)
#define SSNT_STAGE_S(RANK_, FULL_) chain_stage<CPL, RANK_, true, FULL_, 0>(v, g, e0, s0, istr, st_s, sstr_s, ex, UP, lane, c0, max_u, d1, d2)
#ifdef SSNT_BF_DEBUG_VARIANTS
                    const int dbg = (p.debug_skip >> 2) & 3;
                    if (dbg == 1) { if (phase == 0) chain_stage<CPL, 0, false, true, 1>(v, g, e0, s0, istr, st_g, sstr_g, ex, UP, lane, c0, max_u, d1, d2); else chain_stage<CPL, 0, true, true, 1>(v, g, e0, s0, istr, st_s, sstr_s, ex, UP, lane, c0, max_u, d1, d2); }
                    else if (dbg == 2) { if (phase == 0) chain_stage<CPL, 0, false, true, 2>(v, g, e0, s0, istr, st_g, sstr_g, ex, UP, lane, c0, max_u, d1, d2); else chain_stage<CPL, 0, true, true, 2>(v, g, e0, s0, istr, st_s, sstr_s, ex, UP, lane, c0, max_u, d1, d2); }
                    else if (dbg == 3) { if (phase == 0) chain_stage<CPL, 0, false, true, 3>(v, g, e0, s0, istr, st_g, sstr_g, ex, UP, lane, c0, max_u, d1, d2); else chain_stage<CPL, 0, true, true, 3>(v, g, e0, s0, istr, st_s, sstr_s, ex, UP, lane, c0, max_u, d1, d2); }
                    else
#endif
                    if (full_u) {
                        if (rank == 0) { if (phase == 0) SSNT_STAGE_G(0, true); else SSNT_STAGE_S(0, true); }
                        else           { if (phase == 0) SSNT_STAGE_G(1, true); else SSNT_STAGE_S(1, true); }
                    } else {
                        if (rank == 0) { if (phase == 0) SSNT_STAGE_G(0, false); else SSNT_STAGE_S(0, false); }
                        else           { if (phase == 0) SSNT_STAGE_G(1, false); else SSNT_STAGE_S(1, false); }
                    }
#undef SSNT_STAGE_G
#undef SSNT_STAGE_S
                } else {
                    for (int q = 0; q < cnt; ++q) {
                        const int t = Pt0 + dir * (j0 + q);
                        const int idx = dir > 0 ? q : cnt - 1 - q;
                        float* st0 = phase == 0 ? scr + (size_t)(rank == 0 ? t : t + 1) * SU : sp + off_v + q * max_u;
                        if (rank == 0) {
                            if (phase == 0) chain_rows<CPL, 0, false, 1, false>(v, g, sp + off_e + idx * max_u, sp + off_s + idx * max_u, 0, st0, 0, ex, UP, lane, c0, max_u);
                            else chain_rows<CPL, 0, true, 1, false>(v, g, sp + off_e + idx * max_u, sp + off_s + idx * max_u, 0, st0, 0, ex, UP, lane, c0, max_u);
                        } else {
                            if (phase == 0) chain_rows<CPL, 1, false, 1, false>(v, g, sp + off_e + idx * max_u, sp + off_s + idx * max_u, 0, st0, 0, ex, UP, lane, c0, max_u);
                            else chain_rows<CPL, 1, true, 1, false>(v, g, sp + off_e + idx * max_u, sp + off_s + idx * max_u, 0, st0, 0, ex, UP, lane, c0, max_u);
                        }
                    }
                    decide_1(own, nbmag);  // short last stage of a phase: decide from the final state
                    decide_2(own, nbmag);
                }
                if (p.stats) {
                    const long long dt = clock64() - tr1;
                    st_prep += dt;  // chain warp: cycles in the rows
                    if (cnt == kG) { st_min = min(st_min, dt); st_max = max(st_max, dt); ++st_n; st_lt800 += dt < 800; }
                }
                __syncwarp();
                if (lane == 0) {
                    if (phase == 0) {
                        // nobody reads state rows in phase 1; the slot is free once its rows were read
                        mbar_arrive_relaxed_n(smem_u32(state_full + slot), 1);
                        mbar_arrive_relaxed_n(smem_u32(slot_free + slot), 2);
                    } else {
                        mbar_arrive(smem_u32(state_full + slot));
                    }
                }
                tl_mark(0, (int)kg + k, 2);
                if (++slot == NS) { slot = 0; par ^= 1u; }
            }
            kg += (unsigned)Pnst;
        }
    } else {
        // ------------------------------- helpers -------------------------------
        // Six helper warps form three pairs; pair p owns stages k = p, p+3, p+6, ... and each warp
        // of the pair owns one half (4 rows) of the stage, processed together for ILP.
        const int h = warp < 4 ? warp - 1 : warp - 2;  // warps 1,2,3,5,6,7 → 0..5
        const int pair = h >> 1, half = h & 1;
        bool me[CPL], ms[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            me[i] = c0 + i < U;
            ms[i] = c0 + i < U - 1;
        }
        unsigned kg = 0;
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == 1) timed_cluster_sync();
            const Phase& P = ph[phase];
            float f_inv_sum = 0.0f;
            int f_M = 0;
            bool f_dead = false, have_ll = false;
            for (int it = pair; it < P.nst + (phase == 1 ? kPairs : 0); it += kPairs) {
                // ---- prep(it): log-probs → probabilities, in place ----
                if (it < P.nst) {
                    const int k = it;
                    const unsigned kk = kg + (unsigned)k;
                    const int slot = slot_of(kk);
                    if (half == 0) tl_mark(2, (int)kk, 0);
                    timed_wait(0, smem_u32(raw_full + slot), use_of(kk) & 1u);
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
                    if (p.debug_skip & 1) {
                    } else if (nr == kHalf) prep_rows<CPL, kHalf>(e0, s0, dir * max_u, t0r, dir, T, me, ms, c0, max_u);
                    else
                        for (int r = 0; r < nr; ++r)
                            prep_rows<CPL, 1>(e0 + r * dir * max_u, s0 + r * dir * max_u, dir * max_u, t0r + dir * r,
                                              dir, T, me, ms, c0, max_u);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(prep_full + slot));
                    if (half == 0) tl_mark(2, (int)kk, 2);
                    if (p.stats) st_prep += clock64() - tp0;
                }
                // ---- post(it - kPairs): gradients (phase 2 only) ----
                if (phase == 1 && it >= kPairs) {
                    const int k = it - kPairs;
                    const unsigned kk = kg + (unsigned)k;
                    const int slot = slot_of(kk);
                    if (half == 0) tl_mark(3, (int)kk, 0);
                    timed_wait(2, smem_u32(state_full + slot), use_of(kk) & 1u);
                    if (half == 0) tl_mark(3, (int)kk, 1);
                    const long long tq0 = p.stats ? clock64() : 0;
                    float* sp = slot_ptr(slot);
                    const int j0 = k * kG;
                    const int cnt = min(kG, P.n - j0);
                    const int q0 = half * kHalf;
                    const int nr = max(0, min(kHalf, cnt - q0));
                    const int ex_state = reinterpret_cast<const int*>(sp + off_ve)[lane];
                    const bool ll_owner = (k == 0 && half == 0);
                    if (!have_ll && !ll_owner) {  // wait for the log-likelihood of the meeting row
                        named_bar_sync(1, 32 * kHelpers);
                        f_M = __float_as_int(llinfo[0]);
                        f_inv_sum = llinfo[1];
                        f_dead = llinfo[2] != 0.0f;
                        have_ll = true;
                    }
                    for (int r = 0; r < nr; ++r) {
                        const int q = q0 + r;
                        const int j = j0 + q;
                        if ((p.debug_skip & 2) && j != 0) continue;
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
                        float vbn_edge = scale_pow2(__shfl_down_sync(kFull, VB[0], 1), exBn - exB);
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
                                        if (c0 + i == U - 1) bad = !(fabsf(g1[i] - 1.0f) < kBfConsistency);
                                }
                                if (rank == 1 && t == 0 && lane == 0) bad = !(fabsf(g1[0] + g2[0] - 1.0f) < kBfConsistency);
                                if (bad) atomicOr(p.status + b, (unsigned)kBfInconsistent);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(slot_free + slot));
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
        o[8] = st_min; o[9] = st_max; o[10] = st_n; o[11] = st_lt400; o[12] = st_lt800;
    }
}

}  // namespace lattice
}  // namespace ssnt
