// Split-role block-floating-point lattice kernel (kernel kind 4): the latency-bound small-batch path.
//
// Why.  At B=32 only 64 of the 148 SMs have a recursion to run, and in fb_bf_kernel the six helper
// warps of a CTA are saturated in the second half of the sweep (MUFU-bound conversion of the
// log-probabilities plus ~40 instructions per row of gradient work on three SM sub-partitions),
// so the recursion warp waits for them.  Here one CLUSTER OF FOUR CTAs owns an utterance:
//
//   rank 0  alpha recursion CTA     rank 2  gradient CTA for the frames t >= m-1
//   rank 1  beta  recursion CTA     rank 3  gradient CTA for the frames t <  m-1       (m = ceil(T/2))
//
//   recursion CTA d:  twelve prep warps (sub-partitions 1-3) load log_emit/log_shift rows in sweep order
//                     straight from global memory, convert them (EX2, length masks) and write the
//                     probabilities into the shared-memory ring; the recursion warp (alone on
//                     sub-partition 0 but for three mostly sleeping publisher warps) sweeps ALL T rows
//                     without phases or barriers and stores its state rows straight to the global scratch
//                     A_d (alpha(t) rows / beta(t+1) rows, lane exponents in the first row of each stage);
//                     the publisher warps fence at GPU scope and flag the rows to the gradient CTAs.
//   gradient CTA:     sixteen identical warps; a frame's occupancies need alpha(t) and beta(t+1), i.e. a
//                     fresh row of one sweep's second half and an old row of the other sweep's first
//                     half; the probabilities are recomputed from the inputs (still in L2).  The
//                     likelihood comes from the meeting row (rank 2 shares it with rank 3).
//
// CTAs signal each other with single words written into the RECEIVER's shared memory
// (st.shared::cluster), so every wait polls local shared memory; the rows themselves travel through
// L2.  No TMA and no mbarrier: every hand-off is a flag word.  The arithmetic (block floating point,
// skewed recursion, in-kernel log-domain re-run of utterances it cannot hold) is fb_bf.cuh's.
#pragma once
#include "fb_bf.cuh"

namespace ssnt {
namespace lattice {

constexpr int kSplitThreads = 512;
constexpr int kSplitWarps = 16;  // helper warps per helper CTA; stage k is owned by warp k % 16
constexpr int kCopyWarps = 3;   // copy-out warps of a recursion CTA (warps 4, 8, 12)
constexpr int kPrepWarps = 12;  // prep warps of a recursion CTA (sub-partitions 1-3)
constexpr int kSplitHeaderBytes = 2048;
constexpr int kPrefetchIters = 2;  // L2 prefetch distance of a prep warp, in its own iterations (x12 stages)
// Poll intervals of the warps that share an SM with a recursion warp: every probe is a shared-memory load
// that competes with the recursion's own loads, stores and shuffles.
#ifndef SSNT_PREP_POLL_NS
#define SSNT_PREP_POLL_NS 256
#endif
#ifndef SSNT_COPY_POLL_NS
#define SSNT_COPY_POLL_NS 128
#endif
constexpr unsigned kPrepPollNs = SSNT_PREP_POLL_NS, kCopyPollNs = SSNT_COPY_POLL_NS;
constexpr int kRoundRing = 128;  // per-round "rows stored" flags kept by the helpers (value r+1 in slot r % 128)

struct SplitParams {
    FbArgs a;
    float* A;          // [B][2][nstp*8][SU]        state rows + lane exponents, sweep order
    unsigned* status;  // [B]
    unsigned* fallbacks;
    int SU, NS, nstp;
    int ring_bytes;    // dynamic shared memory behind the header (sizes the log-domain re-run's own ring)
    int force_fallback;
    int debug;         // profiling aid (SSNT_SPLIT_DEBUG): 1 skip re-normalisation, 2 skip decision hooks (results wrong)
    unsigned* counter;
    long long* stats;
};

__device__ __forceinline__ uint32_t map_to_rank(const void* local_smem, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem)), "r"(rank));
    return r;
}
// Publishes `v` in another CTA's shared memory after everything this thread (and, through the
// preceding __syncwarp, its warp) wrote to global memory.
__device__ __forceinline__ void remote_publish(uint32_t cluster_addr, int v) {
    __threadfence();
    asm volatile("st.relaxed.cluster.shared::cluster.s32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag_ge(const int* f, int need, unsigned sleep_ns = 64) {
    while (flag_load(f) < need) __nanosleep(sleep_ns);
}
template <int CPL>
__device__ __forceinline__ void ldcg_cells(const float* row, int c0, float (&v)[CPL]) {
    if constexpr (CPL >= 4) {
#pragma unroll
        for (int q = 0; q < CPL / 4; ++q) {
            const float4 w = __ldcg(reinterpret_cast<const float4*>(row + c0 + 4 * q));
            v[4 * q + 0] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
        }
    } else {
        const float2 w = __ldcg(reinterpret_cast<const float2*>(row + c0));
        v[0] = w.x; v[1] = w.y;
    }
}

// NSTG full stages (8 rows each, rows in sweep order) as straight-line code.  Slot layout:
// e[8][max_u] | s[8][max_u] | state[8][max_u] | lane exponents.  For the alpha sweep the s rows are
// stored SHIFTED by one token (s'[c] = s[c-1], s'[0] = 0), so one 128-bit load gives a lane the shift
// probability feeding its first cell and those between its cells.
template <int CPL, int RANK, int NSTG, typename H1, typename H2>
__device__ __forceinline__ void chain_round_split(ChainState<CPL>& cs, const float g, float* const (&sp)[4],
                                                  const int lane, H1 h1, H2 h2, float* grow, const int ex) {
    constexpr int max_u = 32 * CPL;
    constexpr int SU = max_u + 32;
    constexpr int NR = NSTG * kG;
    constexpr int CH = CPL <= 4 ? 4 : 2;  // rows per chunk (register budget)
    constexpr int NC = NR / CH;
    const int c0 = lane * CPL;
    float E[2][CH][CPL], S[2][CH][CPL];
    auto load_chunk = [&](int c, int buf) {
#pragma unroll
        for (int r = 0; r < CH; ++r) {
            const int q = c * CH + r;
            const float* base = sp[q >> 3];
            load_cells<CPL>(base + (q & 7) * max_u, c0, max_u, 0.0f, E[buf][r]);
            load_cells<CPL>(base + (kG + (q & 7)) * max_u, c0, max_u, 0.0f, S[buf][r]);
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
            if (RANK == 0) {
                float Ss[CPL];
#pragma unroll
                for (int i = 0; i + 1 < CPL; ++i) Ss[i] = S[c & 1][r][i + 1];
                Ss[CPL - 1] = 0.0f;
                skew_step<CPL, 0>(cs, E[c & 1][r], Ss, S[c & 1][r][0], g, out);
            } else {
                skew_step<CPL, 1>(cs, E[c & 1][r], S[c & 1][r], 0.0f, g, out);
            }
            // the state row goes straight to the global scratch (values + this lane's exponent): keeping it in
            // the ring for a copy-out warp costs a shared store here plus a shared load there, and the
            // recursion SM is bound by its shared-memory pipe
            float* dst = grow + q * SU;
            store_cells<CPL>(dst, c0, max_u, out);
            if ((q & 7) == 0) reinterpret_cast<int*>(dst)[max_u + lane] = ex;  // one exponent row per stage (rounds start at stage boundaries)
            if (q == NR - 3) h1();
            if (q == NR - 2) h2();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Recursion CTA (rank 0 / 1): warp 0 recursion | warps 4, 8, 12 copy-out | the other twelve warps prep
// ---------------------------------------------------------------------------------------------------
template <int CPL>
__device__ void split_chain_cta(const SplitParams& p, int b, unsigned rank, int T, int U, unsigned char* smem_raw) {
    constexpr int max_u = 32 * CPL, SU = max_u + 32;
    constexpr int stageP = 2 * kG * max_u;                     // floats of one probability stage
    constexpr int slot_floats = stageP;                        // e | s
    const FbArgs& a = p.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NS = p.NS;
    const int d = (int)rank, dir = d == 0 ? 1 : -1;
    const int nst = (T + kG - 1) / kG;
    const size_t slab = (size_t)a.max_t * a.max_u;
    const float* le = a.log_emit + (size_t)b * slab;
    const float* ls = a.log_shift + (size_t)b * slab;
    float* Ad = p.A + ((size_t)b * 2 + d) * (size_t)p.nstp * kG * SU;

    int* ready = reinterpret_cast<int*>(smem_raw + 128);          // [NS] prep → recursion: use+1
    int* state_done = reinterpret_cast<int*>(smem_raw + 256);     // [NS] recursion → copy-out: use+1
    int* slot_free = reinterpret_cast<int*>(smem_raw + 384);      // [NS] copy-out → prep: use+1 once the state rows left
    float* ringm = reinterpret_cast<float*>(smem_raw + kSplitHeaderBytes);
    auto slot_ptr = [&](int slot) { return ringm + (size_t)slot * slot_floats; };
    const int c0 = lane * CPL;
    const bool is_copy = warp != 0 && (warp & 3) == 0;             // warps 4, 8, 12
    const int copy_idx = (warp >> 2) - 1;                          // 0..2
    const int prep_idx = warp - 1 - (warp >> 2);                   // warps 1,2,3,5,6,7,9,... → 0..11

    if (warp == 0) {
        // ------------------------------- recursion -------------------------------
        ChainState<CPL> cs;
        cs.init(d, lane, U);
        int ex = 0, nb_ex = 0;
        const bool edge_lane = d == 0 ? lane == 0 : lane == 31;
        float g = edge_lane ? 0.0f : 1.0f;
        bool have_dec = false;
        int ex_dec = 0, nbex_dec = 0;
        auto apply_decision = [&]() {
            if (!have_dec) return;
            const int shift = ex - ex_dec;
#pragma unroll
            for (int i = 0; i < CPL; ++i) cs.a[i] = scale_pow2(cs.a[i], shift);
            const int fshift = nb_ex - nbex_dec;
            cs.inA = scale_pow2(cs.inA, fshift);
            cs.inB = scale_pow2(cs.inB, fshift);
            ex = ex_dec;
            nb_ex = nbex_dec;
            g = edge_lane ? 0.0f : pow2i(max(-126, min(126, nb_ex - ex)));
            have_dec = false;
        };
        auto decide_1 = [&](int& own, int& nbmag) {
            float mx = cs.a[0];
#pragma unroll
            for (int i = 1; i < CPL; ++i) mx = fmaxf(mx, cs.a[i]);
            own = mx > 0.0f ? ex + ilogb_pos(mx) - kTarget : kNoMass;
            const float edge = d == 0 ? cs.a[CPL - 1] : cs.a[0];
            const int amag = edge > 0.0f ? ex + ilogb_pos(edge) : kNoMass;
            nbmag = d == 0 ? __shfl_up_sync(kFull, amag, 1) : __shfl_down_sync(kFull, amag, 1);
        };
        auto decide_2 = [&](int own, int nbmag) {
            if (edge_lane) nbmag = kNoMass;
            int nw = max(own, nbmag - kTarget - kSlack);
            if (nw <= kNoMass / 2) nw = ex;
            ex_dec = nw;
            nbex_dec = d == 0 ? __shfl_up_sync(kFull, nw, 1) : __shfl_down_sync(kFull, nw, 1);
            have_dec = true;
        };
        // NS is a power of two and a multiple of 4, rounds start at multiples of 4 stages: the four
        // slots of a round are consecutive, share one use count, and their flags are one 16-byte word.
        const int lgNS = NS == 16 ? 4 : (NS == 8 ? 3 : 2);
        int4 fl = make_int4(0, 0, 0, 0);  // ready flags of this round's slots, loaded a round ahead
        const long long st_t0 = p.stats ? clock64() : 0;
        long long st_wait = 0, st_first = 0, st_rows = 0, st_hand = 0, st_mid = 0, st_warm = 0;
        // The first pass over the round body is a WARM-UP on whatever the ring holds: the recursion warp
        // would otherwise idle until the first stages are prepared, and then run ~80 lines of cold
        // straight-line code through instruction-cache misses.  Its results are thrown away.
        bool warmup = true;
        int nwarm = 0;
        const ChainState<CPL> cs0 = cs;
        for (int k = 0; k < nst;) {
            // stages of this round: four, but the last four go 2, 1, 1 so that the copy-out and the
            // gradients of the final rows overlap the recursion's last steps (shorter tail)
            // (starting with small rounds as well was measured slower: the recursion then runs into the
            // prep warps' first burst and stalls later)
            const int rem = nst - k;
            int ns = rem > 4 ? 4 : (rem >= 3 ? 2 : 1);
            if (warmup) ns = 4;  // the warm-up pass touches the body every full round runs
            const int slot0 = k & (NS - 1);
            const int use = (k >> lgNS) + 1;
            const int* rflag = ready + (slot0 & ~3);  // aligned group of four flags that holds this round's
            const int fo = slot0 & 3;                 // first flag of the round inside the group (0 or 2, tail only)
            auto round_ready = [&](const int4& f) {
                const int v[4] = {f.x, f.y, f.z, f.w};
                bool ok = true;
#pragma unroll
                for (int z = 0; z < 4; ++z)
                    if (z >= fo && z < fo + ns) ok = ok && v[z] >= use;
                return ok;
            };
            const long long tw0 = p.stats ? clock64() : 0;
            if (!warmup)
                while (!__all_sync(kFull, round_ready(fl))) {
                    asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(fl.x), "=r"(fl.y), "=r"(fl.z), "=r"(fl.w) : "r"(smem_u32(rflag)) : "memory");
                }
            const long long tm0 = p.stats ? clock64() : 0;
            if (p.stats) { st_wait += tm0 - tw0; if (k == 0) st_first = tm0 - st_t0; }
            // request the flags of the next round's group now; they are looked at a round later
            {
                const int kn = k + ns;
                const int* nflag = ready + ((kn & (NS - 1)) & ~3);
                asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(fl.x), "=r"(fl.y), "=r"(fl.z), "=r"(fl.w) : "r"(smem_u32(nflag)) : "memory");
            }
            float* sp[4];
#pragma unroll
            for (int z = 0; z < 4; ++z) sp[z] = slot_ptr((slot0 + z) & (NS - 1));
            if (!(p.debug & 1)) apply_decision();
            int own = kNoMass, nbmag = kNoMass;
            auto d1 = [&]() { if (!(p.debug & 2)) decide_1(own, nbmag); };
            auto d2 = [&]() { if (!(p.debug & 2)) decide_2(own, nbmag); };
            float* grow = Ad + (size_t)k * kG * SU;  // scratch row of the round's first lattice row (sweep order)
            const long long tr1 = p.stats ? clock64() : 0;
            st_mid += tr1 - tm0;
            if (d == 0) {
                if (ns == 4) chain_round_split<CPL, 0, 4>(cs, g, sp, lane, d1, d2, grow, ex);
                else if (ns == 2) chain_round_split<CPL, 0, 2>(cs, g, sp, lane, d1, d2, grow, ex);
                else chain_round_split<CPL, 0, 1>(cs, g, sp, lane, d1, d2, grow, ex);
            } else {
                if (ns == 4) chain_round_split<CPL, 1, 4>(cs, g, sp, lane, d1, d2, grow, ex);
                else if (ns == 2) chain_round_split<CPL, 1, 2>(cs, g, sp, lane, d1, d2, grow, ex);
                else chain_round_split<CPL, 1, 1>(cs, g, sp, lane, d1, d2, grow, ex);
            }
            const long long tr2 = p.stats ? clock64() : 0;
            if (warmup) {  // forget everything the warm-up pass computed and start over
                st_warm = tr2 - st_t0;
                cs = cs0;
                have_dec = false;
                warmup = ++nwarm < 1;
                fl = make_int4(0, 0, 0, 0);
                continue;
            }
            st_rows += tr2 - tr1;
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                if (ns == 4) {
                    asm volatile("st.volatile.shared.v4.s32 [%0], {%1, %1, %1, %1};" ::"r"(smem_u32(state_done + slot0)), "r"(use) : "memory");
                } else {
#pragma unroll
                    for (int z = 0; z < 2; ++z)
                        if (z < ns)
                            asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(state_done + slot0 + z)), "r"(use) : "memory");
                }
            }
            __syncwarp();
            if (p.stats) st_hand += clock64() - tr2;
            k += ns;
        }
        if (p.stats && lane == 0) {
            long long* o = p.stats + ((size_t)blockIdx.x * 16 + warp) * 16;
            o[0] = clock64() - st_t0; o[1] = st_wait; o[2] = st_first; o[3] = st_rows; o[4] = st_hand; o[5] = st_mid; o[6] = st_warm;
        }
    } else if (!is_copy) {
        // ------------------------------- prep -------------------------------
        // prep warp i owns the stages k = i, i+12, ...: rows of the sweep in sweep order, converted to
        // probabilities with the length masks (tokens >= U: e = s = 0; last token / last frame: s = 0;
        // rows past T: 0, the recursion runs whole stages).
        bool me[CPL], ms[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            me[i] = c0 + i < U;
            ms[i] = c0 + i < U - 1;
        }
        const long long st_t0 = p.stats ? clock64() : 0;
        long long st_w1 = 0, st_first = 0;
        long long st_tl[4] = {0, 0, 0, 0};
        bool first = true;
        for (int k = prep_idx; k < nst; k += kPrepWarps) {
            const int slot = k % NS;
            // inputs of the stage this warp prepares kPrefetchIters iterations from now → L2, one 128-byte
            // line per lane (plain prefetch instructions: cp.async.bulk.prefetch costs hundreds of cycles
            // to ISSUE and serialises across the twelve prep warps — measured 6400 cycles at start-up)
            {
                const int kf = k + kPrefetchIters * kPrepWarps;
                if (kf < nst) {
                    const int j0 = kf * kG, n = min(kG, T - j0);
                    const int r0 = dir > 0 ? j0 : T - j0 - n;
                    const int nlines = n * max_u / 32;  // 128-byte lines of the n rows of one tensor
                    for (int l = lane; l < nlines; l += 32) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(le + (size_t)r0 * max_u + l * 32));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ls + (size_t)r0 * max_u + l * 32));
                    }
                }
            }
            constexpr int HR = kG / 2;  // rows per half stage
            if (k >= NS) {  // the slot's previous occupant has been consumed and its state rows copied out
                const long long t0 = p.stats ? clock64() : 0;
                wait_flag_ge(state_done + slot, k / NS, kPrepPollNs);
                if (p.stats) st_w1 += clock64() - t0;
            }
            float* dst = slot_ptr(slot);
            float RE[2][HR][CPL], RS[2][HR][CPL];
            auto load_half = [&](int h, int buf) {
#pragma unroll
                for (int q = 0; q < HR; ++q) {
                    const int j = k * kG + h * HR + q;
                    const int t = dir > 0 ? j : T - 1 - j;
                    if (j < T) {
                        ldcg_cells<CPL>(le + (size_t)t * max_u, c0, RE[buf][q]);
                        ldcg_cells<CPL>(ls + (size_t)t * max_u, c0, RS[buf][q]);
                    } else {
#pragma unroll
                        for (int i = 0; i < CPL; ++i) { RE[buf][q][i] = -INFINITY; RS[buf][q][i] = -INFINITY; }
                    }
                }
            };
            if (p.stats && first) st_tl[0] = clock64() - st_t0;
            load_half(0, 0);
            load_half(1, 1);
            if (p.stats && first) {
                st_tl[1] = clock64() - st_t0;
                if (RS[1][HR - 1][CPL - 1] == 123456.0f && RE[1][HR - 1][0] == 654321.0f) st_w1 = 1;
                st_tl[2] = clock64() - st_t0;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int q = 0; q < HR; ++q) {
                    const int j = k * kG + h * HR + q;
                    const int t = dir > 0 ? j : T - 1 - j;
                    const bool not_last = t != T - 1;
                    float Ev[CPL], Sv[CPL];
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        // both halves are resident in registers; the select keeps the body loop-invariant
                        const float re = h == 0 ? RE[0][q][i] : RE[1][q][i];
                        const float rs = h == 0 ? RS[0][q][i] : RS[1][q][i];
                        Ev[i] = me[i] ? ex2(re * kLog2e) : 0.0f;
                        Sv[i] = (ms[i] && not_last) ? ex2(rs * kLog2e) : 0.0f;
                    }
                    const int row = h * HR + q;
                    store_cells<CPL>(dst + row * max_u, c0, max_u, Ev);
                    if (d == 0) {  // alpha sweep: s'[c] = s[c-1] (see chain_round_split)
                        float prev = __shfl_up_sync(kFull, Sv[CPL - 1], 1);
                        if (lane == 0) prev = 0.0f;
                        float Sh[CPL];
                        Sh[0] = prev;
#pragma unroll
                        for (int i = 1; i < CPL; ++i) Sh[i] = Sv[i - 1];
                        store_cells<CPL>(dst + (kG + row) * max_u, c0, max_u, Sh);
                    } else {
                        store_cells<CPL>(dst + (kG + row) * max_u, c0, max_u, Sv);
                    }
                }
            }
            __syncwarp();
            if (p.stats && first) st_tl[3] = clock64() - st_t0;
            if (lane == 0) flag_publish(ready + slot, k / NS + 1);
            if (p.stats && first) st_first = clock64() - st_t0;
            first = false;
        }
        if (p.stats && lane == 0) {
            long long* o = p.stats + ((size_t)blockIdx.x * 16 + warp) * 16;
            o[0] = clock64() - st_t0; o[1] = st_w1; o[2] = st_first; o[3] = st_tl[0]; o[4] = st_tl[1]; o[5] = st_tl[2]; o[6] = st_tl[3];
        }
    } else {
        // ------------------------------- copy-out -------------------------------
        // warp c takes the rounds r = c, c+3, ...; a round is the (one or two) stages the recursion hands
        // over together.  The GPU-scope fence each round costs ~1000 cycles, hence several warps.
        const uint32_t h0 = map_to_rank(smem_raw + 1024, 2), h1 = map_to_rank(smem_raw + 1024, 3);
        const int nround = (nst + 1) / 2;
        const long long st_t0 = p.stats ? clock64() : 0;
        long long st_w1 = 0, st_f = 0;
        for (int r = copy_idx; r < nround; r += kCopyWarps) {
            const int kend = min(2 * r + 2, nst);
            for (int k = 2 * r; k < kend; ++k) {
                const int slot = k % NS;
                const long long t0 = p.stats ? clock64() : 0;
                wait_flag_ge(state_done + slot, k / NS + 1, kCopyPollNs);
                if (p.stats) st_w1 += clock64() - t0;
            }
            // The recursion warp stored the rows itself and released them at CTA scope (fence + flag); the
            // GPU-scope fence below, executed after observing that flag, orders them before the flags this
            // warp writes into the gradient CTAs (causality is transitive in the PTX memory model).
            if (lane == 0) {
                const long long t0 = p.stats ? clock64() : 0;
                __threadfence();
                if (p.stats) st_f += clock64() - t0;
                // round_done[d][r % 128] = r + 1 in both helper CTAs
                const uint32_t off = 4u * (uint32_t)(d * kRoundRing + (r % kRoundRing));
                asm volatile("st.relaxed.cluster.shared::cluster.s32 [%0], %1;" ::"r"(h0 + off), "r"(r + 1) : "memory");
                asm volatile("st.relaxed.cluster.shared::cluster.s32 [%0], %1;" ::"r"(h1 + off), "r"(r + 1) : "memory");
            }
        }
        if (p.stats && lane == 0) {
            long long* o = p.stats + ((size_t)blockIdx.x * 16 + warp) * 16;
            o[0] = clock64() - st_t0; o[1] = st_w1; o[2] = st_f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Gradient CTA (rank 2 / 3): sixteen identical warps; warp w owns the stages k = w, w+16, ... of its half
// ---------------------------------------------------------------------------------------------------
template <int CPL>
__device__ void split_helper_cta(const SplitParams& p, int b, unsigned rank, int T, int U, unsigned char* smem_raw) {
    constexpr int max_u = 32 * CPL, SU = max_u + 32;
    constexpr int GB = CPL <= 4 ? 4 : 2;  // gradient rows per batch (all loads of a batch are issued before any use)
    const FbArgs& a = p.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = (int)rank - 2, dir = d == 0 ? 1 : -1;
    const int nst = (T + kG - 1) / kG;
    const int m = (T + 1) >> 1;
    const int n1 = d == 0 ? m - 1 : T - m + 1;   // first sweep row (of direction d) whose gradients this CTA emits
    const size_t slab = (size_t)a.max_t * a.max_u;
    const float* le = a.log_emit + (size_t)b * slab;
    const float* ls = a.log_shift + (size_t)b * slab;
    float* ge = a.grad_emit + (size_t)b * slab;
    float* gs = a.grad_shift + (size_t)b * slab;
    const float* A0 = p.A + ((size_t)b * 2 + 0) * (size_t)p.nstp * kG * SU;  // row j = alpha(j)
    const float* A1 = p.A + ((size_t)b * 2 + 1) * (size_t)p.nstp * kG * SU;  // row j = beta(T - j)
    const int* round_done = reinterpret_cast<const int*>(smem_raw + 1024);  // [2][128], written by the recursion CTAs
    int* ll_flag = reinterpret_cast<int*>(smem_raw + 704);       // [0] 1 once llinfo is valid
    float* llinfo = reinterpret_cast<float*>(smem_raw + 720);    // [0] M (int bits) [1] 1/sum [2] dead
    const int c0 = lane * CPL;

    // all sweep rows < n of direction `dd` are in global memory? (rows come in rounds of 16)
    auto rows_ready = [&](int dd, int n) {
        if (n <= 0) return true;
        const int r = (n - 1) / (2 * kG);  // the round that holds row n-1
        return flag_load(round_done + dd * kRoundRing + (r % kRoundRing)) >= r + 1;
    };
    auto range_ready = [&](int dd, int lo, int hi) {  // rows [lo, hi): at most two rounds
        if (hi <= lo) return true;
        return rows_ready(dd, lo + 1) && rows_ready(dd, hi);
    };

    bool me[CPL], ms[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        me[i] = c0 + i < U;
        ms[i] = c0 + i < U - 1;
    }
    float f_inv_sum = 0.0f;
    int f_M = 0;
    bool f_dead = false, have_ll = false;
    const long long st_t0 = p.stats ? clock64() : 0;
    long long st_grad = 0, st_grad_start = 0, st_wait = 0;

    // work units of GB consecutive sweep rows, starting at the first row of this CTA's half; warp w takes
    // the units w, w+16, ... (small units keep the tail after the recursion's last row short)
    const int nunits = (T - n1 + GB - 1) / GB;
    for (int un = warp; un < nunits; un += kSplitWarps) {
        const int j_lo = n1 + un * GB;
        const int j_hi = min(j_lo + GB, T);
        const bool ll_producer = d == 0 && un == 0;  // this warp computes the likelihood itself
        {
            // own direction: rows [j_lo, j_hi); other direction: rows T-1-t, i.e. [T - j_hi, T - j_lo)
            const long long t0 = p.stats ? clock64() : 0;
            while (!(range_ready(d, j_lo, j_hi) && range_ready(1 - d, T - j_hi, T - j_lo))) __nanosleep(200);
            if (!have_ll && !ll_producer) {
                wait_flag_ge(ll_flag, 1, 200);
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
                f_M = __float_as_int(*reinterpret_cast<volatile float*>(llinfo + 0));
                f_inv_sum = *reinterpret_cast<volatile float*>(llinfo + 1);
                f_dead = *reinterpret_cast<volatile float*>(llinfo + 2) != 0.0f;
                have_ll = true;
            }
            if (p.stats) st_wait += clock64() - t0;
        }
        // The rows behind the flags were fenced at GPU scope by their writers before the flags were set;
        // they are read with ld.global.cg (L2), control-dependent on the flag values.
        const long long tg0 = p.stats ? clock64() : 0;
        if (p.stats && st_grad_start == 0) st_grad_start = tg0 - st_t0;
        {
            const int jb = j_lo;
            float E[GB][CPL], S[GB][CPL], VA[GB][CPL], VB[GB][CPL];
            int exA[GB], exB[GB];
#pragma unroll
            for (int r = 0; r < GB; ++r) {
                const int j = min(jb + r, j_hi - 1);  // rows past the end repeat the last one (not stored)
                const int t = dir > 0 ? j : T - 1 - j;
                const float* arow = A0 + (size_t)t * SU;            // alpha(t)
                const float* brow = A1 + (size_t)(T - 1 - t) * SU;  // beta(t+1)
                ldcg_cells<CPL>(le + (size_t)t * max_u, c0, E[r]);
                ldcg_cells<CPL>(ls + (size_t)t * max_u, c0, S[r]);
                ldcg_cells<CPL>(arow, c0, VA[r]);
                ldcg_cells<CPL>(brow, c0, VB[r]);
                // lane exponents: kept in the first row of each 8-row stage of the sweep that wrote the row
                exA[r] = __ldcg(reinterpret_cast<const int*>(A0 + (size_t)(t & ~7) * SU) + max_u + lane);
                exB[r] = __ldcg(reinterpret_cast<const int*>(A1 + (size_t)((T - 1 - t) & ~7) * SU) + max_u + lane);
            }
            float edge[GB];
            int exBn[GB];
#pragma unroll
            for (int r = 0; r < GB; ++r) {
                const int j = min(jb + r, j_hi - 1);
                const int t = dir > 0 ? j : T - 1 - j;
                const bool not_last = t != T - 1;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    E[r][i] = me[i] ? ex2(E[r][i] * kLog2e) : 0.0f;
                    S[r][i] = (ms[i] && not_last) ? ex2(S[r][i] * kLog2e) : 0.0f;
                }
                exBn[r] = __shfl_down_sync(kFull, exB[r], 1);
                edge[r] = __shfl_down_sync(kFull, VB[r][0], 1);
            }
#pragma unroll
            for (int r = 0; r < GB; ++r) {
                const int j = jb + r;
                if (j >= j_hi) break;  // warp-uniform
                const int t = dir > 0 ? j : T - 1 - j;
                float vbn_edge = scale_pow2(edge[r], exBn[r] - exB[r]);
                if (lane == 31) vbn_edge = 0.0f;
                const int EL = exA[r] + exB[r];
                if (d == 0 && t == m - 1) {
                    // likelihood from the meeting row: Z = sum_u alpha(m-1,u) (e beta(m,u) + s beta(m,u+1))
                    float w = 0.0f;
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        const float nb = (i + 1 < CPL) ? VB[r][i + 1] : vbn_edge;
                        w += VA[r][i] * (E[r][i] * VB[r][i] + S[r][i] * nb);
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
                        if (st) atomicOr(p.status + b, st);
                        const double ll2 = (double)lg2(sum) + (double)M;
                        a.log_likelihood[b] = st ? -INFINITY : (float)(ll2 * kLn2);
                        // share with this CTA's other warps and with the other gradient CTA
                        llinfo[0] = __int_as_float(M);
                        llinfo[1] = f_inv_sum;
                        llinfo[2] = f_dead ? 1.0f : 0.0f;
                        __threadfence_block();
                        asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(ll_flag)), "r"(1) : "memory");
                        const uint32_t rl = map_to_rank(llinfo, 3), rf = map_to_rank(ll_flag, 3);
                        asm volatile("st.relaxed.cluster.shared::cluster.s32 [%0], %1;" ::"r"(rl), "r"(M) : "memory");
                        asm volatile("st.relaxed.cluster.shared::cluster.f32 [%0], %1;" ::"r"(rl + 4u), "f"(f_inv_sum) : "memory");
                        asm volatile("st.relaxed.cluster.shared::cluster.f32 [%0], %1;" ::"r"(rl + 8u), "f"(f_dead ? 1.0f : 0.0f) : "memory");
                        asm volatile("fence.acq_rel.cluster;" ::: "memory");
                        asm volatile("st.relaxed.cluster.shared::cluster.s32 [%0], %1;" ::"r"(rf), "r"(1) : "memory");
                    }
                }
                // gamma = alpha·p·beta / Z, exponents split over two factors
                const int kf = max(-252, min(252, EL - f_M));
                const int kh = kf >> 1;
                const float fa = pow2i(max(-126, kh));
                const float fb = pow2i(max(-126, kf - kh)) * f_inv_sum;
                float g1[CPL], g2[CPL];
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const float nb = (i + 1 < CPL) ? VB[r][i + 1] : vbn_edge;
                    const float va = VA[r][i] * fa;
                    g1[i] = f_dead ? 0.0f : va * ((E[r][i] * VB[r][i]) * fb);
                    g2[i] = f_dead ? 0.0f : va * ((S[r][i] * nb) * fb);
                }
                store_cells_cs<CPL>(ge + (size_t)t * max_u, c0, max_u, g1);
                store_cells_cs<CPL>(gs + (size_t)t * max_u, c0, max_u, g2);
                if (!f_dead && (t == T - 1 || t == 0)) {
                    bool bad = false;
                    if (d == 0 && t == T - 1) {
#pragma unroll
                        for (int i = 0; i < CPL; ++i)
                            if (c0 + i == U - 1) bad = !(fabsf(g1[i] - 1.0f) < kBfConsistency);
                    }
                    if (d == 1 && t == 0 && lane == 0) bad = !(fabsf(g1[0] + g2[0] - 1.0f) < kBfConsistency);
                    if (bad) atomicOr(p.status + b, (unsigned)kBfInconsistent);
                }
            }
        }
        if (p.stats) st_grad += clock64() - tg0;
    }
    if (p.stats && lane == 0) {
        long long* o = p.stats + ((size_t)blockIdx.x * 16 + warp) * 16;
        o[0] = clock64() - st_t0; o[3] = st_grad; o[5] = st_grad_start; o[6] = st_wait;
    }
}

}  // namespace lattice
}  // namespace ssnt
