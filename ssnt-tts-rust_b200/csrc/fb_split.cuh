// Split-role block-floating-point lattice kernel (kernel kind 4): the latency-bound small-batch path.
//
// Why.  At B=32 only 64 of the 148 SMs have a recursion to run, and inside fb_bf_kernel the
// recursion warp shares its SM's memory-instruction pipeline with the helper warps that convert
// log-probabilities (MUFU-bound) and emit gradients: measured 57-78 cycles/row in situ against 30
// in isolation (tools/skew_microbench2.cu).  Here one CLUSTER OF FOUR CTAs owns an utterance and
// the recursion SMs run nothing but the recursion:
//
//   rank 0  alpha recursion CTA     rank 2  alpha-side helper CTA
//   rank 1  beta  recursion CTA     rank 3  beta-side  helper CTA
//
//   helper d:   log_emit/log_shift rows (global, in sweep order) --EX2, masks--> probability stages
//               P_d (global scratch ring, L2 resident) ............. ahead of the recursion
//   chain d:    P_d --TMA--> shared ring --recursion warp--> state rows (shared) --copy-out warp-->
//               A_d (global scratch: alpha(t) rows / beta(t+1) rows with their lane exponents)
//   helper d:   gradients of the rows its recursion walks in the SECOND half of its sweep, from the
//               fresh A_d rows, the other direction's first-half rows and the P_d stages it wrote
//               itself; the likelihood comes from the meeting row (helper 0 shares it with helper 1).
//
// Both recursions sweep all T rows without ever stopping: there are no phases and no cluster
// barrier on the critical path.  CTAs signal each other with single words written into the
// RECEIVER's shared memory (st.shared::cluster), so every wait polls local shared memory; the data
// itself travels through L2.  The arithmetic (block floating point, skewed recursion, in-kernel
// log-domain re-run of utterances it cannot hold) is fb_bf.cuh's.
#pragma once
#include "fb_bf.cuh"

namespace ssnt {
namespace lattice {

constexpr int kSplitThreads = 512;
constexpr int kSplitWarps = 16;  // helper warps per helper CTA; stage k is owned by warp k % 16
constexpr int kCopyWarps = 4;   // copy-out warps of a recursion CTA (warps 3..6)
constexpr int kSplitHeaderBytes = 2048;
constexpr int kPrefetchIters = 2;  // L2 prefetch distance of a helper warp, in its own iterations (x16 stages)
constexpr int kRoundRing = 128;  // per-round "rows stored" flags kept by the helpers (value r+1 in slot r % 128)

struct SplitParams {
    FbArgs a;
    float* P;          // [B][2][ring][2*8*max_u]   probability stages (e rows | s rows), sweep order
    float* A;          // [B][2][nstp*8][SU]        state rows + lane exponents, sweep order
    float* log_scratch;  // [B][max_t+1][SU]        only touched by the log-domain re-run
    unsigned* status;  // [B]
    unsigned* fallbacks;
    int SU, NS, ring, nstp;
    int force_fallback;
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

// ---------------------------------------------------------------------------------------------------
// Recursion CTA (rank 0 / 1): warp 0 recursion | warp 1 loader (TMA) | warp 2 notifier | warp 3 copy-out
// ---------------------------------------------------------------------------------------------------
template <int CPL>
__device__ void split_chain_cta(const SplitParams& p, int b, unsigned rank, int T, int U, unsigned char* smem_raw) {
    constexpr int max_u = 32 * CPL, SU = max_u + 32;
    constexpr int stageP = 2 * kG * max_u;                     // floats of one probability stage
    constexpr int slot_floats = stageP + kG * max_u + 32;      // e | s | state rows | lane exponents
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NS = p.NS, ring = p.ring;
    const int d = (int)rank;
    const int nst = (T + kG - 1) / kG;
    const float* Pd = p.P + ((size_t)b * 2 + d) * (size_t)ring * stageP;
    float* Ad = p.A + ((size_t)b * 2 + d) * (size_t)p.nstp * kG * SU;

    uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem_raw);   // [NS] TMA completion
    int* ready = reinterpret_cast<int*>(smem_raw + 128);          // [NS] notifier → recursion: use+1
    int* state_done = reinterpret_cast<int*>(smem_raw + 256);     // [NS] recursion → copy-out: use+1
    int* slot_free = reinterpret_cast<int*>(smem_raw + 384);      // [NS] copy-out → loader: use+1 once the state rows left
    int* prep_done = reinterpret_cast<int*>(smem_raw + 512);      // [8] written by helper d's warps (remote)
    float* ringm = reinterpret_cast<float*>(smem_raw + kSplitHeaderBytes);
    auto slot_ptr = [&](int slot) { return ringm + (size_t)slot * slot_floats; };
    const int c0 = lane * CPL;

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
        int slot = 0, use1 = 1;
        int fA = 0, fB = 0;  // ready flags of this round's slots, loaded a round ahead
        const long long st_t0 = p.stats ? clock64() : 0;
        long long st_wait = 0, st_first = 0, st_rows = 0, st_hand = 0;
        for (int k = 0; k < nst;) {
            const bool two = nst - k >= 2;
            const int slot2 = slot + 1 == NS ? 0 : slot + 1;
            const int use2 = slot + 1 == NS ? use1 + 1 : use1;
            const long long tw0 = p.stats ? clock64() : 0;
            while (!__all_sync(kFull, fA >= use1)) fA = flag_load(ready + slot);
            if (two)
                while (!__all_sync(kFull, fB >= use2)) fB = flag_load(ready + slot2);
            if (p.stats) { const long long t = clock64(); st_wait += t - tw0; if (k == 0) st_first = t - st_t0; }
            const int adv = two ? 2 : 1;
            int slot_n = slot, use_n = use1;
            for (int z = 0; z < adv; ++z)
                if (++slot_n == NS) { slot_n = 0; ++use_n; }
            fA = flag_load(ready + slot_n);
            fB = flag_load(ready + (slot_n + 1 == NS ? 0 : slot_n + 1));
            float* spA = slot_ptr(slot);
            float* spB = slot_ptr(slot2);
            apply_decision();
            reinterpret_cast<int*>(spA + stageP + kG * max_u)[lane] = ex;
            if (two) reinterpret_cast<int*>(spB + stageP + kG * max_u)[lane] = ex;
            int own = kNoMass, nbmag = kNoMass;
            auto d1 = [&]() { decide_1(own, nbmag); };
            auto d2 = [&]() { decide_2(own, nbmag); };
            // stages hold their rows in SWEEP order for both directions, so both walk them forwards
            const long long tr1 = p.stats ? clock64() : 0;
            if (two) {
                if (d == 0) chain_round_skew<CPL, 0, true, 16, decltype(d1), decltype(d2), true>(cs, g, spA, spA + kG * max_u, spA + stageP, ex, lane, d1, d2, spB, spB + kG * max_u, spB + stageP);
                else chain_round_skew<CPL, 1, true, 16, decltype(d1), decltype(d2), true>(cs, g, spA, spA + kG * max_u, spA + stageP, ex, lane, d1, d2, spB, spB + kG * max_u, spB + stageP);
            } else {
                if (d == 0) chain_round_skew<CPL, 0, true, 8, decltype(d1), decltype(d2), true>(cs, g, spA, spA + kG * max_u, spA + stageP, ex, lane, d1, d2);
                else chain_round_skew<CPL, 1, true, 8, decltype(d1), decltype(d2), true>(cs, g, spA, spA + kG * max_u, spA + stageP, ex, lane, d1, d2);
            }
            const long long tr2 = p.stats ? clock64() : 0;
            st_rows += tr2 - tr1;
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(state_done + slot)), "r"(use1) : "memory");
                if (two)
                    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(state_done + slot2)), "r"(use2) : "memory");
            }
            __syncwarp();
            if (p.stats) st_hand += clock64() - tr2;
            k += adv;
            slot = slot_n;
            use1 = use_n;
        }
        if (p.stats && lane == 0) {
            long long* o = p.stats + ((size_t)blockIdx.x * 16 + warp) * 16;
            o[0] = clock64() - st_t0; o[1] = st_wait; o[2] = st_first; o[3] = st_rows; o[4] = st_hand;
        }
    } else if (warp == 1) {
        // ------------------------------- loader -------------------------------
        // four lanes, four stages per cp.async.bulk instruction (the issue cost is per instruction)
        const long long st_t0 = p.stats ? clock64() : 0;
        long long st_w1 = 0, st_w2 = 0;
        for (int k0 = 0; k0 < nst; k0 += 4) {
            const int k = k0 + lane;
            const bool mine = lane < 4 && k < nst;
            const int slot = k % NS;
            const uint32_t bar = smem_u32(raw_full + slot);
            if (mine) {
                const long long t0 = p.stats ? clock64() : 0;
                if (k >= NS) wait_flag_ge(slot_free + slot, k / NS);          // previous occupant copied out
                const long long t1 = p.stats ? clock64() : 0;
                wait_flag_ge(prep_done + (k % kSplitWarps), k / kSplitWarps + 1);  // helper warp k%16 finished stage k
                if (p.stats) { st_w1 += t1 - t0; st_w2 += clock64() - t1; }
                // the stage was written with generic-proxy stores and fenced at GPU scope before the flag;
                // it is read by the async proxy, from L2
                fence_proxy_async();
                mbar_expect_tx(bar, (uint32_t)stageP * 4u);
            }
            __syncwarp();
            if (mine) bulk_g2s(smem_u32(slot_ptr(slot)), Pd + (size_t)(k % ring) * stageP, (uint32_t)stageP * 4u, bar);
        }
        if (p.stats && lane == 0) {
            long long* o = p.stats + ((size_t)blockIdx.x * 16 + warp) * 16;
            o[0] = clock64() - st_t0; o[1] = st_w1; o[2] = st_w2;
        }
    } else if (warp == 2) {
        // ------------------------------- notifier -------------------------------
        for (int k = 0; k < nst; ++k) {
            const int slot = k % NS;
            mbar_wait_warp(smem_u32(raw_full + slot), (unsigned)(k / NS) & 1u);
            if (lane == 0) flag_publish(ready + slot, k / NS + 1);
        }
    } else if (warp < 3 + kCopyWarps) {
        // ------------------------------- copy-out -------------------------------
        // warp c takes the rounds r = c-3, c-3+4, ...; a round is the (one or two) stages the recursion
        // hands over together.  The global fence each round costs ~700 cycles, hence several warps.
        const uint32_t h0 = map_to_rank(smem_raw + 1024, 2), h1 = map_to_rank(smem_raw + 1024, 3);
        const int nround = (nst + 1) / 2;
        const long long st_t0 = p.stats ? clock64() : 0;
        long long st_w1 = 0, st_f = 0;
        for (int r = warp - 3; r < nround; r += kCopyWarps) {
            const int kend = min(2 * r + 2, nst);
            for (int k = 2 * r; k < kend; ++k) {
                const int slot = k % NS;
                const long long t0 = p.stats ? clock64() : 0;
                wait_flag_ge(state_done + slot, k / NS + 1, 32);
                if (p.stats) st_w1 += clock64() - t0;
                const float* sp = slot_ptr(slot);
                const int exs = reinterpret_cast<const int*>(sp + stageP + kG * max_u)[lane];
                copy_out_rows<CPL, kG, true>(sp + stageP, Ad + (size_t)k * kG * SU, (long long)SU, exs, c0, max_u, max_u, lane);
                __syncwarp();
                if (lane == 0) asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(slot_free + slot)), "r"(k / NS + 1) : "memory");
            }
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
// Helper CTA (rank 2 / 3): eight identical warps; warp w owns the stages k = w, w+8, ...
// ---------------------------------------------------------------------------------------------------
template <int CPL>
__device__ void split_helper_cta(const SplitParams& p, int b, unsigned rank, int T, int U, unsigned char* smem_raw) {
    constexpr int max_u = 32 * CPL, SU = max_u + 32;
    constexpr int stageP = 2 * kG * max_u;
    constexpr int GB = CPL <= 4 ? 4 : 2;  // gradient rows per batch (all loads of a batch are issued before any use)
    const FbArgs& a = p.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = (int)rank - 2, dir = d == 0 ? 1 : -1;
    const int ring = p.ring;
    const int nst = (T + kG - 1) / kG;
    const int m = (T + 1) >> 1;
    const int n1 = d == 0 ? m - 1 : T - m + 1;   // first sweep row whose gradients this helper emits
    const int k1 = n1 / kG;                      // … and the stage that holds it
    const size_t slab = (size_t)a.max_t * a.max_u;
    const float* le = a.log_emit + (size_t)b * slab;
    const float* ls = a.log_shift + (size_t)b * slab;
    float* ge = a.grad_emit + (size_t)b * slab;
    float* gs = a.grad_shift + (size_t)b * slab;
    float* Pd = p.P + ((size_t)b * 2 + d) * (size_t)ring * stageP;
    const float* A0 = p.A + ((size_t)b * 2 + 0) * (size_t)p.nstp * kG * SU;  // row j = alpha(j)
    const float* A1 = p.A + ((size_t)b * 2 + 1) * (size_t)p.nstp * kG * SU;  // row j = beta(T - j)
    const int* round_done = reinterpret_cast<const int*>(smem_raw + 1024);  // [2][128], written by the recursion CTAs
    int* ll_flag = reinterpret_cast<int*>(smem_raw + 704);       // [0] 1 once llinfo is valid
    float* llinfo = reinterpret_cast<float*>(smem_raw + 720);    // [0] M (int bits) [1] 1/sum [2] dead
    const int c0 = lane * CPL;
    const uint32_t chain_prep_done = map_to_rank(smem_raw + 512, (unsigned)d) + 4u * warp;

    // all sweep rows < n of direction `dd` are in global memory? (rows come in rounds of 16)
    auto rows_ready = [&](int dd, int n) {
        if (n <= 0) return true;
        const int r = (n - 1) / (2 * kG);  // the round that holds row n-1; earlier rounds are checked by whoever needs them
        return flag_load(round_done + dd * kRoundRing + (r % kRoundRing)) >= r + 1;
    };
    // rows [lo, hi) of direction dd
    auto range_ready = [&](int dd, int lo, int hi) {
        if (hi <= lo) return true;
        return rows_ready(dd, lo + 1) && rows_ready(dd, hi);  // a range of <= 16 rows touches at most two rounds
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

    // raw log-prob rows of the next stage to prepare, requested one iteration ahead
    float RE[kG][CPL], RS[kG][CPL];
    auto load_raw = [&](int k) {
#pragma unroll
        for (int q = 0; q < kG; ++q) {
            const int j = k * kG + q;
            const int t = dir > 0 ? j : T - 1 - j;
            if (j < T) {
                ldcg_cells<CPL>(le + (size_t)t * max_u, c0, RE[q]);
                ldcg_cells<CPL>(ls + (size_t)t * max_u, c0, RS[q]);
            } else {
#pragma unroll
                for (int i = 0; i < CPL; ++i) { RE[q][i] = -INFINITY; RS[q][i] = -INFINITY; }
            }
        }
    };

    int kp = warp;                                    // next stage to prepare
    int kq = k1 + ((warp - k1) % kSplitWarps + kSplitWarps) % kSplitWarps;  // next stage whose gradients to emit (≡ warp mod 8)
    int nprep = 0;
    bool pending = false;  // a prepared stage whose publication is still owed
    const long long st_t0 = p.stats ? clock64() : 0;
    long long st_tl[5] = {0, 0, 0, 0, 0};
    long long st_prep = 0, st_fence = 0, st_grad = 0, st_prep_end = 0, st_grad_start = 0, st_load = 0, st_first_pub = 0;
    while (kp < nst || kq < nst) {
        // ---- prep(kp): has priority, the recursion waits for it ----
        bool can_prep = kp < nst;
        if (can_prep && kp >= ring) {
            // previous occupant of the ring slot (this warp owns it: ring % 8 == 0): the recursion has
            // consumed it (its round is copied out) and this warp has emitted its gradients
            const int old = kp - ring;
            can_prep = rows_ready(d, (old + 1) * kG) && (old < k1 || kq > old);
        }
        if (can_prep) {
            const long long tp0 = p.stats ? clock64() : 0;
            float* dst = Pd + (size_t)(kp % ring) * stageP;
            // keep the inputs of the stages this warp will prepare next on their way into L2 (the very
            // first demand loads go first: the recursion is waiting for them)
            if (nprep == 0) load_raw(kp);
            if (lane < (nprep == 0 ? kPrefetchIters : 1)) {
                const int kf = kp + (nprep == 0 ? lane + 1 : kPrefetchIters) * kSplitWarps;
                if (kf < nst) {
                    const int j0 = kf * kG, n = min(kG, T - j0);
                    const int r0 = dir > 0 ? j0 : T - j0 - n;
                    prefetch_l2(le + (size_t)r0 * max_u, (uint32_t)n * max_u * 4u);
                    prefetch_l2(ls + (size_t)r0 * max_u, (uint32_t)n * max_u * 4u);
                }
            }
            if (nprep != 0) load_raw(kp);
            if (p.stats) {  // wait for the loads so that their latency is counted separately
                if (RS[kG - 1][CPL - 1] == 123456.0f && RE[kG - 1][0] == 654321.0f) st_grad_start = 1;
                st_load += clock64() - tp0;
                if (nprep == 0) st_tl[0] = clock64() - st_t0, st_tl[4] = tp0 - st_t0;
            }
            float E[kG][CPL], S[kG][CPL];
#pragma unroll
            for (int q = 0; q < kG; ++q) {
                const int j = kp * kG + q;
                const int t = dir > 0 ? j : T - 1 - j;
                const bool not_last = t != T - 1;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    E[q][i] = me[i] ? ex2(RE[q][i] * kLog2e) : 0.0f;
                    S[q][i] = (ms[i] && not_last) ? ex2(RS[q][i] * kLog2e) : 0.0f;
                }
            }
            if (p.stats && nprep == 0) {
                if (E[kG - 1][CPL - 1] == 123456.0f && S[kG - 1][0] == 654321.0f) st_grad_start = 1;
                st_tl[1] = clock64() - st_t0;
            }
            const long long tp1 = p.stats ? clock64() : 0;
            // The stage prepared one iteration ago is published only now: its stores have long
            // drained, so the GPU-scope fence finds nothing to wait for.
            if (pending) {
                if (lane == 0) remote_publish(chain_prep_done, nprep);
                __syncwarp();
                pending = false;
            }
            const long long tp2 = p.stats ? clock64() : 0;
#pragma unroll
            for (int q = 0; q < kG; ++q) {
                store_cells<CPL>(dst + q * max_u, c0, max_u, E[q]);
                store_cells<CPL>(dst + (kG + q) * max_u, c0, max_u, S[q]);
            }
            __syncwarp();
            if (p.stats && nprep == 0) st_tl[2] = clock64() - st_t0;
            ++nprep;
            if (nprep <= 1) {  // the first stage of every warp goes out at once: the recursion is waiting for it
                if (lane == 0) remote_publish(chain_prep_done, nprep);
                __syncwarp();
                if (p.stats) st_first_pub = clock64() - st_t0;
            } else {
                pending = true;
            }
            kp += kSplitWarps;
            if (p.stats) { const long long t = clock64(); st_prep += (tp1 - tp0) + (t - tp2); st_fence += tp2 - tp1; st_prep_end = t - st_t0; }
            continue;
        }
        if (pending) {  // nothing to prepare right now: do not sit on a finished stage
            const long long tp1 = p.stats ? clock64() : 0;
            if (lane == 0) remote_publish(chain_prep_done, nprep);
            __syncwarp();
            pending = false;
            if (p.stats) st_fence += clock64() - tp1;
        }
        // ---- grad(kq): rows of stage kq that lie in this helper's half ----
        bool can_grad = kq < nst;
        if (can_grad) {
            const int j_lo = max(kq * kG, n1);
            const int j_hi = min((kq + 1) * kG, T);
            // own direction: rows [j_lo, j_hi); other direction: rows T-1-t, i.e. [T - j_hi, T - j_lo)
            can_grad = range_ready(d, j_lo, j_hi) && range_ready(1 - d, T - j_hi, T - j_lo);
            // every warp but the one that produces the likelihood (helper 0, stage k1) needs it first
            if (can_grad && !have_ll && !(d == 0 && kq == k1)) can_grad = flag_load(ll_flag) >= 1;
            if (can_grad) {
                const long long tg0 = p.stats ? clock64() : 0;
                if (p.stats && st_grad_start == 0) st_grad_start = tg0 - st_t0;
                // The rows behind the flags were fenced at GPU scope by their writers before the flags were
                // set; they are read below with ld.global.cg (L2), control-dependent on the flag values.
                if (!have_ll && !(d == 0 && kq == k1)) {
                    asm volatile("fence.acq_rel.cluster;" ::: "memory");
                    f_M = __float_as_int(*reinterpret_cast<volatile float*>(llinfo + 0));
                    f_inv_sum = *reinterpret_cast<volatile float*>(llinfo + 1);
                    f_dead = *reinterpret_cast<volatile float*>(llinfo + 2) != 0.0f;
                    have_ll = true;
                }
                const float* Pst = Pd + (size_t)(kq % ring) * stageP;
                for (int jb = j_lo; jb < j_hi; jb += GB) {
                    float E[GB][CPL], S[GB][CPL], VA[GB][CPL], VB[GB][CPL];
                    int exA[GB], exB[GB];
#pragma unroll
                    for (int r = 0; r < GB; ++r) {
                        const int j = min(jb + r, j_hi - 1);  // rows past the end repeat the last one (not stored)
                        const int q = j - kq * kG;
                        const int t = dir > 0 ? j : T - 1 - j;
                        const float* arow = A0 + (size_t)t * SU;            // alpha(t)
                        const float* brow = A1 + (size_t)(T - 1 - t) * SU;  // beta(t+1)
                        ldcg_cells<CPL>(Pst + q * max_u, c0, E[r]);
                        ldcg_cells<CPL>(Pst + (kG + q) * max_u, c0, S[r]);
                        ldcg_cells<CPL>(arow, c0, VA[r]);
                        ldcg_cells<CPL>(brow, c0, VB[r]);
                        exA[r] = __ldcg(reinterpret_cast<const int*>(arow) + max_u + lane);
                        exB[r] = __ldcg(reinterpret_cast<const int*>(brow) + max_u + lane);
                    }
                    float edge[GB];
                    int exBn[GB];
#pragma unroll
                    for (int r = 0; r < GB; ++r) {
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
                                // share with this CTA's other warps and with helper 1
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
                kq += kSplitWarps;
                if (p.stats) st_grad += clock64() - tg0;
                continue;
            }
        }
        __nanosleep(100);
    }
    if (pending) {
        if (lane == 0) remote_publish(chain_prep_done, nprep);
        __syncwarp();
    }
    if (p.stats && lane == 0) {
        long long* o = p.stats + ((size_t)blockIdx.x * 16 + warp) * 16;
        o[0] = clock64() - st_t0; o[1] = st_prep; o[2] = st_fence; o[3] = st_grad; o[4] = st_prep_end; o[5] = st_grad_start; o[6] = st_load; o[7] = st_first_pub; o[8] = st_tl[0]; o[9] = st_tl[1]; o[10] = st_tl[2]; o[11] = st_tl[4];
    }
}

}  // namespace lattice
}  // namespace ssnt
