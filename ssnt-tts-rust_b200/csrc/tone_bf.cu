// Tone-latent marginalised lattice, block-floating-point split-role kernel (SURVEY.md §8 a-TL).
//
// Same organisation as fb_split.cuh (one cluster of four CTAs per utterance: two recursion CTAs with
// prep and copy-out warps, two gradient CTAs; flag words in the receiver's shared memory; rows through
// L2), with a K-vector of tone classes per token:
//
//   alpha'(u,k) = alpha(u,k) e(u,k) + tone(u,k) X(u-1),      X(u) = sum_k alpha(u,k) s(u,k)
//   beta (u,k)  = e(u,k) beta'(u,k) + s(u,k) Y'(u+1),        Y(u) = sum_k tone(u,k) beta(u,k)
//
// in the probability domain with one shared power-of-two exponent per lane (its CPL tokens x K tones).
// A row is max_u*K floats (2 KB at U=128, K=4), so a step carries ~60 instructions of independent
// work and the one cross-lane value per row (X of the lane's last token / Y of its first) is shuffled
// at the top of the step and consumed at its end: no skew is needed here.
//
// Utterances the block-float arithmetic cannot hold are flagged and re-run by the log-domain kernel
// (tone_fb_kernels.cu) in a second launch that skips everything else.
#include "fb_split.cuh"

namespace ssnt {
using namespace lattice;

namespace {

constexpr int kTThreads = 256;
constexpr int kTWarps = 8;
constexpr int kSR = 4;            // rows per stage
constexpr int kTHeader = 2048;
constexpr int kTPrepWarps = 4;    // warps 1,2,3,5 (sub-partitions 1-3)
constexpr int kTCopyWarps = 3;    // warps 4,6,7

struct ToneBfParams {
    ToneFbArgs a;
    float* A;          // [B][2][nrows][RW + 32]
    float* GT;         // [B][2][kTWarps][max_u*K] partial tone gradients
    unsigned* status;  // [B]
    int NS, nrows;
    unsigned* counter;
    long long* stats;  // profiling aid: [4B CTAs][8 warps][8] cycle counters, or null
};

template <int W>
__device__ __forceinline__ void ld_row(const float* p, float (&v)[W]) {  // W consecutive floats, 16-byte aligned
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(p + 4 * q);
        v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
    }
}
template <int W>
__device__ __forceinline__ void ldcg_row(const float* p, float (&v)[W]) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
        const float4 w = __ldcg(reinterpret_cast<const float4*>(p + 4 * q));
        v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
    }
}
template <int W>
__device__ __forceinline__ void st_row(float* p, const float (&v)[W]) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q)
        *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
template <int W>
__device__ __forceinline__ void stcs_row(float* p, const float (&v)[W]) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q)
        __stcs(reinterpret_cast<float4*>(p + 4 * q), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
}

// Lane-interleaved row layout for everything this kernel owns (shared-memory ring, scratch rows A): the
// q-th float4 of lane l sits at float offset (q*32 + l)*4, so one 128-bit access of a warp covers 512
// contiguous bytes.  In the natural layout (a lane's W floats contiguous, lanes 4*W bytes apart) every
// 128-bit shared-memory access is a 4-way bank conflict at W = 16 (measured: the recursion CTA was
// shared-memory-bandwidth bound at ~1100 cycles per row).
template <int W>
__device__ __forceinline__ void ldp_row(const float* row, int lane, float (&v)[W]) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(row + (q * 32 + lane) * 4);
        v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
    }
}
template <int W>
__device__ __forceinline__ void ldcgp_row(const float* row, int lane, float (&v)[W]) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
        const float4 w = __ldcg(reinterpret_cast<const float4*>(row + (q * 32 + lane) * 4));
        v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
    }
}
template <int W>
__device__ __forceinline__ void stp_row(float* row, int lane, const float (&v)[W]) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q)
        *reinterpret_cast<float4*>(row + (q * 32 + lane) * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// ---------------------------------------------------------------------------------------------------
template <int CPL, int K>
__device__ void tone_chain_cta(const ToneBfParams& p, int b, int d, int T, int U, unsigned char* smem_raw) {
    constexpr int W = CPL * K;              // floats of a row owned by one lane
    constexpr int RW = 32 * W;              // floats of a row
    constexpr int RS = RW + 32;             // scratch row stride (lane exponents behind the row)
    // e/s rows in the ring: token u (recursion lane u/CPL, its (u%CPL)-th float4) at float4 (u%CPL)*S4 + u/CPL.
    // S4 = 32 + 8/CPL makes BOTH the recursion's reads (a lane's q-th float4, lanes consecutive) and the prep
    // warps' token-strided writes (lanes walk u) conflict-free; the state rows keep stride 32.
    constexpr int S4 = CPL == 1 ? 32 : 32 + 8 / CPL;
    constexpr int RWP = CPL * S4 * 4;       // floats of a padded e/s row
    constexpr int slot_floats = 2 * kSR * RWP;
    constexpr int off_s = kSR * RWP;
    const ToneFbArgs& a = p.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NS = p.NS, dir = d == 0 ? 1 : -1;
    const int nst = (T + kSR - 1) / kSR;
    const size_t slab = (size_t)a.max_t * RW;
    const float* le = a.log_emit + (size_t)b * slab;
    const float* ls = a.log_shift + (size_t)b * slab;
    const float* lt = a.log_tone + (size_t)b * RW;
    float* Ad = p.A + ((size_t)b * 2 + d) * (size_t)p.nrows * RS;
    int* ready = reinterpret_cast<int*>(smem_raw + 128);
    int* state_done = reinterpret_cast<int*>(smem_raw + 256);
    int* slot_free = reinterpret_cast<int*>(smem_raw + 384);
    float* ringm = reinterpret_cast<float*>(smem_raw + kTHeader);
    auto slot_ptr = [&](int slot) { return ringm + (size_t)slot * slot_floats; };
    const int f0 = lane * W;                // first float of this lane inside a row
    auto ldpp = [&](const float* row, float (&x)[W]) {  // this lane's CPL float4s of a padded e/s row
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(row + (q * S4 + lane) * 4);
            x[4 * q] = w.x; x[4 * q + 1] = w.y; x[4 * q + 2] = w.z; x[4 * q + 3] = w.w;
        }
    };
    const bool is_copy = warp == 4 || warp >= 6;
    const int copy_idx = warp == 4 ? 0 : warp - 5;        // warps 4,6,7 → 0,1,2
    const int prep_idx = warp - 1 - (warp > 4 ? 1 : 0);  // warps 1,2,3,5 → 0..3

    if (warp == 0) {
        // ------------------------------- recursion -------------------------------
        float tone[W], v[W];
        {
            float raw[W];
            ld_row<W>(lt + f0, raw);
#pragma unroll
            for (int i = 0; i < W; ++i) tone[i] = (lane * CPL + i / K < U) ? ex2(raw[i] * kLog2e) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < W; ++i) v[i] = 0.0f;
        if (d == 0) {
            if (lane == 0)
#pragma unroll
                for (int k = 0; k < K; ++k) v[k] = tone[k];  // alpha(0,0,k) = tone(0,k)
        } else {
#pragma unroll
            for (int i = 0; i < W; ++i)
                if (lane * CPL + i / K == U - 1) v[i] = 1.0f;  // virtual terminal row beta(T, U-1, k) = 1
        }
        int ex = 0, nb_ex = 0;
        const bool edge_lane = d == 0 ? lane == 0 : lane == 31;
        float g = edge_lane ? 0.0f : 1.0f;
        bool have_dec = false;
        int ex_dec = 0, nbex_dec = 0;
        const int lgNS = NS == 8 ? 3 : (NS == 4 ? 2 : 4);
        int4 fl = make_int4(0, 0, 0, 0);
        const long long st0 = p.stats ? clock64() : 0;
        long long st_w = 0, st_f = 0;
        for (int k = 0; k < nst;) {
            const int rem = nst - k;
            const int ns = rem >= 2 ? 2 : 1;  // stages of this round: 8 rows (the ring holds 8 stages; short rounds
                                              // recycle slots early enough for the prep warps to stay ahead)
            const int slot0 = k & (NS - 1);
            const int use = (k >> lgNS) + 1;
            const int* rflag = ready + (slot0 & ~3);
            const int fo = slot0 & 3;
            auto round_ready = [&](const int4& f) {
                const int q[4] = {f.x, f.y, f.z, f.w};
                bool ok = true;
#pragma unroll
                for (int z = 0; z < 4; ++z)
                    if (z >= fo && z < fo + ns) ok = ok && q[z] >= use;
                return ok;
            };
            const long long tw0 = p.stats ? clock64() : 0;
            while (!__all_sync(kFull, round_ready(fl)))
                asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(fl.x), "=r"(fl.y), "=r"(fl.z), "=r"(fl.w) : "r"(smem_u32(rflag)) : "memory");
            const long long tw1 = p.stats ? clock64() : 0;
            st_w += tw1 - tw0;
            {
                const int* nflag = ready + (((k + ns) & (NS - 1)) & ~3);
                asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(fl.x), "=r"(fl.y), "=r"(fl.z), "=r"(fl.w) : "r"(smem_u32(nflag)) : "memory");
            }
            // ---- apply the re-normalisation decided in the previous round ----
            if (have_dec) {
                const int shift = ex - ex_dec;
#pragma unroll
                for (int i = 0; i < W; ++i) v[i] = scale_pow2(v[i], shift);
                ex = ex_dec;
                nb_ex = nbex_dec;
                g = edge_lane ? 0.0f : pow2i(max(-126, min(126, nb_ex - ex)));
                have_dec = false;
            }
            int own = kNoMass, nbmag = kNoMass;
            // rows of the round, the next row's probabilities requested before the current row is computed
            // (the compiler cannot hoist those loads itself: they might alias the state rows stored in between)
            const int nrow = ns * kSR;
            float E[2][W], S[2][W];
            {
                const float* sp0 = slot_ptr(slot0);
                ldpp(sp0, E[0]);
                ldpp(sp0 + off_s, S[0]);
            }
#pragma unroll 2
            for (int r = 0; r < nrow; ++r) {
                float* sp = slot_ptr((slot0 + (r >> 2)) & (NS - 1));
                const int q = r & (kSR - 1);
                const int cb = r & 1;
                if (r + 1 < nrow) {
                    const float* spn = slot_ptr((slot0 + ((r + 1) >> 2)) & (NS - 1));
                    const int qn = (r + 1) & (kSR - 1);
                    if (cb == 0) { ldpp(spn + qn * RWP, E[1]); ldpp(spn + off_s + qn * RWP, S[1]); }
                    else { ldpp(spn + qn * RWP, E[0]); ldpp(spn + off_s + qn * RWP, S[0]); }
                }
                {
                    // the state BEFORE the step is this row; it goes straight to the global scratch (interleaved
                    // layout), with the lane exponents in the first row of each stage
                    float* dst = Ad + (size_t)(k * kSR + r) * RS;
                    stp_row<W>(dst, lane, v);
                    if (q == 0) reinterpret_cast<int*>(dst)[RW + lane] = ex;
                }
                auto step = [&](const float (&Ec)[W], const float (&Sc)[W]) {
                    if (d == 0) {
                        float X[CPL];
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            float x = 0.0f;
#pragma unroll
                            for (int kk = 0; kk < K; ++kk) x = fmaf(v[i * K + kk], Sc[i * K + kk], x);
                            X[i] = x;
                        }
                        const float in = __shfl_up_sync(kFull, X[CPL - 1], 1) * g;
#pragma unroll
                        for (int i = CPL - 1; i >= 1; --i)
#pragma unroll
                            for (int kk = 0; kk < K; ++kk)
                                v[i * K + kk] = fmaf(tone[i * K + kk], X[i - 1], v[i * K + kk] * Ec[i * K + kk]);
#pragma unroll
                        for (int kk = 0; kk < K; ++kk) v[kk] = fmaf(tone[kk], in, v[kk] * Ec[kk]);
                    } else {
                        float Y[CPL];
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            float y = 0.0f;
#pragma unroll
                            for (int kk = 0; kk < K; ++kk) y = fmaf(tone[i * K + kk], v[i * K + kk], y);
                            Y[i] = y;
                        }
                        const float in = __shfl_down_sync(kFull, Y[0], 1) * g;
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            const float yn = (i + 1 < CPL) ? Y[i + 1] : in;
#pragma unroll
                            for (int kk = 0; kk < K; ++kk)
                                v[i * K + kk] = fmaf(Ec[i * K + kk], v[i * K + kk], Sc[i * K + kk] * yn);
                        }
                    }
                };
                if (cb == 0) step(E[0], S[0]);
                else step(E[1], S[1]);
            }
            {
                // decide the next frame from the state at the round's end (the shuffles overlap the hand-off)
                float mx = v[0];
#pragma unroll
                for (int i = 1; i < W; ++i) mx = fmaxf(mx, v[i]);
                own = mx > 0.0f ? ex + ilogb_pos(mx) - kTarget : kNoMass;
                float edge = 0.0f;  // magnitude of what the neighbour lane will receive from this one
                if (d == 0) {
#pragma unroll
                    for (int kk = 0; kk < K; ++kk) edge = fmaxf(edge, v[(CPL - 1) * K + kk]);
                } else {
#pragma unroll
                    for (int kk = 0; kk < K; ++kk) edge = fmaxf(edge, v[kk]);
                }
                const int amag = edge > 0.0f ? ex + ilogb_pos(edge) : kNoMass;
                nbmag = d == 0 ? __shfl_up_sync(kFull, amag, 1) : __shfl_down_sync(kFull, amag, 1);
                if (edge_lane) nbmag = kNoMass;
                int nw = max(own, nbmag - kTarget - kSlack);
                if (nw <= kNoMass / 2) nw = ex;
                ex_dec = nw;
                nbex_dec = d == 0 ? __shfl_up_sync(kFull, nw, 1) : __shfl_down_sync(kFull, nw, 1);
                have_dec = true;
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                for (int z = 0; z < ns; ++z)
                    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(state_done + ((slot0 + z) & (NS - 1)))), "r"(use) : "memory");
            }
            __syncwarp();
            k += ns;
        }
        if (p.stats && lane == 0) {
            long long* o = p.stats + ((size_t)blockIdx.x * 8 + warp) * 8;
            o[0] = clock64() - st0; o[1] = st_w; o[2] = st_f;
        }
    } else if (!is_copy) {
        // ------------------------------- prep -------------------------------
        // token-strided ownership (lane l: tokens q*32 + l, one float4 of K tones each): coalesced global loads
        const float4* le4 = reinterpret_cast<const float4*>(le);
        const float4* ls4 = reinterpret_cast<const float4*>(ls);
        constexpr int R4 = RW / 4;
        bool me[CPL], ms[CPL];
        int ppos[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int u = q * 32 + lane;
            me[q] = u < U;
            ms[q] = u < U - 1;
            ppos[q] = ((u % CPL) * S4 + u / CPL) * 4;
        }
        const long long st0 = p.stats ? clock64() : 0;
        long long st_w = 0;
        for (int k = prep_idx; k < nst; k += kTPrepWarps) {
            const int slot = k & (NS - 1);
            {
                const int kf = k + 2 * kTPrepWarps;  // L2 prefetch two iterations ahead, one 128-byte line per lane
                if (kf < nst) {
                    const int j0 = kf * kSR, n = min(kSR, T - j0);
                    const int r0 = dir > 0 ? j0 : T - j0 - n;
                    const int nlines = n * RW / 32;
                    for (int l = lane; l < nlines; l += 32) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(le + (size_t)r0 * RW + l * 32));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ls + (size_t)r0 * RW + l * 32));
                    }
                }
            }
            float4 RE[kSR][CPL], RSv[kSR][CPL];
#pragma unroll
            for (int q = 0; q < kSR; ++q) {
                const int j = k * kSR + q;
                const int t = dir > 0 ? j : T - 1 - j;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    if (j < T) {
                        RE[q][c] = __ldcg(le4 + (size_t)t * R4 + c * 32 + lane);
                        RSv[q][c] = __ldcg(ls4 + (size_t)t * R4 + c * 32 + lane);
                    } else {
                        RE[q][c] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                        RSv[q][c] = RE[q][c];
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kSR; ++q) {
                const int j = k * kSR + q;
                const int t = dir > 0 ? j : T - 1 - j;
                const bool not_last = t != T - 1;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    float4& E = RE[q][c];
                    float4& S = RSv[q][c];
                    if (me[c]) E = make_float4(ex2(E.x * kLog2e), ex2(E.y * kLog2e), ex2(E.z * kLog2e), ex2(E.w * kLog2e));
                    else E = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ms[c] && not_last) S = make_float4(ex2(S.x * kLog2e), ex2(S.y * kLog2e), ex2(S.z * kLog2e), ex2(S.w * kLog2e));
                    else S = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            { const long long t0 = p.stats ? clock64() : 0;
              // the slot's e/s region is free once the recursion has consumed its previous occupant
              if (k >= NS) wait_flag_ge(state_done + slot, k / NS, 128);
              if (p.stats) st_w += clock64() - t0; }
            float* dst = slot_ptr(slot);
#pragma unroll
            for (int q = 0; q < kSR; ++q)
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    *reinterpret_cast<float4*>(dst + q * RWP + ppos[c]) = RE[q][c];
                    *reinterpret_cast<float4*>(dst + off_s + q * RWP + ppos[c]) = RSv[q][c];
                }
            __syncwarp();
            if (lane == 0) flag_publish(ready + slot, k / NS + 1);
        }
        if (p.stats && lane == 0) {
            long long* o = p.stats + ((size_t)blockIdx.x * 8 + warp) * 8;
            o[0] = clock64() - st0; o[1] = st_w;
        }
    } else {
        // ------------------------------- copy-out -------------------------------
        const uint32_t h0 = map_to_rank(smem_raw + 1024, 2), h1 = map_to_rank(smem_raw + 1024, 3);
        const int nround = (nst + 3) / 4;  // publish per four stages (16 rows)
        for (int r = copy_idx; r < nround; r += kTCopyWarps) {
            const int kend = min(4 * r + 4, nst);
            for (int k = 4 * r; k < kend; ++k) wait_flag_ge(state_done + (k & (NS - 1)), k / NS + 1, 128);
            // the recursion warp stored and CTA-released the rows itself; the GPU-scope fence below, after
            // observing its flags, orders them before the flags written into the gradient CTAs
            if (lane == 0) {
                __threadfence();
                const uint32_t off = 4u * (uint32_t)(d * kRoundRing + (r % kRoundRing));
                asm volatile("st.relaxed.cluster.shared::cluster.s32 [%0], %1;" ::"r"(h0 + off), "r"(r + 1) : "memory");
                asm volatile("st.relaxed.cluster.shared::cluster.s32 [%0], %1;" ::"r"(h1 + off), "r"(r + 1) : "memory");
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Gradient CTA.  Lane ownership is TOKEN-STRIDED here: lane l owns the tokens u = q*32 + l (q < CPL), each a
// float4 of K = 4 tones, so every access to the [.., U, K] tensors is one fully coalesced 512-byte request per
// instruction (the recursion's "CPL consecutive tokens per lane" would touch 16 lines per instruction).  The
// scratch rows are in the recursion's interleaved layout: token u sits at float4 (u % CPL)*32 + u / CPL and
// carries the exponent of recursion lane u / CPL.
template <int CPL, int K>
__device__ void tone_grad_cta(const ToneBfParams& p, int b, int d, int T, int U, unsigned char* smem_raw) {
    static_assert(K == 4, "one float4 per token");
    constexpr int RW = 32 * CPL * K, RS = RW + 32;
    constexpr int ROWS = 16;  // rows per published round
    const ToneFbArgs& a = p.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dir = d == 0 ? 1 : -1;
    const int m = (T + 1) >> 1;
    const int n1 = d == 0 ? m - 1 : T - m + 1;
    const size_t slab = (size_t)a.max_t * RW;
    const float4* le = reinterpret_cast<const float4*>(a.log_emit + (size_t)b * slab);
    const float4* ls = reinterpret_cast<const float4*>(a.log_shift + (size_t)b * slab);
    const float4* lt = reinterpret_cast<const float4*>(a.log_tone + (size_t)b * RW);
    float4* ge = reinterpret_cast<float4*>(a.grad_emit + (size_t)b * slab);
    float4* gs = reinterpret_cast<float4*>(a.grad_shift + (size_t)b * slab);
    const float* A0 = p.A + ((size_t)b * 2 + 0) * (size_t)p.nrows * RS;  // row j = alpha(j)
    const float* A1 = p.A + ((size_t)b * 2 + 1) * (size_t)p.nrows * RS;  // row j = beta(T - j)
    float4* GT = reinterpret_cast<float4*>(p.GT + (((size_t)b * 2 + d) * kTWarps + warp) * RW);
    const int* round_done = reinterpret_cast<const int*>(smem_raw + 1024);
    int* ll_flag = reinterpret_cast<int*>(smem_raw + 704);
    float* llinfo = reinterpret_cast<float*>(smem_raw + 720);
    constexpr int R4 = RW / 4;  // float4s per row (= tokens)

    auto rows_ready = [&](int dd, int n) {
        if (n <= 0) return true;
        const int r = (n - 1) / ROWS;
        return flag_load(round_done + dd * kRoundRing + (r % kRoundRing)) >= r + 1;
    };
    auto ready = [&](int jj) { return rows_ready(d, jj + 1) && rows_ready(1 - d, T - jj); };

    int apos[CPL], alane[CPL];  // where this lane's tokens sit in a scratch row / whose exponent they carry
    bool me[CPL], ms[CPL];
    float4 tone[CPL], gacc[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const int u = q * 32 + lane;
        apos[q] = (u % CPL) * 32 + u / CPL;
        alane[q] = u / CPL;
        me[q] = u < U;
        ms[q] = u < U - 1;
        const float4 r = lt[u];
        tone[q] = me[q] ? make_float4(ex2(r.x * kLog2e), ex2(r.y * kLog2e), ex2(r.z * kLog2e), ex2(r.w * kLog2e))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        gacc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float f_inv_sum = 0.0f;
    int f_M = 0;
    bool f_dead = false, have_ll = false;
    const long long st0 = p.stats ? clock64() : 0;
    long long st_w = 0, st_start = 0;

    struct Row {
        float4 E[CPL], S[CPL], VA[CPL], VB[CPL];
        int exA[CPL], exB[CPL];
    };
    auto load = [&](int jj, Row& r) {
        const int t = dir > 0 ? jj : T - 1 - jj;
        const float* arow = A0 + (size_t)t * RS;            // alpha(t)
        const float* brow = A1 + (size_t)(T - 1 - t) * RS;  // beta(t+1)
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            r.E[q] = __ldcg(le + (size_t)t * R4 + q * 32 + lane);
            r.S[q] = __ldcg(ls + (size_t)t * R4 + q * 32 + lane);
            r.VA[q] = __ldcg(reinterpret_cast<const float4*>(arow) + apos[q]);
            r.VB[q] = __ldcg(reinterpret_cast<const float4*>(brow) + apos[q]);
            // lane exponents: in the first row of the 4-row stage of the sweep that wrote the row
            r.exA[q] = __ldcg(reinterpret_cast<const int*>(A0 + (size_t)(t & ~(kSR - 1)) * RS) + RW + alane[q]);
            r.exB[q] = __ldcg(reinterpret_cast<const int*>(A1 + (size_t)((T - 1 - t) & ~(kSR - 1)) * RS) + RW + alane[q]);
        }
    };
    auto prefetch_row = [&](int jj) {
        if (jj >= T) return;
        const int t = dir > 0 ? jj : T - 1 - jj;
        const float* src[4] = {reinterpret_cast<const float*>(le + (size_t)t * R4), reinterpret_cast<const float*>(ls + (size_t)t * R4),
                               A0 + (size_t)t * RS, A1 + (size_t)(T - 1 - t) * RS};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int nl = (r < 2 ? RW : RS) / 32 + (r < 2 ? 0 : 1);
            if (lane < nl) asm volatile("prefetch.global.L2 [%0];" ::"l"(src[r] + lane * 32));
        }
    };
    auto dot4 = [](const float4& x, const float4& y) { return fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, x.w * y.w))); };
    auto compute = [&](int j, Row& r) {
        const int t = dir > 0 ? j : T - 1 - j;
        const bool ll_producer = d == 0 && j == n1;
        if (!have_ll && !ll_producer) {
            wait_flag_ge(ll_flag, 1, 200);
            asm volatile("fence.acq_rel.cluster;" ::: "memory");
            f_M = __float_as_int(*reinterpret_cast<volatile float*>(llinfo + 0));
            f_inv_sum = *reinterpret_cast<volatile float*>(llinfo + 1);
            f_dead = *reinterpret_cast<volatile float*>(llinfo + 2) != 0.0f;
            have_ll = true;
        }
        const bool not_last = t != T - 1;
        float Y[CPL], X[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            float4& E = r.E[q];
            float4& S = r.S[q];
            if (me[q]) E = make_float4(ex2(E.x * kLog2e), ex2(E.y * kLog2e), ex2(E.z * kLog2e), ex2(E.w * kLog2e));
            else E = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ms[q] && not_last) S = make_float4(ex2(S.x * kLog2e), ex2(S.y * kLog2e), ex2(S.z * kLog2e), ex2(S.w * kLog2e));
            else S = make_float4(0.f, 0.f, 0.f, 0.f);
            Y[q] = dot4(tone[q], r.VB[q]);  // Y(t+1, u) = sum_k tone(u,k) beta(t+1,u,k)
            X[q] = dot4(r.VA[q], S);        // X(t, u)   = sum_k alpha(t,u,k) s(t,u,k)
        }
        // neighbours: token u+1 is lane+1 (same q) or lane 0 of q+1; token u-1 is lane-1 or lane 31 of q-1;
        // both re-framed into this token's exponents
        float yn[CPL], xp[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            float y = __shfl_down_sync(kFull, Y[q], 1);
            int ey = __shfl_down_sync(kFull, r.exB[q], 1);
            const float y0 = q + 1 < CPL ? __shfl_sync(kFull, Y[q + 1 < CPL ? q + 1 : q], 0) : 0.0f;
            const int ey0 = q + 1 < CPL ? __shfl_sync(kFull, r.exB[q + 1 < CPL ? q + 1 : q], 0) : 0;
            if (lane == 31) { y = y0; ey = ey0; }
            yn[q] = (q + 1 < CPL || lane < 31) ? scale_pow2(y, ey - r.exB[q]) : 0.0f;
            float x = __shfl_up_sync(kFull, X[q], 1);
            int exx = __shfl_up_sync(kFull, r.exA[q], 1);
            const float x31 = q > 0 ? __shfl_sync(kFull, X[q > 0 ? q - 1 : q], 31) : 0.0f;
            const int ex31 = q > 0 ? __shfl_sync(kFull, r.exA[q > 0 ? q - 1 : q], 31) : 0;
            if (lane == 0) { x = x31; exx = ex31; }
            xp[q] = (q > 0 || lane > 0) ? scale_pow2(x, exx - r.exA[q]) : 0.0f;
        }
        if (ll_producer) {
            // Z = sum_{u,k} alpha(m-1,u,k) (e beta(m,u,k) + s Y(m,u+1)), tokens carry different exponents
            float w[CPL];
            int key = kNoMass;
            bool finite = true;
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const float4 &E = r.E[q], &S = r.S[q], &VA = r.VA[q], &VB = r.VB[q];
                w[q] = VA.x * (E.x * VB.x + S.x * yn[q]) + VA.y * (E.y * VB.y + S.y * yn[q]) +
                       VA.z * (E.z * VB.z + S.z * yn[q]) + VA.w * (E.w * VB.w + S.w * yn[q]);
                finite = finite && (w[q] == w[q] && w[q] < 3.0e38f);
                if (w[q] > 0.0f && w[q] < 3.0e38f) key = max(key, r.exA[q] + r.exB[q] + ilogb_pos(w[q]));
            }
            int M = key;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) M = max(M, __shfl_xor_sync(kFull, M, o));
            float part = 0.0f;
#pragma unroll
            for (int q = 0; q < CPL; ++q)
                if (w[q] > 0.0f && w[q] < 3.0e38f) part += scale_pow2(w[q], r.exA[q] + r.exB[q] - M);
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
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int u = q * 32 + lane;
            const int kf = max(-252, min(252, r.exA[q] + r.exB[q] - f_M));
            const int kh = kf >> 1;
            const float fa = pow2i(max(-126, kh));
            const float fb = pow2i(max(-126, kf - kh)) * f_inv_sum;
            const float4 &E = r.E[q], &S = r.S[q], &VA = r.VA[q], &VB = r.VB[q];
            const float4 va = make_float4(VA.x * fa, VA.y * fa, VA.z * fa, VA.w * fa);
            float4 g1 = make_float4(va.x * ((E.x * VB.x) * fb), va.y * ((E.y * VB.y) * fb), va.z * ((E.z * VB.z) * fb), va.w * ((E.w * VB.w) * fb));
            const float sy = yn[q] * fb;
            float4 g2 = make_float4(va.x * (S.x * sy), va.y * (S.y * sy), va.z * (S.z * sy), va.w * (S.w * sy));
            if (f_dead) { g1 = make_float4(0.f, 0.f, 0.f, 0.f); g2 = g1; }
            __stcs(ge + (size_t)t * R4 + u, g1);
            __stcs(gs + (size_t)t * R4 + u, g2);
            if (!f_dead) {
                // d LL / d log_tone(u,k): entering token u with tone k at frame t+1 (X(t,u-1) tone beta / Z) ...
                const float xf = xp[q] * fa;
                gacc[q].x += xf * ((tone[q].x * VB.x) * fb);
                gacc[q].y += xf * ((tone[q].y * VB.y) * fb);
                gacc[q].z += xf * ((tone[q].z * VB.z) * fb);
                gacc[q].w += xf * ((tone[q].w * VB.w) * fb);
                if (t == 0 && u == 0) {  // ... and token 0 draws its tone at the start: occupancy of (0, 0, k)
                    gacc[q].x += g1.x + g2.x; gacc[q].y += g1.y + g2.y; gacc[q].z += g1.z + g2.z; gacc[q].w += g1.w + g2.w;
                }
                bool bad = false;
                if (d == 0 && t == T - 1 && u == U - 1) bad = !(fabsf(g1.x + g1.y + g1.z + g1.w - 1.0f) < kBfConsistency);
                if (d == 1 && t == 0 && u == 0)
                    bad = !(fabsf(g1.x + g1.y + g1.z + g1.w + g2.x + g2.y + g2.z + g2.w - 1.0f) < kBfConsistency);
                if (bad) atomicOr(p.status + b, (unsigned)kBfInconsistent);
            }
        }
    };
    auto wait_ready = [&](int jj) {
        const long long tw0 = p.stats ? clock64() : 0;
        while (!ready(jj)) __nanosleep(200);
        if (p.stats) { const long long t1 = clock64(); st_w += t1 - tw0; if (!st_start) st_start = t1 - st0; }
    };
    {
        // one sweep row per work unit; warp w takes the rows n1 + w, n1 + w + 8, ...; the next row's loads are
        // issued before the current row is computed (two register sets) and L2 prefetches run further ahead
        Row ra, rb;
        bool hasA = false, hasB = false;
        int j = n1 + warp;
        constexpr int kAhead = 4;
        for (int q = 1; q <= kAhead; ++q) prefetch_row(j + q * kTWarps);
        while (j < T) {
            prefetch_row(j + (kAhead + 1) * kTWarps);
            prefetch_row(j + (kAhead + 2) * kTWarps);
            if (!hasA) { wait_ready(j); load(j, ra); }
            hasA = false;
            const int j2 = j + kTWarps;
            if (j2 < T && ready(j2)) { load(j2, rb); hasB = true; }
            compute(j, ra);
            j = j2;
            if (j >= T) break;
            if (!hasB) { wait_ready(j); load(j, rb); }
            hasB = false;
            const int j3 = j + kTWarps;
            if (j3 < T && ready(j3)) { load(j3, ra); hasA = true; }
            compute(j, rb);
            j = j3;
        }
    }
#pragma unroll
    for (int q = 0; q < CPL; ++q) GT[q * 32 + lane] = gacc[q];
    if (p.stats && lane == 0) {
        long long* o = p.stats + ((size_t)blockIdx.x * 8 + warp) * 8;
        o[0] = clock64() - st0; o[1] = st_w; o[2] = st_start;
    }
}

template <int CPL, int K>
__global__ void __launch_bounds__(kTThreads, 1) tone_split_kernel(const ToneBfParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int W = CPL * K, RW = 32 * W;
    const int tid = threadIdx.x;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int b = blockIdx.x >> 2;
    const ToneFbArgs& a = p.a;
    int T = a.t_len ? a.t_len[b] : a.max_t;
    int U = a.u_len ? a.u_len[b] : a.max_u;
    T = min(max(T, 0), a.max_t);
    U = min(max(U, 0), a.max_u);
    const size_t slab = (size_t)a.max_t * RW;
    if (T <= 0 || U <= 0 || U > T) {
        if (rank >= 2) {
            float4* g = reinterpret_cast<float4*>((rank == 2 ? a.grad_emit : a.grad_shift) + (size_t)b * slab);
            for (size_t i = tid; i < slab / 4; i += kTThreads) __stcs(g + i, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        if (rank == 0) {
            for (int i = tid; i < RW; i += kTThreads) a.grad_tone[(size_t)b * RW + i] = 0.0f;
            if (tid == 0) {
                a.log_likelihood[b] = -INFINITY;
                p.status[b] = 0u;  // nothing to re-run: the follow-up log-domain kernel must not read an unwritten word
            }
        }
    } else {
        for (int i = tid; i < (kTHeader - 128) / 4; i += kTThreads) reinterpret_cast<int*>(smem_raw + 128)[i] = 0;
        if (rank == 0 && tid == 0) p.status[b] = 0u;
        __syncthreads();
        cluster.sync();
        if (rank < 2) tone_chain_cta<CPL, K>(p, b, (int)rank, T, U, smem_raw);
        else tone_grad_cta<CPL, K>(p, b, (int)rank - 2, T, U, smem_raw);
        if (rank >= 2) {  // padded frames
            float4* g = reinterpret_cast<float4*>((rank == 2 ? a.grad_emit : a.grad_shift) + (size_t)b * slab + (size_t)T * RW);
            const size_t n4 = (size_t)(a.max_t - T) * RW / 4;
            for (size_t i = tid; i < n4; i += kTThreads) __stcs(g + i, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        __threadfence();
        cluster.sync();
        // deterministic reduction of the sixteen per-warp partial tone gradients (rank 0, fixed order)
        if (rank == 0) {
            const float* GT = p.GT + (size_t)b * 2 * kTWarps * RW;
            for (int i = tid; i < RW; i += kTThreads) {
                float acc = 0.0f;
                for (int q = 0; q < 2 * kTWarps; ++q) acc += __ldcg(GT + (size_t)q * RW + i);
                a.grad_tone[(size_t)b * RW + i] = (i / K) < U ? acc : 0.0f;
            }
        }
    }
    if (rank == 0) finish_loss(a.log_likelihood, a.loss, a.batch_size, p.counter, tid, kTThreads);
}

template <int CPL, int K>
void launch_tone_split(const ToneBfParams& p, size_t smem, cudaStream_t stream) {
    static size_t configured_[64] = {};  // per device
    size_t& configured = configured_[device_ordinal()];
    if (configured == 0) configured = 48 * 1024;
    if (smem > configured) {
        SSNT_CUDA(cudaFuncSetAttribute(tone_split_kernel<CPL, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)p.a.batch_size * 4u);
    cfg.blockDim = dim3(kTThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SSNT_CUDA(cudaLaunchKernelEx(&cfg, tone_split_kernel<CPL, K>, p));
}

}  // namespace

// Shapes the block-float tone kernel takes: K = 4, max_u in {32, 64, 128}, 16-byte aligned tensors.
bool tone_bf_supported(const ToneFbArgs& a) {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    return a.tone_class_size == 4 && (a.max_u == 32 || a.max_u == 64 || a.max_u == 128) && al(a.log_emit) &&
           al(a.log_shift) && al(a.log_tone) && al(a.grad_emit) && al(a.grad_shift) && al(a.grad_tone);
}
size_t tone_bf_workspace_bytes(int B, int max_t, int max_u, int K) {
    if (K != 4 || !(max_u == 32 || max_u == 64 || max_u == 128)) return 0;
    const size_t RW = (size_t)max_u * K, RS = RW + 32;
    const size_t nrows = (((size_t)max_t + kSR - 1) / kSR) * kSR;
    return ((size_t)B * 2 * nrows * RS + (size_t)B * 2 * kTWarps * RW) * sizeof(float) +
           (((size_t)B * sizeof(unsigned) + 255) & ~(size_t)255) + 512;
}
// Runs the block-float kernel; returns the device pointer of the [B] status words (non-zero = utterance
// must be re-run in the log domain).
unsigned* launch_tone_bf(const ToneFbArgs& a, void* ws, unsigned* counter, cudaStream_t stream) {
    ToneBfParams p;
    p.a = a;
    const int K = 4;
    const size_t RW = (size_t)a.max_u * K, RS = RW + 32;
    p.nrows = ((a.max_t + kSR - 1) / kSR) * kSR;
    p.A = (float*)ws;
    p.GT = p.A + (size_t)a.batch_size * 2 * p.nrows * RS;
    p.status = (unsigned*)(p.GT + (size_t)a.batch_size * 2 * kTWarps * RW);
    p.counter = counter;
    p.stats = fb_get_stats_buffer();
    const int cpl = a.max_u / 32;
    const size_t RWP = (size_t)cpl * (cpl == 1 ? 32 : 32 + 8 / cpl) * 4;  // padded e/s row, see tone_chain_cta
    const size_t slot_bytes = ((size_t)2 * kSR * RWP) * sizeof(float);
    int NS = (int)((size_t)(224 * 1024 - kTHeader) / slot_bytes);
    NS = NS >= 16 ? 16 : (NS >= 8 ? 8 : 4);
    p.NS = NS;
    const size_t smem = kTHeader + (size_t)NS * slot_bytes;
    if (a.max_u == 32) launch_tone_split<1, 4>(p, smem, stream);
    else if (a.max_u == 64) launch_tone_split<2, 4>(p, smem, stream);
    else launch_tone_split<4, 4>(p, smem, stream);
    return p.status;
}

}  // namespace ssnt
