// Beam-search single-step kernels: v1 Emit/Shift (src/lib.rs:149-230), v2 duration classes with
// diagonal-band pruning (src/v2.rs:94-166, 269-339) and tone-latent (src/tone_latent.rs:79-95,
// 184-234).  All three share the reference's skeleton:
//
//   expand every live beam w into its class candidates in (w asc, class asc) order
//   → stable sort by log-prob descending → drop an element equal (ignoring the parent) to the
//   one before it → [v2: remember the first on-diagonal survivor] → pad cyclically to W →
//   truncate to W → [v2: overwrite the last slot with the diagonal candidate].
//
// One warp owns one batch entry.  The candidate table lives in that warp's shared-memory slice;
// compaction and de-duplication use ballot/popc scans, the stable descending sort is a
// rank-by-counting select in which each lane ranks its candidates against keys broadcast across
// the warp (ties broken by candidate position, which is exactly what makes a sort "stable").
// The results must equal the Rust path bit for bit, so the v2 band arithmetic is spelled with
// __f*_rn intrinsics (no FMA contraction) and __float2int_rz (saturating, NaN→0 like `as i32`).
#include "ssnt_common.cuh"

namespace ssnt {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpsPerBlock = 4;

enum Variant { kV1 = 0, kV2 = 1, kTone = 2 };

struct BeamParams {
    // inputs
    const float* h;              // [B, W, C]
    const float* hist;           // [B, W]
    const bool* fin;             // [B, W]
    const int* total;            // [B, W]   (v2)
    const int* dur_table;        // [C]      (v2)
    const int* t;                // [B, W]
    const int* u;                // [B, W]
    const int* in_len;           // [B]      (v2, tone); v1 uses max_t
    const int* out_len;          // [B]      (v2)
    int B, W, C;
    int max_t;                   // v1
    int special_id;              // v2: zero_duration_id, tone: empty_tone_id
    bool allow_skip, test_mode;  // v2
    // outputs [B, W]
    int* prediction;
    float* log_probs;
    int* next_t;
    int* next_u;
    bool* next_fin;
    int* next_total;  // v2
    int* parent;
    unsigned* err;
};

// Candidate table, structure-of-arrays in shared memory (N = W*C slots per warp).
struct Table {
    float* lp;
    int* pred;
    int* nt;
    int* nu;
    int* tot;
    int* par;
    unsigned char* fin;
    unsigned char* valid;
    int* order;  // compacted → sorted → de-duplicated slot indices (reused in place)
    int* tmp;
    float* key;  // log-probs in compacted order (sort keys)
};

__device__ __forceinline__ bool same_bucket(const Table& tb, int i, int j) {
    // eq_ignore_parent: src/lib.rs:80-88, src/v2.rs:180-189, src/tone_latent.rs:108-116
    return tb.pred[i] == tb.pred[j] && tb.lp[i] == tb.lp[j] && tb.nt[i] == tb.nt[j] &&
           tb.nu[i] == tb.nu[j] && tb.fin[i] == tb.fin[j] && tb.tot[i] == tb.tot[j];
}

// src/v2.rs:94-104
__device__ __forceinline__ void v2_bounds(int in_len, int out_len, int t, int& lo, int& hi) {
    const float ratio = __fdiv_rn((float)out_len, (float)in_len);
    const float diagonal = __fmul_rn(ratio, (float)(t + 1));
    const float upper_range = __fmul_rn((float)out_len, 0.1f);
    const float lower_range = __fmul_rn((float)out_len, 0.05f);
    float lb = __fsub_rn(diagonal, lower_range);
    lb = (lb != lb) ? 0.0f : (lb > 0.0f ? lb : 0.0f);
    float ub = __fadd_rn(diagonal, upper_range);
    const float ol = (float)out_len;
    ub = (ub != ub) ? ol : (ub < ol ? ub : ol);
    lo = __float2int_rz(lb);
    hi = __float2int_rz(ub);
}
// src/v2.rs:113-117
__device__ __forceinline__ bool v2_on_diagonal(int in_len, int out_len, int next_t, int total) {
    const float ratio = __fdiv_rn((float)out_len, (float)in_len);
    const float diagonal = __fmul_rn(ratio, (float)next_t);
    const float diff = __fsub_rn((float)total, diagonal);
    return diff >= -20.0f && diff <= 0.0f;
}

// One batch entry's state and result rows ([W] each; global or shared memory).
struct BeamRow {
    const float* h;      // [W, C]
    const float* hist;   // [W]
    const bool* fin;
    const int* total;    // v2
    const int* t;
    const int* u;
    int* prediction;
    float* log_probs;
    int* next_t;
    int* next_u;
    bool* next_fin;
    int* next_total;     // v2
    int* parent;
};

__device__ __forceinline__ Table carve_table(unsigned char* base, int N) {
    Table tb;
    tb.lp = reinterpret_cast<float*>(base);
    tb.pred = reinterpret_cast<int*>(tb.lp + N);
    tb.nt = tb.pred + N;
    tb.nu = tb.nt + N;
    tb.tot = tb.nu + N;
    tb.par = tb.tot + N;
    tb.order = tb.par + N;
    tb.tmp = tb.order + N;
    tb.key = reinterpret_cast<float*>(tb.tmp + N);
    tb.fin = reinterpret_cast<unsigned char*>(tb.key + N);
    tb.valid = tb.fin + N;
    return tb;
}
__host__ __device__ inline size_t table_bytes(size_t N) { return ((N * (9 * 4 + 2) + 16) + 15) & ~(size_t)15; }

// One beam step of one batch entry, executed by one warp.  Returns false if no candidate survived
// (src/v2.rs:292 assert_ne! / `i % 0` in src/tone_latent.rs:199): the error flag is raised, nothing is written.
template <int V>
__device__ bool beam_step_warp(const BeamParams& p, const Table& tb, const BeamRow& r, long long in_len, long long out_len, int lane) {
    const int W = p.W, C = p.C, N = W * C;
    const float* h = r.h;
    const float* hist = r.hist;
    const bool* fin = r.fin;
    const int* tt = r.t;
    const int* uu = r.u;

    // ---- 1. expand ------------------------------------------------------------------------
    for (int s = lane; s < N; s += 32) {
        const int w = s / C, c = s - w * C;
        const int t = tt[w], u = uu[w];
        const float hp = hist[w];
        const bool defined = t >= 0 && (long long)t < in_len;  // usize compare in the reference
        bool valid = false, f = false;
        int pred = 0, nt = t, nu = u, tot = 0;
        float lp = hp;
        if (!defined || fin[w]) {
            // "End of input. Return values to fill padding region." — one filler per beam
            if (c == 0) {
                valid = true;
                f = true;
                pred = V == kV1 ? 0 : p.special_id;
                tot = V == kV2 ? r.total[w] : 0;
            }
        } else if (V == kV1) {
            const bool last = (long long)t == in_len - 1;
            valid = true;
            if (c == 0) {  // Emit
                lp = __fadd_rn(hp, h[w * 2 + 0]);
                if (last) { f = true; } else { nu = u + 1; }
            } else if (last) {  // Shift is prohibited at the last input position
                pred = 0; lp = hp; f = true;
            } else {  // Shift
                pred = 1; lp = __fadd_rn(hp, h[w * 2 + 1]); nt = t + 1; nu = u + 1;
            }
        } else if (V == kV2) {
            const int duration = p.dur_table[c];
            tot = r.total[w] + duration;
            int lo, hi;
            v2_bounds((int)in_len, (int)out_len, t, lo, hi);
            const unsigned long long remaining = (unsigned long long)(in_len - ((long long)t + 1));
            const bool overrun = remaining * 3ull > (unsigned long long)out_len;  // src/v2.rs:106-111
            const bool last = (long long)t == in_len - 1;
            bool keep = true;
            if (!p.test_mode && (tot < lo || tot > hi)) keep = false;
            else if (!p.test_mode && overrun) keep = false;
            else if (last) {
                if (!p.test_mode && tot != (int)out_len) keep = false;
                else if (!p.allow_skip && c == p.special_id) keep = false;
                else f = true;
            } else if (!p.allow_skip && c == p.special_id) keep = false;
            valid = keep;
            pred = c;
            lp = __fadd_rn(hp, h[s]);
            if (!f) { nt = t + 1; nu = u + 1; }
        } else {  // tone latent: every class survives, never finishes here
            valid = true;
            pred = c;
            lp = __fadd_rn(hp, h[s]);
            nt = t + 1; nu = u + 1;
        }
        tb.lp[s] = lp; tb.pred[s] = pred; tb.nt[s] = nt; tb.nu[s] = nu; tb.tot[s] = tot;
        tb.par[s] = w; tb.fin[s] = f ? 1 : 0; tb.valid[s] = valid ? 1 : 0;
    }
    __syncwarp();

    int kept = 0, diag = -1;
    if (N <= 64) {
        // Few candidates (e.g. the tone-latent step, 8 beams x 4 tones): a full stable sort by rank counting is cheaper
        // than the extraction rounds below.
        // ---- 2. compact valid slots, keeping (w, class) order -----------------------------------
        int n = 0;
        for (int s0 = 0; s0 < N; s0 += 32) {
            const int s = s0 + lane;
            const bool v = s < N && tb.valid[s];
            const unsigned bal = __ballot_sync(kFull, v);
            if (v) {
                const int k = n + __popc(bal & ((1u << lane) - 1u));
                tb.tmp[k] = s;
                // Sort keys form a total order: the reference's comparator (partial_cmp().unwrap_or(Equal), src/lib.rs:161)
                // leaves the position of NaN log-probs to its sort algorithm; here NaN ranks below everything, after -inf
                // entries of the same run, so the ranks below are always a permutation.
                const float k0 = tb.lp[s];
                tb.key[k] = (k0 != k0) ? -INFINITY : k0;
            }
            n += __popc(bal);
        }
        __syncwarp();
        if (n == 0) {  // src/v2.rs:292 assert_ne! / `i % 0` in src/tone_latent.rs:199
            if (lane == 0) atomicOr(p.err, V == kV2 ? kErrV2EmptyBeam : kErrToneEmptyBeam);
            return false;
        }
        // ---- 3. stable descending sort: rank = #{j : lp_j > lp_i or (lp_j == lp_i and j before i)} ---
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const float mine = i < n ? tb.key[i] : 0.0f;
            int rank = 0;
#pragma unroll 4
            for (int j = 0; j < n; ++j) {
                const float other = tb.key[j];  // same address across the warp → broadcast
                rank += (other > mine || (other == mine && j < i)) ? 1 : 0;
            }
            if (i < n) tb.order[rank] = tb.tmp[i];
        }
        __syncwarp();
        // ---- 4. drop consecutive duplicates (first of a run survives) ------------------------------
        for (int r0 = 0; r0 < n; r0 += 32) {
            const int q = r0 + lane;
            bool keep = false;
            int slot = 0;
            if (q < n) {
                slot = tb.order[q];
                keep = q == 0 || !same_bucket(tb, slot, tb.order[q - 1]);
            }
            const unsigned bal = __ballot_sync(kFull, keep);
            if (keep) tb.tmp[kept + __popc(bal & ((1u << lane) - 1u))] = slot;
            kept += __popc(bal);
        }
        __syncwarp();
        // ---- 5. v2: first survivor on the diagonal ---------------------------------------------------
        if (V == kV2 && !p.test_mode) {
            for (int r0 = 0; r0 < kept && diag < 0; r0 += 32) {
                const int q = r0 + lane;
                bool on = false;
                if (q < kept) {
                    const int s = tb.tmp[q];
                    on = v2_on_diagonal((int)in_len, (int)out_len, tb.nt[s], tb.tot[s]);
                }
                const unsigned bal = __ballot_sync(kFull, on);
                if (bal) diag = tb.tmp[r0 + __ffs(bal) - 1];
            }
        }
    } else {
        // ---- 2-5. the first W survivors of "stable sort by log-prob descending, then drop an element equal to the one
        // before it" WITHOUT sorting everything: candidates are extracted one by one in sorted order — a warp arg-max over
        // (key descending, slot ascending), which is exactly the stable order since slots are in (w, class) order — and
        // each is kept unless it equals the last kept one (an equivalence, so "last kept" and "immediately preceding"
        // agree).  W + (duplicates met) rounds of two warp reductions instead of N^2/32 comparisons per lane.
        // Keys form a total order: -0.0 counts as +0.0 (partial_cmp says Equal); NaN, whose position the reference's
        // comparator (partial_cmp().unwrap_or(Equal), src/lib.rs:161) leaves to its sort algorithm, ranks below everything.
        auto okey = [&](int slot) -> unsigned {
            float k0 = tb.lp[slot] + 0.0f;
            k0 = (k0 != k0) ? -INFINITY : k0;
            const unsigned bits = __float_as_uint(k0);
            return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);   // order-preserving image of the float
        };
        // this lane's best live candidate among its slots lane, lane+32, ...  (0 / -1 if none)
        auto lane_best = [&](unsigned& bk, int& bs) {
            bk = 0u; bs = -1;
            for (int s2 = lane; s2 < N; s2 += 32) {
                if (!tb.valid[s2]) continue;
                const unsigned k2 = okey(s2);
                if (bs < 0 || k2 > bk) { bk = k2; bs = s2; }   // ascending scan: the first of equal keys stays
            }
        };
        // v2: the first survivor on the diagonal = the best on-diagonal candidate (it cannot be a duplicate of the element
        // before it, which would be on the diagonal too and earlier)
        if (V == kV2 && !p.test_mode) {
            unsigned bk = 0u; int bs = -1;
            for (int s2 = lane; s2 < N; s2 += 32) {
                if (!tb.valid[s2] || !v2_on_diagonal((int)in_len, (int)out_len, tb.nt[s2], tb.tot[s2])) continue;
                const unsigned k2 = okey(s2);
                if (bs < 0 || k2 > bk) { bk = k2; bs = s2; }
            }
            const unsigned top = __reduce_max_sync(kFull, bs >= 0 ? bk : 0u);
            const int first = (int)__reduce_min_sync(kFull, (bs >= 0 && bk == top) ? (unsigned)bs : 0xffffffffu);
            diag = __any_sync(kFull, bs >= 0) ? first : -1;
        }
        unsigned bk; int bs;
        lane_best(bk, bs);
        int last = -1;
        for (;;) {
            if (!__any_sync(kFull, bs >= 0)) break;   // every candidate extracted
            const unsigned top = __reduce_max_sync(kFull, bs >= 0 ? bk : 0u);
            const int slot = (int)__reduce_min_sync(kFull, (bs >= 0 && bk == top) ? (unsigned)bs : 0xffffffffu);
            if (last < 0 || !same_bucket(tb, slot, last)) {
                if (lane == 0) tb.tmp[kept] = slot;
                last = slot;
                if (++kept == W) break;
            }
            if (slot % 32 == lane) {   // the owner retires it and looks for its next best
                tb.valid[slot] = 0;
                lane_best(bk, bs);
            }
        }
        __syncwarp();
        if (kept == 0) {  // src/v2.rs:292 assert_ne! / `i % 0` in src/tone_latent.rs:199
            if (lane == 0) atomicOr(p.err, V == kV2 ? kErrV2EmptyBeam : kErrToneEmptyBeam);
            return false;
        }
    }
    // ---- 6. pad cyclically, truncate, append the diagonal candidate last -------------------------
    for (int i = lane; i < W; i += 32) {
        int s = tb.tmp[i % kept];
        if (diag >= 0 && i == W - 1) s = diag;
        r.prediction[i] = tb.pred[s];
        r.log_probs[i] = tb.lp[s];
        r.next_t[i] = tb.nt[s];
        r.next_u[i] = tb.nu[s];
        r.next_fin[i] = tb.fin[s] != 0;
        r.parent[i] = tb.par[s];
        if (V == kV2) r.next_total[i] = tb.tot[s];
    }
    return true;
}

// One step for every batch entry: one warp per entry, blockDim.x / 32 entries per block.
template <int V>
__global__ void beam_step_kernel(const BeamParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + wib;
    if (b >= p.B) return;
    const int W = p.W, N = W * p.C;
    const Table tb = carve_table(smem + (size_t)wib * table_bytes(N), N);
    const size_t o = (size_t)b * W;
    BeamRow r;
    r.h = p.h + (size_t)b * N;
    r.hist = p.hist + o; r.fin = p.fin + o; r.total = V == kV2 ? p.total + o : nullptr; r.t = p.t + o; r.u = p.u + o;
    r.prediction = p.prediction + o; r.log_probs = p.log_probs + o; r.next_t = p.next_t + o; r.next_u = p.next_u + o;
    r.next_fin = p.next_fin + o; r.next_total = V == kV2 ? p.next_total + o : nullptr; r.parent = p.parent + o;
    const long long in_len = V == kV1 ? (long long)p.max_t : (long long)p.in_len[b];
    const long long out_len = V == kV2 ? (long long)p.out_len[b] : 0;
    beam_step_warp<V>(p, tb, r, in_len, out_len, lane);
}

template <int V>
void launch(const BeamParams& p, cudaStream_t stream) {
    if (p.B <= 0 || p.W <= 0) return;
    const size_t per_warp = table_bytes((size_t)p.W * p.C);
    // Warps (batch entries) per block: four when their candidate tables fit the shared memory together, else two or one.
    // One warp's table holds beam_width * classes up to ~6100 candidates; the reference has no limit, this library does.
    int wpb = kWarpsPerBlock;
    while (wpb > 1 && per_warp * wpb > 227 * 1024) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    SSNT_ASSERT(smem <= 227 * 1024, "beam step: beam_width * classes exceeds 6100 candidates (shared-memory candidate table)");
    if (smem > 48 * 1024)
        SSNT_CUDA(cudaFuncSetAttribute(beam_step_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (p.B + wpb - 1) / wpb;
    beam_step_kernel<V><<<blocks, 32 * wpb, smem, stream>>>(p);
    SSNT_CUDA(cudaGetLastError());
}

// ---- whole-loop decoding -------------------------------------------------------------------
// The reference's ABI is one call per output step (ssnt_tts_tensorflow/__init__.py:33-73, meant for a tf.while_loop),
// followed by order_beam_branch and upsample_source_indexes (src/v2_util.rs:6-66).  When the per-step scores are known
// up front — h [B, S, W, C] — one launch does all of it: one warp per batch entry runs the S steps with the beam state
// in shared memory, records prediction and beam_branch per step, then walks every final beam back (final_branch =
// 0..W-1), gathers the durations along each branch and expands them into source indexes.  Step for step it is
// beam_step_warp above, so the results equal the per-step calls bit for bit.
struct LoopParams {
    BeamParams bp;            // h = [B, S, W, C]; hist/fin/total/t/u = initial state [B, W] (null: zeros / false)
    int S, max_u;
    int* pred_hist;           // [B, S, W]
    int* branch_hist;         // [B, S, W]
    int* ordered;             // [B, W, S]
    int* ordered_pred;        // [B, W, S]  prediction along the branch (v2: mapped through the duration table → durations)
    int* upsampled;           // [B, W, max_u] (v2, may be null), caller pre-filled
    int hist_in_smem;         // the two histories are also kept in shared memory (back-trace without global round trips)
};

template <int V>
__global__ void __launch_bounds__(32) decode_loop_kernel(const LoopParams lp) {
    extern __shared__ __align__(16) unsigned char smem[];
    const BeamParams& p = lp.bp;
    const int lane = threadIdx.x, b = blockIdx.x;
    const int W = p.W, C = p.C, N = W * C, S = lp.S;
    const Table tb = carve_table(smem, N);
    // state, double-buffered: [2] x {hist f32[W], total i32[W], t i32[W], u i32[W], fin u8[W]}
    unsigned char* sp = smem + table_bytes(N);
    const size_t WP = (size_t)((W + 3) & ~3);
    float* s_hist = reinterpret_cast<float*>(sp);
    int* s_total = reinterpret_cast<int*>(s_hist + 2 * WP);
    int* s_t = s_total + 2 * WP;
    int* s_u = s_t + 2 * WP;
    bool* s_fin = reinterpret_cast<bool*>(s_u + 2 * WP);
    float* s_h = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_fin) + 2 * WP);   // [2][N] scores of this / the next step
    int* s_ph = reinterpret_cast<int*>(s_h + 2 * (size_t)N);  // [S, W] if hist_in_smem
    int* s_bh = s_ph + (lp.hist_in_smem ? (size_t)S * W : 0);
    const size_t ob = (size_t)b * W;
    for (int w = lane; w < W; w += 32) {
        s_hist[w] = p.hist ? p.hist[ob + w] : 0.0f;
        s_total[w] = (V == kV2 && p.total) ? p.total[ob + w] : 0;
        s_t[w] = p.t ? p.t[ob + w] : 0;
        s_u[w] = p.u ? p.u[ob + w] : 0;
        s_fin[w] = p.fin ? p.fin[ob + w] : false;
    }
    // the step's scores travel global -> shared memory asynchronously, one step ahead of their use
    const float* hb = p.h + (size_t)b * S * N;
    auto fetch = [&](int s) {
        float* dst = s_h + (size_t)(s & 1) * N;
        const float* src = hb + (size_t)s * N;
        for (int i = lane; i < N; i += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst + i)), "l"(src + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch(0);
    __syncwarp();
    const long long in_len = (long long)p.in_len[b];
    const long long out_len = V == kV2 ? (long long)p.out_len[b] : 0;
    int* gph = lp.pred_hist + (size_t)b * S * W;
    int* gbh = lp.branch_hist + (size_t)b * S * W;
    int cur = 0;
    for (int s = 0; s < S; ++s) {
        const int nxt = cur ^ 1;
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (s + 1 < S) fetch(s + 1);
        BeamRow r;
        r.h = s_h + (size_t)(s & 1) * N;
        r.hist = s_hist + cur * WP; r.fin = s_fin + cur * WP; r.total = s_total + cur * WP; r.t = s_t + cur * WP; r.u = s_u + cur * WP;
        r.prediction = lp.hist_in_smem ? s_ph + (size_t)s * W : gph + (size_t)s * W;
        r.parent = lp.hist_in_smem ? s_bh + (size_t)s * W : gbh + (size_t)s * W;
        r.log_probs = s_hist + nxt * WP; r.next_t = s_t + nxt * WP; r.next_u = s_u + nxt * WP;
        r.next_fin = s_fin + nxt * WP; r.next_total = s_total + nxt * WP;
        if (!beam_step_warp<V>(p, tb, r, in_len, out_len, lane)) {  // flag raised; the reference would have panicked here
            asm volatile("cp.async.wait_all;" ::: "memory");
            return;
        }
        __syncwarp();
        if (lp.hist_in_smem)
            for (int w = lane; w < W; w += 32) {
                gph[(size_t)s * W + w] = s_ph[(size_t)s * W + w];
                gbh[(size_t)s * W + w] = s_bh[(size_t)s * W + w];
            }
        cur = nxt;
    }
    // final state
    for (int w = lane; w < W; w += 32) {
        p.log_probs[ob + w] = s_hist[cur * WP + w];
        p.next_t[ob + w] = s_t[cur * WP + w];
        p.next_u[ob + w] = s_u[cur * WP + w];
        p.next_fin[ob + w] = s_fin[cur * WP + w];
        if (V == kV2) p.next_total[ob + w] = s_total[cur * WP + w];
    }
    __syncwarp();
    __threadfence_block();
    // back-trace of every final beam (src/v2_util.rs:26-36 with final_branch = w): a lane per beam
    for (int w = lane; w < W; w += 32) {
        int c = w;
        int* ord = lp.ordered + ((size_t)b * W + w) * S;
        int* opr = lp.ordered_pred + ((size_t)b * W + w) * S;
        for (int s = S - 1; s >= 0; --s) {
            const int pr = lp.hist_in_smem ? s_ph[(size_t)s * W + c] : __ldcg(gph + (size_t)s * W + c);
            const int par = lp.hist_in_smem ? s_bh[(size_t)s * W + c] : __ldcg(gbh + (size_t)s * W + c);
            ord[s] = c;
            opr[s] = V == kV2 ? p.dur_table[pr] : pr;
            c = par;  // always in [0, W): written by beam_step_warp
        }
    }
    if (V != kV2 || lp.upsampled == nullptr) return;
    __syncwarp();
    __threadfence_block();
    // source indexes: index s repeated duration[s] times (src/v2_util.rs:39-66); length = the beam's total duration
    for (int w = 0; w < W; ++w) {
        const int* d = lp.ordered_pred + ((size_t)b * W + w) * S;
        const int len = s_total[cur * WP + w];
        int total = 0;
        bool bad = false;
        for (int s0 = 0; s0 < S; s0 += 32) {
            const int v = s0 + lane < S ? __ldcg(d + s0 + lane) : 0;
            bad |= v < 0;
            total += v > 0 ? v : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
        if (__any_sync(kFull, bad) || total != len) {  // assert_eq!(upsampled.len(), output_length[0]) src/v2_util.rs:58
            if (lane == 0) atomicOr(p.err, kErrUpsampleLength);
            continue;
        }
        int* out = lp.upsampled + ((size_t)b * W + w) * lp.max_u;
        int base = 0;
        for (int s0 = 0; s0 < S; s0 += 32) {
            const int s = s0 + lane;
            const int v = s < S ? __ldcg(d + s) : 0;
            int x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(kFull, x, o);
                if (lane >= o) x += y;
            }
            const int start = base + x - v;
            for (int k = 0; k < v && start + k < lp.max_u; ++k) out[start + k] = s;
            base += __shfl_sync(kFull, x, 31);
        }
    }
}

template <int V>
void launch_loop(LoopParams lp, cudaStream_t stream) {
    const BeamParams& p = lp.bp;
    if (p.B <= 0 || p.W <= 0 || lp.S <= 0) return;
    const size_t WP = (size_t)((p.W + 3) & ~3);
    const size_t base = table_bytes((size_t)p.W * p.C) + 2 * WP * (4 * 4 + 1) + 16 + (size_t)2 * p.W * p.C * sizeof(float);
    const size_t hist = (size_t)2 * lp.S * p.W * sizeof(int);
    lp.hist_in_smem = base + hist <= 200 * 1024 ? 1 : 0;
    const size_t smem = base + (lp.hist_in_smem ? hist : 0);
    SSNT_ASSERT(smem <= 227 * 1024, "decode loop: beam_width * classes exceeds 6000 candidates (shared-memory candidate table)");
    if (smem > 48 * 1024)
        SSNT_CUDA(cudaFuncSetAttribute(decode_loop_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decode_loop_kernel<V><<<p.B, 32, smem, stream>>>(lp);
    SSNT_CUDA(cudaGetLastError());
}

}  // namespace

// Device-pointer compute layer (the analogue of the Rust crate's trait objects; the C-ABI in
// c_api.cu marshals into these exactly as ssnt_tts_c/src/lib.rs marshals into the crate).
void v1_beam_search_decode(const float* h, const float* hist, const bool* fin, const int* t,
                           const int* u, int batch_size, int max_t, int beam_width, int* prediction,
                           float* log_probs, int* next_t, int* next_u, bool* next_fin, int* parent,
                           cudaStream_t stream) {
    BeamParams p{};
    p.h = h; p.hist = hist; p.fin = fin; p.t = t; p.u = u;
    p.B = batch_size; p.W = beam_width; p.C = 2; p.max_t = max_t;
    p.prediction = prediction; p.log_probs = log_probs; p.next_t = next_t; p.next_u = next_u;
    p.next_fin = next_fin; p.parent = parent; p.err = device_error_flag();
    launch<kV1>(p, stream);
}

void v2_beam_search_decode(const float* h, const float* hist, const bool* fin, const int* total,
                           const int* dur_table, const int* t, const int* u, const int* in_len,
                           const int* out_len, int batch_size, int beam_width, int classes,
                           int zero_duration_id, bool allow_skip, bool test_mode, int* prediction,
                           float* log_probs, int* next_t, int* next_u, bool* next_fin,
                           int* next_total, int* parent, cudaStream_t stream) {
    BeamParams p{};
    p.h = h; p.hist = hist; p.fin = fin; p.total = total; p.dur_table = dur_table; p.t = t; p.u = u;
    p.in_len = in_len; p.out_len = out_len;
    p.B = batch_size; p.W = beam_width; p.C = classes; p.special_id = zero_duration_id;
    p.allow_skip = allow_skip; p.test_mode = test_mode;
    p.prediction = prediction; p.log_probs = log_probs; p.next_t = next_t; p.next_u = next_u;
    p.next_fin = next_fin; p.next_total = next_total; p.parent = parent; p.err = device_error_flag();
    SSNT_ASSERT(classes > 0, "duration_class_size must be positive");
    launch<kV2>(p, stream);
}

void tone_beam_search_decode(const float* h, const float* hist, const bool* fin, const int* t,
                             const int* u, const int* in_len, int batch_size, int beam_width,
                             int classes, int empty_tone_id, int* prediction, float* log_probs,
                             int* next_t, int* next_u, bool* next_fin, int* parent,
                             cudaStream_t stream) {
    BeamParams p{};
    p.h = h; p.hist = hist; p.fin = fin; p.t = t; p.u = u; p.in_len = in_len;
    p.B = batch_size; p.W = beam_width; p.C = classes; p.special_id = empty_tone_id;
    p.prediction = prediction; p.log_probs = log_probs; p.next_t = next_t; p.next_u = next_u;
    p.next_fin = next_fin; p.parent = parent; p.err = device_error_flag();
    SSNT_ASSERT(classes > 0, "tone_class_size must be positive");
    launch<kTone>(p, stream);
}

void v2_decode_loop(const float* h, const int* dur_table, const int* in_len, const int* out_len, const float* hist0,
                    const bool* fin0, const int* total0, const int* t0, const int* u0, int batch_size, int steps,
                    int beam_width, int classes, int zero_duration_id, bool allow_skip, bool test_mode, int max_u,
                    int* pred_hist, int* branch_hist, float* log_probs, int* final_t, int* final_u, bool* final_fin,
                    int* final_total, int* ordered, int* duration, int* upsampled, cudaStream_t stream) {
    LoopParams lp{};
    BeamParams& p = lp.bp;
    p.h = h; p.hist = hist0; p.fin = fin0; p.total = total0; p.dur_table = dur_table; p.t = t0; p.u = u0;
    p.in_len = in_len; p.out_len = out_len;
    p.B = batch_size; p.W = beam_width; p.C = classes; p.special_id = zero_duration_id;
    p.allow_skip = allow_skip; p.test_mode = test_mode;
    p.log_probs = log_probs; p.next_t = final_t; p.next_u = final_u; p.next_fin = final_fin; p.next_total = final_total;
    p.err = device_error_flag();
    lp.S = steps; lp.max_u = max_u;
    lp.pred_hist = pred_hist; lp.branch_hist = branch_hist; lp.ordered = ordered; lp.ordered_pred = duration;
    lp.upsampled = upsampled;
    SSNT_ASSERT(classes > 0, "duration_class_size must be positive");
    launch_loop<kV2>(lp, stream);
}

void tone_decode_loop(const float* h, const int* in_len, const float* hist0, const bool* fin0, const int* t0, const int* u0,
                      int batch_size, int steps, int beam_width, int classes, int empty_tone_id, int* pred_hist,
                      int* branch_hist, float* log_probs, int* final_t, int* final_u, bool* final_fin, int* ordered,
                      int* ordered_tone, cudaStream_t stream) {
    LoopParams lp{};
    BeamParams& p = lp.bp;
    p.h = h; p.hist = hist0; p.fin = fin0; p.t = t0; p.u = u0; p.in_len = in_len;
    p.B = batch_size; p.W = beam_width; p.C = classes; p.special_id = empty_tone_id;
    p.log_probs = log_probs; p.next_t = final_t; p.next_u = final_u; p.next_fin = final_fin;
    p.err = device_error_flag();
    lp.S = steps; lp.max_u = 0;
    lp.pred_hist = pred_hist; lp.branch_hist = branch_hist; lp.ordered = ordered; lp.ordered_pred = ordered_tone;
    lp.upsampled = nullptr;
    SSNT_ASSERT(classes > 0, "tone_class_size must be positive");
    launch_loop<kTone>(lp, stream);
}

}  // namespace ssnt
