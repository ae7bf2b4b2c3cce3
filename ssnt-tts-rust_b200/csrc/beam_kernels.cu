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

template <int V>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) beam_step_kernel(const BeamParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int b = blockIdx.x * kWarpsPerBlock + wib;
    if (b >= p.B) return;
    const int W = p.W, C = p.C, N = W * C;
    // carve this warp's slice
    const size_t per_warp = (size_t)N * (9 * 4 + 2 * 1) + 16;
    unsigned char* base = smem + (size_t)wib * ((per_warp + 15) & ~(size_t)15);
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

    const float* h = p.h + (size_t)b * N;
    const float* hist = p.hist + (size_t)b * W;
    const bool* fin = p.fin + (size_t)b * W;
    const int* tt = p.t + (size_t)b * W;
    const int* uu = p.u + (size_t)b * W;
    const long long in_len = V == kV1 ? (long long)p.max_t : (long long)p.in_len[b];
    const long long out_len = V == kV2 ? (long long)p.out_len[b] : 0;

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
                tot = V == kV2 ? p.total[(size_t)b * W + w] : 0;
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
            tot = p.total[(size_t)b * W + w] + duration;
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

    // ---- 2. compact valid slots, keeping (w, class) order -----------------------------------
    int n = 0;
    for (int s0 = 0; s0 < N; s0 += 32) {
        const int s = s0 + lane;
        const bool v = s < N && tb.valid[s];
        const unsigned bal = __ballot_sync(kFull, v);
        if (v) {
            const int k = n + __popc(bal & ((1u << lane) - 1u));
            tb.tmp[k] = s;
            tb.key[k] = tb.lp[s];
        }
        n += __popc(bal);
    }
    __syncwarp();
    if (n == 0) {  // src/v2.rs:292 assert_ne! / `i % 0` in src/tone_latent.rs:199
        if (lane == 0) atomicOr(p.err, V == kV2 ? kErrV2EmptyBeam : kErrToneEmptyBeam);
        return;
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
    int kept = 0;
    for (int r0 = 0; r0 < n; r0 += 32) {
        const int r = r0 + lane;
        bool keep = false;
        int slot = 0;
        if (r < n) {
            slot = tb.order[r];
            keep = r == 0 || !same_bucket(tb, slot, tb.order[r - 1]);
        }
        const unsigned bal = __ballot_sync(kFull, keep);
        if (keep) tb.tmp[kept + __popc(bal & ((1u << lane) - 1u))] = slot;
        kept += __popc(bal);
    }
    __syncwarp();
    // ---- 5. v2: first survivor on the diagonal ---------------------------------------------------
    int diag = -1;
    if (V == kV2 && !p.test_mode) {
        for (int r0 = 0; r0 < kept && diag < 0; r0 += 32) {
            const int r = r0 + lane;
            bool on = false;
            if (r < kept) {
                const int s = tb.tmp[r];
                on = v2_on_diagonal((int)in_len, (int)out_len, tb.nt[s], tb.tot[s]);
            }
            const unsigned bal = __ballot_sync(kFull, on);
            if (bal) diag = tb.tmp[r0 + __ffs(bal) - 1];
        }
    }
    // ---- 6. pad cyclically, truncate, append the diagonal candidate last -------------------------
    for (int i = lane; i < W; i += 32) {
        int s = tb.tmp[i % kept];
        if (diag >= 0 && i == W - 1) s = diag;
        const size_t o = (size_t)b * W + i;
        p.prediction[o] = tb.pred[s];
        p.log_probs[o] = tb.lp[s];
        p.next_t[o] = tb.nt[s];
        p.next_u[o] = tb.nu[s];
        p.next_fin[o] = tb.fin[s] != 0;
        p.parent[o] = tb.par[s];
        if (V == kV2) p.next_total[o] = tb.tot[s];
    }
}

template <int V>
void launch(const BeamParams& p, cudaStream_t stream) {
    if (p.B <= 0 || p.W <= 0) return;
    const size_t N = (size_t)p.W * p.C;
    const size_t per_warp = ((N * (9 * 4 + 2) + 16) + 15) & ~(size_t)15;
    const size_t smem = per_warp * kWarpsPerBlock;
    SSNT_ASSERT(smem <= 227 * 1024, "beam step: beam_width * classes too large for shared memory");
    if (smem > 48 * 1024)
        SSNT_CUDA(cudaFuncSetAttribute(beam_step_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (p.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
    beam_step_kernel<V><<<blocks, 32 * kWarpsPerBlock, smem, stream>>>(p);
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

}  // namespace ssnt
