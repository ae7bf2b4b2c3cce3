// TEST INFRASTRUCTURE ONLY — NOT PRODUCT CODE.
//
// CPU oracle for the ssnt-tts-rust hot path.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library, and only as the
// checker (or the CPU arm being timed) — never as part of the shipped CUDA path.
//
// Two kinds of content live here:
//
//  (1) A line-faithful C++ restatement of the reference Rust crate (cited file:line below,
//      paths relative to the reference checkout).  The Rust crate cannot be built in this
//      image (no cargo/rustc), so there is no oracle/_ref.  This part is PINNED against the
//      reference's own golden vectors (tests/test_edit_distance.rs:9-107,
//      tests/test_decoding.rs:53-131, ssnt-tts-tensorflow/tests/
//      test_upsample_source_indexes.py:13-53) by tests/test_oracle_golden.py.
//
//  (2) The SSNT forward-backward lattice (a-FB) and its tone-latent variant (a-TL).  The
//      reference contains NO forward-backward / loss / gradient code, so
//      this is an authored specification (SURVEY.md §8 a-FB / a-TL) restated in fp32 and
//      fp64.  PARITY UNPINNED by the reference for this part: the oracle is hardened by a
//      brute-force path enumerator, finite differences and occupancy invariants instead
//      (tests/test_oracle_lattice.py).
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off, so the f32 band arithmetic of
// src/v2.rs:94-117 is evaluated op by op exactly as rustc emits it).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------
// rayon stand-in: the reference fans the batch axis out over rayon's global pool
// (src/lib.rs:122-133, src/v2.rs:227-242, src/edit_distance.rs:13-22).  Order-preserving,
// no arithmetic lives there, so a static block partition over std::thread is equivalent.
// ---------------------------------------------------------------------------------------
int g_threads = 0;  // 0 → hardware_concurrency

int pool_size() {
    if (g_threads > 0) return g_threads;
    unsigned n = std::thread::hardware_concurrency();
    return n == 0 ? 1 : (int)n;
}

// Persistent worker pool (rayon's global pool is created once, too): the workers sleep on a condition
// variable between calls, so a call does not pay for thread creation.
class Pool {
public:
    explicit Pool(int n) {
        for (int k = 0; k < n; ++k) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }
    void run(int n, const std::function<void(int)>& body) {
        std::unique_lock<std::mutex> l(mu_);
        body_ = &body;
        n_ = n;
        next_.store(0);
        pending_ = (int)workers_.size();
        ++epoch_;
        cv_.notify_all();
        done_.wait(l, [this] { return pending_ == 0; });
        body_ = nullptr;
    }

private:
    void loop() {
        unsigned seen = 0;
        for (;;) {
            const std::function<void(int)>* body;
            int n;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return stop_ || epoch_ != seen; });
                if (stop_) return;
                seen = epoch_;
                body = body_;
                n = n_;
            }
            for (;;) {
                int i = next_.fetch_add(1);
                if (i >= n) break;
                (*body)(i);
            }
            {
                std::lock_guard<std::mutex> l(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* body_ = nullptr;
    std::atomic<int> next_{0};
    int n_ = 0, pending_ = 0;
    unsigned epoch_ = 0;
    bool stop_ = false;
};

std::mutex g_pool_mu;     // one parallel_for at a time (callers are the single-threaded tests / bench)
Pool* g_pool = nullptr;   // intentionally leaked at exit: workers must not be joined from a static destructor

void parallel_for(int n, const std::function<void(int)>& body) {
    const int want = pool_size();
    if (want <= 1 || n <= 1) {
        for (int i = 0; i < n; ++i) body(i);
        return;
    }
    std::lock_guard<std::mutex> l(g_pool_mu);
    if (!g_pool || g_pool->size() != want) {
        delete g_pool;
        g_pool = new Pool(want);
    }
    g_pool->run(n, body);
}

[[noreturn]] void rust_panic(const char* msg) {
    // Rust: assert!/assert_eq! panic → unwinds into `extern fn` → abort.
    std::fprintf(stderr, "ssnt_oracle: panic: %s\n", msg);
    std::abort();
}

// Rust `f32 as i32`: truncate toward zero, saturating, NaN → 0.
inline int32_t rust_f32_as_i32(float v) {
    if (std::isnan(v)) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}

// ---------------------------------------------------------------------------------------
// Beam-search single step.  One DecodeResult type serves v1 (src/lib.rs:70-88), v2
// (src/v2.rs:169-189, adds total_duration) and tone-latent (src/tone_latent.rs:99-116).
// ---------------------------------------------------------------------------------------
struct DecodeResult {
    int32_t prediction;
    float log_prob;
    int64_t next_t;  // usize in the reference
    int64_t next_u;
    bool is_finished;
    int64_t parent_branch;
    int32_t total_duration;  // v2 only; 0 elsewhere so it never breaks equality
};

// eq_ignore_parent: src/lib.rs:80-88, src/v2.rs:180-189, src/tone_latent.rs:108-116.
inline bool eq_ignore_parent(const DecodeResult& a, const DecodeResult& b) {
    return a.prediction == b.prediction && a.log_prob == b.log_prob && a.next_t == b.next_t &&
           a.next_u == b.next_u && a.is_finished == b.is_finished &&
           a.total_duration == b.total_duration;
}

// results.sort_by(|a,b| a.log_prob.partial_cmp(&b.log_prob).unwrap_or(Equal).reverse())
// (src/lib.rs:161, src/v2.rs:280, src/tone_latent.rs:195): stable, descending.  For non-NaN
// keys `a.lp > b.lp` is the same strict weak order; NaN keys are outside the contract.
void sort_desc_stable(std::vector<DecodeResult>& r) {
    std::stable_sort(r.begin(), r.end(),
                     [](const DecodeResult& a, const DecodeResult& b) { return a.log_prob > b.log_prob; });
}

// Vec::dedup_by(|a, b| a.eq_ignore_parent(b)) (src/lib.rs:162): drop an element iff it
// equals the last *retained* one; the first of a run survives.
void dedup_consecutive(std::vector<DecodeResult>& r) {
    if (r.empty()) return;
    size_t w = 1;
    for (size_t i = 1; i < r.size(); ++i)
        if (!eq_ignore_parent(r[i], r[w - 1])) r[w++] = r[i];
    r.resize(w);
}

// ---- v1 (src/lib.rs) --------------------------------------------------------------------
// beam_search_kernel_internal, src/lib.rs:172-230 with decode_beam_at :57-67.
void v1_expand(const float* h, int w, int64_t t, int64_t u, float hist, bool finished,
               int64_t input_length, std::vector<DecodeResult>& out) {
    const bool defined = t >= 0 && t < input_length;  // is_defined_at :53-55
    if (!defined || finished) {                       // None arm :175-184
        out.push_back({0, hist, t, u, true, w, 0});
        return;
    }
    const float branch[2] = {h[w * 2 + 0], h[w * 2 + 1]};  // Emit=0, Shift=1 (:65)
    for (int c = 0; c < 2; ++c) {
        const bool last = (t == input_length - 1);
        if (c == 0 && last)
            out.push_back({0, hist + branch[0], t, u, true, w, 0});  // :187-195
        else if (c == 1 && last)
            out.push_back({0, hist, t, u, true, w, 0});  // Shift prohibited :196-205
        else if (c == 1)
            out.push_back({1, hist + branch[1], t + 1, u + 1, false, w, 0});  // :206-215
        else
            out.push_back({0, hist + branch[0], t, u + 1, false, w, 0});  // :216-225
    }
}

// beam_search_kernel, src/lib.rs:149-170.
std::vector<DecodeResult> v1_kernel(const float* h, const float* hist, const bool* fin,
                                    const int32_t* t, const int32_t* u, int64_t input_length,
                                    int W, int maxW) {
    std::vector<DecodeResult> r;
    for (int w = 0; w < W; ++w)
        v1_expand(h, w, (int64_t)t[w], (int64_t)u[w], hist[w], fin[w], input_length, r);
    sort_desc_stable(r);
    dedup_consecutive(r);
    if ((int)r.size() < maxW) {  // :163-167 — results[i] with the vector growing ≡ results[i % n]
        size_t add = maxW - r.size();
        for (size_t i = 0; i < add; ++i) r.push_back(r[i]);
    }
    r.resize(maxW);  // truncate :168
    return r;
}

// ---- v2 (src/v2.rs) ---------------------------------------------------------------------
struct V2Table {
    const float* h;
    const float* hist;
    const bool* fin;
    const int32_t* total;
    const int32_t* dur_table;
    int D;
    int64_t in_len, out_len;
    int W, maxW;
    int32_t zero_id;
};

// total_duration_bounds, src/v2.rs:94-104 (f32 arithmetic, op by op).
void v2_bounds(const V2Table& tb, int64_t t, int32_t& lo, int32_t& hi) {
    float ratio = (float)tb.out_len / (float)tb.in_len;
    float diagonal = ratio * (float)(t + 1);
    float upper_range = (float)tb.out_len * 0.1f;
    float lower_range = (float)tb.out_len * 0.05f;
    float lb = diagonal - lower_range;
    lb = std::isnan(lb) ? 0.0f : (lb > 0.0f ? lb : 0.0f);  // f32::max(0.0): NaN → the other operand
    float ub = diagonal + upper_range;
    float ol = (float)tb.out_len;
    ub = std::isnan(ub) ? ol : (ub < ol ? ub : ol);  // f32::min(out)
    lo = rust_f32_as_i32(lb);
    hi = rust_f32_as_i32(ub);
}

// will_overrun, src/v2.rs:106-111 (usize arithmetic).
bool v2_will_overrun(const V2Table& tb, int64_t t) {
    uint64_t remaining = (uint64_t)(tb.in_len - (t + 1));
    uint64_t min_total = remaining * 3u;
    return min_total > (uint64_t)tb.out_len;
}

// on_diagonal, src/v2.rs:113-117.
bool v2_on_diagonal(const V2Table& tb, const DecodeResult& r) {
    float ratio = (float)tb.out_len / (float)tb.in_len;
    float diagonal = ratio * (float)r.next_t;
    float diff = (float)r.total_duration - diagonal;
    return diff >= -20.0f && diff <= 0.0f;
}

// beam_search_kernel_internal src/v2.rs:311-339 + decode_beam_at :119-166.
void v2_expand(const V2Table& tb, int w, int64_t t, int64_t u, bool allow_skip, bool test_mode,
               std::vector<DecodeResult>& out) {
    const float hist = tb.hist[w];
    const bool defined = t >= 0 && t < tb.in_len;  // :90-92 (usize: negative t wraps to huge → false)
    if (!defined || tb.fin[w]) {                   // :314-323
        out.push_back({tb.zero_id, hist, t, u, true, w, tb.total[w]});
        return;
    }
    for (int i = 0; i < tb.D; ++i) {
        const float v = tb.h[w * tb.D + i];
        const int32_t duration = tb.dur_table[i];
        const int32_t total = tb.total[w] + duration;
        int32_t lo, hi;
        v2_bounds(tb, t, lo, hi);
        bool finished;
        if (!test_mode && (total < lo || total > hi)) continue;      // :131-132
        else if (!test_mode && v2_will_overrun(tb, t)) continue;     // :133-134
        else if (t == tb.in_len - 1) {                               // :135-150
            if (!test_mode && total != (int32_t)tb.out_len) continue;
            if (!allow_skip && i == tb.zero_id) continue;
            finished = true;
        } else {                                                     // :151-162
            if (!allow_skip && i == tb.zero_id) continue;
            finished = false;
        }
        out.push_back({i, hist + v, finished ? t : t + 1, finished ? u : u + 1, finished, w, total});  // :326-336
    }
}

// beam_search_kernel src/v2.rs:269-309.  Returns false when n_results == 0 (assert_ne! :292).
bool v2_kernel(const V2Table& tb, const int32_t* t, const int32_t* u, bool allow_skip,
               bool test_mode, std::vector<DecodeResult>& r) {
    r.clear();
    for (int w = 0; w < tb.W; ++w) v2_expand(tb, w, (int64_t)t[w], (int64_t)u[w], allow_skip, test_mode, r);
    sort_desc_stable(r);
    dedup_consecutive(r);
    bool have_diag = false;
    DecodeResult diag{};
    if (!test_mode)
        for (const auto& x : r)
            if (v2_on_diagonal(tb, x)) { diag = x; have_diag = true; break; }  // :283-289
    const size_t n = r.size();
    if (n == 0) return false;
    if ((int)n < tb.maxW)
        for (size_t i = 0; i < tb.maxW - n; ++i) r.push_back(r[i % n]);  // :293-297
    if (have_diag) {  // :298-303
        r.resize(tb.maxW - 1);
        r.push_back(diag);
    } else {
        r.resize(tb.maxW);
    }
    return true;
}

// ---- tone latent (src/tone_latent.rs) -----------------------------------------------------
// beam_search_kernel_internal :208-234 + decode_beam_at :79-95.
void tone_expand(const float* h, const float* hist, const bool* fin, int K, int64_t in_len,
                 int32_t empty_id, int w, int64_t t, int64_t u, std::vector<DecodeResult>& out) {
    const bool defined = t >= 0 && t < in_len;
    if (!defined || fin[w]) {
        out.push_back({empty_id, hist[w], t, u, true, w, 0});  // :211-219
        return;
    }
    for (int i = 0; i < K; ++i)  // is_finished always false (:87-93) → t+1,u+1 (:225-229)
        out.push_back({i, hist[w] + h[w * K + i], t + 1, u + 1, false, w, 0});
}

// beam_search_kernel src/tone_latent.rs:184-206.
bool tone_kernel(const float* h, const float* hist, const bool* fin, const int32_t* t,
                 const int32_t* u, int K, int64_t in_len, int32_t empty_id, int W, int maxW,
                 std::vector<DecodeResult>& r) {
    r.clear();
    for (int w = 0; w < W; ++w) tone_expand(h, hist, fin, K, in_len, empty_id, w, t[w], u[w], r);
    sort_desc_stable(r);
    dedup_consecutive(r);
    const size_t n = r.size();
    if (n == 0) return false;  // `i % n_results` would divide by zero → panic
    if ((int)n < maxW)
        for (size_t i = 0; i < maxW - n; ++i) r.push_back(r[i % n]);
    r.resize(maxW);
    return true;
}

// ---------------------------------------------------------------------------------------
// Lattice forward-backward (authored spec, SURVEY.md §8 a-FB).  T = output frames (serial
// axis), U = input tokens.  Per frame the path Emits (stays on u) or Shifts (u→u+1); Shift
// from the last token is prohibited and the last frame must be an Emit at (T-1, U-1) —
// lifted from the decoding rules src/lib.rs:187-225.
// ---------------------------------------------------------------------------------------
template <typename R>
inline R neg_inf() { return -std::numeric_limits<R>::infinity(); }

template <typename R>
inline R logaddexp(R a, R b) {
    if (a == neg_inf<R>()) return b;
    if (b == neg_inf<R>()) return a;
    R m = a > b ? a : b;
    R d = a > b ? b - a : a - b;  // -|a-b|
    return m + std::log1p(std::exp(d));
}

// One utterance.  le/ls/ge/gs: [max_t, max_u] row-major slabs; T,U the true lengths.
template <typename R>
void fb_one(const float* le, const float* ls, int T, int U, int max_t, int max_u, float* ll_out,
            float* ge, float* gs) {
    const size_t slab = (size_t)max_t * max_u;
    if (ge) std::memset(ge, 0, slab * sizeof(float));
    if (gs) std::memset(gs, 0, slab * sizeof(float));
    if (T <= 0 || U <= 0 || U > T) {  // no monotonic path with U-1 shifts in T-1 frames
        *ll_out = -std::numeric_limits<float>::infinity();
        return;
    }
    const R NI = neg_inf<R>();
    std::vector<R> alpha((size_t)T * U, NI), beta((size_t)(T + 1) * (U + 1), NI);
    auto A = [&](int t, int u) -> R& { return alpha[(size_t)t * U + u]; };
    auto Bt = [&](int t, int u) -> R& { return beta[(size_t)t * (U + 1) + u]; };
    auto LE = [&](int t, int u) -> R { return (R)le[(size_t)t * max_u + u]; };
    auto LS = [&](int t, int u) -> R { return (R)ls[(size_t)t * max_u + u]; };
    A(0, 0) = 0;
    for (int t = 1; t < T; ++t)
        for (int u = 0; u < U; ++u) {
            R stay = A(t - 1, u) + LE(t - 1, u);
            R shift = u > 0 ? A(t - 1, u - 1) + LS(t - 1, u - 1) : NI;
            A(t, u) = logaddexp(stay, shift);
        }
    const R LL = A(T - 1, U - 1) + LE(T - 1, U - 1);
    // beta(T-1,U-1) = le(T-1,U-1); beta = -inf elsewhere on the last row (last frame must emit).
    Bt(T - 1, U - 1) = LE(T - 1, U - 1);
    for (int t = T - 2; t >= 0; --t)
        for (int u = 0; u < U; ++u) {
            R stay = LE(t, u) + Bt(t + 1, u);
            R shift = u + 1 < U ? LS(t, u) + Bt(t + 1, u + 1) : NI;
            Bt(t, u) = logaddexp(stay, shift);
        }
    *ll_out = (float)LL;
    if (!(LL > NI) || !ge || !gs) return;  // LL = -inf (or NaN): gradients stay 0
    for (int t = 0; t < T; ++t)
        for (int u = 0; u < U; ++u) {
            const size_t o = (size_t)t * max_u + u;
            if (t == T - 1) {
                if (u == U - 1) ge[o] = (float)std::exp(A(t, u) + LE(t, u) - LL);
                continue;
            }
            ge[o] = (float)std::exp(A(t, u) + LE(t, u) + Bt(t + 1, u) - LL);
            if (u + 1 < U) gs[o] = (float)std::exp(A(t, u) + LS(t, u) + Bt(t + 1, u + 1) - LL);
        }
}

// Tone-latent marginalised lattice (authored spec, SURVEY.md §8 a-TL).  State (t,u,k):
// a stay keeps the token's tone k, a shift redraws the next token's tone from log_tone.
// The K tone classes mirror tone_class_size of src/tone_latent.rs:79-95.
template <typename R>
void tone_fb_one(const float* le, const float* ls, const float* ltone, int T, int U, int K,
                 int max_t, int max_u, float* ll_out, float* ge, float* gs, float* gt) {
    const size_t slab = (size_t)max_t * max_u * K;
    if (ge) std::memset(ge, 0, slab * sizeof(float));
    if (gs) std::memset(gs, 0, slab * sizeof(float));
    if (gt) std::memset(gt, 0, (size_t)max_u * K * sizeof(float));
    if (T <= 0 || U <= 0 || U > T || K <= 0) {
        *ll_out = -std::numeric_limits<float>::infinity();
        return;
    }
    const R NI = neg_inf<R>();
    std::vector<R> alpha((size_t)T * U * K, NI), beta((size_t)(T + 1) * (U + 1) * K, NI);
    std::vector<R> S((size_t)T * U, NI);         // S(t,u) = LSE_k(alpha(t,u,k)+ls(t,u,k))
    std::vector<R> Bm((size_t)(T + 1) * (U + 1), NI);  // Bm(t,u) = LSE_k(ltone(u,k)+beta(t,u,k))
    auto A = [&](int t, int u, int k) -> R& { return alpha[((size_t)t * U + u) * K + k]; };
    auto Bt = [&](int t, int u, int k) -> R& { return beta[((size_t)t * (U + 1) + u) * K + k]; };
    auto LE = [&](int t, int u, int k) -> R { return (R)le[((size_t)t * max_u + u) * K + k]; };
    auto LS = [&](int t, int u, int k) -> R { return (R)ls[((size_t)t * max_u + u) * K + k]; };
    auto LT = [&](int u, int k) -> R { return (R)ltone[(size_t)u * K + k]; };
    for (int k = 0; k < K; ++k) A(0, 0, k) = LT(0, k);
    for (int t = 0; t < T; ++t) {
        if (t > 0)
            for (int u = 0; u < U; ++u)
                for (int k = 0; k < K; ++k) {
                    R stay = A(t - 1, u, k) + LE(t - 1, u, k);
                    R shift = u > 0 ? LT(u, k) + S[(size_t)(t - 1) * U + u - 1] : NI;
                    A(t, u, k) = logaddexp(stay, shift);
                }
        for (int u = 0; u < U; ++u) {
            R acc = NI;
            for (int k = 0; k < K; ++k) acc = logaddexp(acc, A(t, u, k) + LS(t, u, k));
            S[(size_t)t * U + u] = acc;
        }
    }
    R LL = NI;
    for (int k = 0; k < K; ++k) LL = logaddexp(LL, A(T - 1, U - 1, k) + LE(T - 1, U - 1, k));
    for (int k = 0; k < K; ++k) Bt(T - 1, U - 1, k) = LE(T - 1, U - 1, k);
    for (int t = T - 1; t >= 0; --t) {
        if (t < T - 1)
            for (int u = 0; u < U; ++u)
                for (int k = 0; k < K; ++k) {
                    R stay = LE(t, u, k) + Bt(t + 1, u, k);
                    R shift = u + 1 < U ? LS(t, u, k) + Bm[(size_t)(t + 1) * (U + 1) + u + 1] : NI;
                    Bt(t, u, k) = logaddexp(stay, shift);
                }
        for (int u = 0; u < U; ++u) {
            R acc = NI;
            for (int k = 0; k < K; ++k) acc = logaddexp(acc, LT(u, k) + Bt(t, u, k));
            Bm[(size_t)t * (U + 1) + u] = acc;
        }
    }
    *ll_out = (float)LL;
    if (!(LL > NI) || !ge || !gs || !gt) return;
    std::vector<double> tone_acc((size_t)U * K, 0.0);
    for (int k = 0; k < K; ++k) tone_acc[k] = (double)std::exp(LT(0, k) + Bt(0, 0, k) - LL);
    for (int t = 0; t < T; ++t)
        for (int u = 0; u < U; ++u)
            for (int k = 0; k < K; ++k) {
                const size_t o = ((size_t)t * max_u + u) * K + k;
                if (t == T - 1) {
                    if (u == U - 1) ge[o] = (float)std::exp(A(t, u, k) + LE(t, u, k) - LL);
                    continue;
                }
                ge[o] = (float)std::exp(A(t, u, k) + LE(t, u, k) + Bt(t + 1, u, k) - LL);
                if (u + 1 < U) {
                    gs[o] = (float)std::exp(A(t, u, k) + LS(t, u, k) +
                                            Bm[(size_t)(t + 1) * (U + 1) + u + 1] - LL);
                    // entering token u+1 at frame t+1 with tone k
                    tone_acc[(size_t)(u + 1) * K + k] += (double)std::exp(
                        S[(size_t)t * U + u] + LT(u + 1, k) + Bt(t + 1, u + 1, k) - LL);
                }
            }
    for (int u = 0; u < U; ++u)
        for (int k = 0; k < K; ++k) gt[(size_t)u * K + k] = (float)tone_acc[(size_t)u * K + k];
}

}  // namespace

extern "C" {

void ssnt_oracle_set_threads(int n) { g_threads = n; }
int ssnt_oracle_get_threads() { return pool_size(); }

// ---- C-ABI twins of ssnt_tts_c/src/lib.rs (same argument order) ---------------------------

// ssnt_tts_c/src/lib.rs:10-83 (batch_size fixed to 1) → src/lib.rs:121-147.
void oracle_ssnt_tts_beam_search_decode(const float* h, const float* log_prob_history,
                                        const bool* is_finished, const int* t, const int* u,
                                        int max_t, int beam_width, int* prediction,
                                        float* log_probs, int* next_t, int* next_u,
                                        bool* next_is_finished, int* beam_branch) {
    if (!h || !log_prob_history || !is_finished || !t || !u || !prediction || !log_probs ||
        !next_t || !next_u || !next_is_finished || !beam_branch)
        rust_panic("null pointer");
    auto r = v1_kernel(h, log_prob_history, is_finished, t, u, (int64_t)max_t, beam_width, beam_width);
    for (int i = 0; i < beam_width; ++i) {
        prediction[i] = r[i].prediction;
        log_probs[i] = r[i].log_prob;
        next_t[i] = (int)r[i].next_t;
        next_u[i] = (int)r[i].next_u;
        beam_branch[i] = (int)r[i].parent_branch;
        next_is_finished[i] = r[i].is_finished;
    }
}

// ssnt_tts_c/src/lib.rs:86-116 → src/util.rs:20-33.
void oracle_ssnt_extract_best_beam_branch(int best_final_branch, const int* beam_branch,
                                          const int* t_history, int beam_width, int max_u,
                                          int* best_beam_branch, int* best_t_history) {
    if (!beam_branch || !t_history || !best_beam_branch || !best_t_history) rust_panic("null pointer");
    int cur = best_final_branch;
    for (int u = max_u - 1; u >= 0; --u) {  // rfold over rows, push_front
        best_beam_branch[u] = cur;
        best_t_history[u] = t_history[(size_t)u * beam_width + cur];
        cur = beam_branch[(size_t)u * beam_width + cur];
    }
}

// ssnt_tts_c/src/lib.rs:118-218 → src/v2.rs:221-267.
void oracle_ssnt_tts_v2_beam_search_decode(
    const float* h, const float* log_prob_history, const bool* is_finished,
    const int* total_duration, const int* duration_table, const int* t, const int* u,
    const int* input_length, const int* output_length, int batch_size, int beam_width,
    int duration_class_size, int zero_duration_id, bool allow_skip, bool test_mode,
    int* prediction, float* log_probs, int* next_t, int* next_u, bool* next_is_finished,
    int* next_total_duration, int* beam_branch) {
    if (!h || !log_prob_history || !is_finished || !total_duration || !duration_table || !t ||
        !u || !input_length || !output_length || !prediction || !log_probs || !next_t ||
        !next_u || !next_is_finished || !next_total_duration || !beam_branch)
        rust_panic("null pointer");
    const int W = beam_width, D = duration_class_size;
    std::atomic<bool> failed(false);
    parallel_for(batch_size, [&](int b) {
        V2Table tb{h + (size_t)b * W * D, log_prob_history + (size_t)b * W,
                   is_finished + (size_t)b * W, total_duration + (size_t)b * W, duration_table,
                   D, (int64_t)input_length[b], (int64_t)output_length[b], W, W, zero_duration_id};
        std::vector<DecodeResult> r;
        if (!v2_kernel(tb, t + (size_t)b * W, u + (size_t)b * W, allow_skip, test_mode, r)) {
            failed = true;
            return;
        }
        for (int i = 0; i < W; ++i) {
            const size_t o = (size_t)b * W + i;
            prediction[o] = r[i].prediction;
            log_probs[o] = r[i].log_prob;
            next_t[o] = (int)r[i].next_t;
            next_u[o] = (int)r[i].next_u;
            beam_branch[o] = (int)r[i].parent_branch;
            next_is_finished[o] = r[i].is_finished;
            next_total_duration[o] = r[i].total_duration;
        }
    });
    if (failed)
        rust_panic("Beam search could not find a duration sequence with compatible output length "
                   "(src/v2.rs:292)");
}

// Same as above but reports the src/v2.rs:292 panic as a return code (tests use it to check
// that the CUDA path flags the same condition without killing the test process).
int oracle_ssnt_tts_v2_beam_search_decode_checked(
    const float* h, const float* log_prob_history, const bool* is_finished,
    const int* total_duration, const int* duration_table, const int* t, const int* u,
    const int* input_length, const int* output_length, int batch_size, int beam_width,
    int duration_class_size, int zero_duration_id, bool allow_skip, bool test_mode,
    int* prediction, float* log_probs, int* next_t, int* next_u, bool* next_is_finished,
    int* next_total_duration, int* beam_branch) {
    const int W = beam_width, D = duration_class_size;
    int bad = 0;
    for (int b = 0; b < batch_size; ++b) {
        V2Table tb{h + (size_t)b * W * D, log_prob_history + (size_t)b * W,
                   is_finished + (size_t)b * W, total_duration + (size_t)b * W, duration_table,
                   D, (int64_t)input_length[b], (int64_t)output_length[b], W, W, zero_duration_id};
        std::vector<DecodeResult> r;
        if (!v2_kernel(tb, t + (size_t)b * W, u + (size_t)b * W, allow_skip, test_mode, r)) {
            ++bad;
            continue;
        }
        for (int i = 0; i < W; ++i) {
            const size_t o = (size_t)b * W + i;
            prediction[o] = r[i].prediction;
            log_probs[o] = r[i].log_prob;
            next_t[o] = (int)r[i].next_t;
            next_u[o] = (int)r[i].next_u;
            beam_branch[o] = (int)r[i].parent_branch;
            next_is_finished[o] = r[i].is_finished;
            next_total_duration[o] = r[i].total_duration;
        }
    }
    return bad;
}

// ssnt_tts_c/src/lib.rs:220-241 → src/v2_util.rs:6-36.
void oracle_ssnt_order_beam_branch(const int* final_branch, const int* beam_branch,
                                   int batch_size, int beam_width, int max_t,
                                   int* ordered_beam_branch) {
    if (!final_branch || !beam_branch || !ordered_beam_branch) rust_panic("null pointer");
    parallel_for(batch_size, [&](int b) {
        const int* bb = beam_branch + (size_t)b * max_t * beam_width;  // (T, W)
        for (int w = 0; w < beam_width; ++w) {
            int cur = final_branch[(size_t)b * beam_width + w];
            int* out = ordered_beam_branch + ((size_t)b * beam_width + w) * max_t;  // (W, T)
            for (int t = max_t - 1; t >= 0; --t) {
                out[t] = cur;
                cur = bb[(size_t)t * beam_width + cur];
            }
        }
    });
}

// ssnt_tts_c/src/lib.rs:244-265 → src/v2_util.rs:39-66.  Returns through panic like Rust.
static int upsample_impl(const int* duration, const int* output_length, int batch_size,
                         int beam_width, int max_t, int max_u, int* out, bool checked) {
    std::atomic<int> bad(0);
    parallel_for(batch_size * beam_width, [&](int bw) {
        const int* d = duration + (size_t)bw * max_t;
        const int len = output_length[bw];
        int64_t total = 0;
        for (int t = 0; t < max_t; ++t) {
            if (d[t] < 0) { ++bad; return; }  // `*d as usize` of a negative → capacity overflow panic
            total += d[t];
        }
        if (total != (int64_t)len) { ++bad; return; }  // assert_eq! :58
        int* o = out + (size_t)bw * max_u;
        const int lim = std::min(len, max_u);  // zip with the max_u-long chunk, .take(len)
        int pos = 0;
        for (int t = 0; t < max_t && pos < lim; ++t)
            for (int k = 0; k < d[t] && pos < lim; ++k) o[pos++] = t;
    });
    if (bad && !checked) rust_panic("upsample: sum(duration) != output_length (src/v2_util.rs:58)");
    return bad;
}

void oracle_ssnt_upsample_source_indexes(const int* duration, const int* output_length,
                                         int batch_size, int beam_width, int max_t, int max_u,
                                         int* upsampled_source_indexes) {
    if (!duration || !output_length || !upsampled_source_indexes) rust_panic("null pointer");
    upsample_impl(duration, output_length, batch_size, beam_width, max_t, max_u,
                  upsampled_source_indexes, false);
}

int oracle_ssnt_upsample_source_indexes_checked(const int* duration, const int* output_length,
                                                int batch_size, int beam_width, int max_t,
                                                int max_u, int* upsampled_source_indexes) {
    return upsample_impl(duration, output_length, batch_size, beam_width, max_t, max_u,
                         upsampled_source_indexes, true);
}

// ssnt_tts_c/src/lib.rs:267-343 → src/tone_latent.rs:144-182.
void oracle_tone_latent_beam_search_decode(const float* h, const float* log_prob_history,
                                           const bool* is_finished, const int* t, const int* u,
                                           const int* input_length, int batch_size,
                                           int beam_width, int tone_class_size, int empty_tone_id,
                                           int* prediction, float* log_probs, int* next_t,
                                           int* next_u, bool* next_is_finished, int* beam_branch) {
    if (!h || !log_prob_history || !is_finished || !t || !u || !input_length || !prediction ||
        !log_probs || !next_t || !next_u || !next_is_finished || !beam_branch)
        rust_panic("null pointer");
    const int W = beam_width, K = tone_class_size;
    std::atomic<bool> failed(false);
    parallel_for(batch_size, [&](int b) {
        std::vector<DecodeResult> r;
        if (!tone_kernel(h + (size_t)b * W * K, log_prob_history + (size_t)b * W,
                         is_finished + (size_t)b * W, t + (size_t)b * W, u + (size_t)b * W, K,
                         (int64_t)input_length[b], empty_tone_id, W, W, r)) {
            failed = true;
            return;
        }
        for (int i = 0; i < W; ++i) {
            const size_t o = (size_t)b * W + i;
            prediction[o] = r[i].prediction;
            log_probs[o] = r[i].log_prob;
            next_t[o] = (int)r[i].next_t;
            next_u[o] = (int)r[i].next_u;
            beam_branch[o] = (int)r[i].parent_branch;
            next_is_finished[o] = r[i].is_finished;
        }
    });
    if (failed) rust_panic("tone_latent: empty candidate set (i % 0)");
}

// ssnt_tts_c/src/lib.rs:346-381 → src/edit_distance.rs:6-60.
static int32_t edit_kernel(const int32_t* a, int M, const int32_t* b, int N) {
    std::vector<int32_t> e(N + 1), e_tmp(N + 1, -1);
    for (int n = 0; n <= N; ++n) e[n] = n;  // :31
    for (int m = 1; m <= M; ++m) {
        e_tmp[0] = e[0] + 1;
        for (int n = 1; n <= N; ++n) {
            int32_t term1 = e[n - 1] + (a[m - 1] == b[n - 1] ? 0 : 1);
            int32_t term2 = e[n] + 1;
            int32_t term3 = e_tmp[n - 1] + 1;
            e_tmp[n] = std::min(term1, std::min(term2, term3));
        }
        e = e_tmp;  // :46
    }
    return e[N];
}

void oracle_tone_latent_levenshtein_edit_distance(const int* a, const int* b,
                                                  const int* a_lengths, const int* b_lengths,
                                                  int batch_size, int max_length, int* distance) {
    if (!a || !b || !a_lengths || !b_lengths || !distance) rust_panic("null pointer");
    parallel_for(batch_size, [&](int i) {
        distance[i] = edit_kernel(a + (size_t)i * max_length, a_lengths[i],
                                  b + (size_t)i * max_length, b_lengths[i]);
    });
}

// ---- lattice (authored spec) ---------------------------------------------------------------
// precision: 0 = fp32 arithmetic ("port"), 1 = fp64 arithmetic (truth for the parity tests).
// t_len / u_len may be NULL → full lengths.  grad_* may be NULL → log-likelihood only.
// loss[0] = -sum_b ll[b] accumulated in double.
void oracle_ssnt_tts_forward_backward(const float* log_emit, const float* log_shift,
                                      const int* t_len, const int* u_len, int batch_size,
                                      int max_t, int max_u, int precision, float* log_likelihood,
                                      float* loss, float* grad_emit, float* grad_shift) {
    const size_t slab = (size_t)max_t * max_u;
    parallel_for(batch_size, [&](int b) {
        const int T = t_len ? t_len[b] : max_t, U = u_len ? u_len[b] : max_u;
        float* ge = grad_emit ? grad_emit + b * slab : nullptr;
        float* gs = grad_shift ? grad_shift + b * slab : nullptr;
        if (precision)
            fb_one<double>(log_emit + b * slab, log_shift + b * slab, T, U, max_t, max_u,
                           log_likelihood + b, ge, gs);
        else
            fb_one<float>(log_emit + b * slab, log_shift + b * slab, T, U, max_t, max_u,
                          log_likelihood + b, ge, gs);
    });
    if (loss) {
        double acc = 0;
        for (int b = 0; b < batch_size; ++b) acc -= (double)log_likelihood[b];
        loss[0] = (float)acc;
    }
}

void oracle_tone_latent_forward_backward(const float* log_emit, const float* log_shift,
                                         const float* log_tone, const int* t_len,
                                         const int* u_len, int batch_size, int max_t, int max_u,
                                         int tone_class_size, int precision,
                                         float* log_likelihood, float* loss, float* grad_emit,
                                         float* grad_shift, float* grad_tone) {
    const int K = tone_class_size;
    const size_t slab = (size_t)max_t * max_u * K, tslab = (size_t)max_u * K;
    parallel_for(batch_size, [&](int b) {
        const int T = t_len ? t_len[b] : max_t, U = u_len ? u_len[b] : max_u;
        float* ge = grad_emit ? grad_emit + b * slab : nullptr;
        float* gs = grad_shift ? grad_shift + b * slab : nullptr;
        float* gt = grad_tone ? grad_tone + b * tslab : nullptr;
        if (precision)
            tone_fb_one<double>(log_emit + b * slab, log_shift + b * slab, log_tone + b * tslab,
                                T, U, K, max_t, max_u, log_likelihood + b, ge, gs, gt);
        else
            tone_fb_one<float>(log_emit + b * slab, log_shift + b * slab, log_tone + b * tslab,
                               T, U, K, max_t, max_u, log_likelihood + b, ge, gs, gt);
    });
    if (loss) {
        double acc = 0;
        for (int b = 0; b < batch_size; ++b) acc -= (double)log_likelihood[b];
        loss[0] = (float)acc;
    }
}

}  // extern "C"
