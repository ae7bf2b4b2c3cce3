// Tone-latent marginalised lattice forward-backward (SURVEY.md §8 a-TL; no counterpart in the
// reference, whose src/tone_latent.rs only holds the beam step — the K tone classes and the
// "one tone per input token" structure come from tone_class_size there, :79-95).
//
// State (t, u, k): a frame either Emits (stays on token u, keeps its tone k) or Shifts (moves to
// token u+1 and draws that token's tone from log_tone[u+1, .]).
//   alpha(0,0,k) = lt(0,k)
//   alpha(t+1,u,k) = lae( alpha(t,u,k) + le(t,u,k),  lt(u,k) + S(t,u-1) ),  S(t,u) = LSE_k(alpha(t,u,k) + ls(t,u,k))
//   LL = LSE_k( alpha(T-1,U-1,k) + le(T-1,U-1,k) )
//   beta(t,u,k) = lae( le(t,u,k) + beta(t+1,u,k),  ls(t,u,k) + Bm(t+1,u+1) ),  Bm(t,u) = LSE_k(lt(u,k) + beta(t,u,k))
// One CTA per utterance, one thread per token (K tones in a loop), rows double-buffered in
// shared memory, log2 domain with integer row offsets exactly as in fb_kernels.cu.  The
// gradient w.r.t. log_tone is accumulated per (u,k) in shared memory over the backward sweep in
// frame order (deterministic).
#include "ssnt_common.cuh"

namespace ssnt {
namespace {

constexpr float kNeg = -1.0e30f;
constexpr float kNegTest = -1.0e29f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLn2 = 0.6931471805599453;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lae2(float x, float y) {
    const float m = fmaxf(x, y);
    const float n = fminf(x, y);
    return m + lg2(1.0f + ex2(n - m));
}
__device__ __forceinline__ float to_log2(float v) { return fmaxf(v * kLog2e, kNeg); }

__device__ float block_max(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float r = l < ((blockDim.x + 31) >> 5) ? red[l] : kNeg;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(kFull, r, o));
    return r;
}

// ---- asynchronous row prefetch (cp.async) -------------------------------------------------------
// The recursion is t-serial and every row needs 2·K fresh inputs per token (3·K in the backward
// sweep, with the stored alpha row): read in the loop they put one global-memory latency (~1 us)
// on the dependency chain of EVERY row.  Instead each thread copies its own tokens' values of the
// row kPF steps ahead into a shared-memory ring with cp.async and reads back only what it wrote
// itself, so no barrier is involved — only cp.async.wait_group.
constexpr int kPF = 8;  // rows in flight
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pf() { asm volatile("cp.async.wait_group %0;" ::"n"(kPF) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct ToneParams {
    ToneFbArgs a;
    float* scratch;  // [B][T][U][K] alpha~
    float* offs;     // [B][T]
    unsigned* counter;
    unsigned* fallbacks;   // cumulative count of re-run utterances (ssnt_tts_fb_fallback_count)
    const unsigned* only;  // optional [B]: run only the utterances whose word is non-zero (re-run of the
                           // ones the block-float kernel flagged); the others keep their results
};

// KT > 0: tone_class_size known at compile time (the loops over the tones unroll and their
// shared-memory loads batch); KT == 0: any K.
template <int KT>
__global__ void tone_fb_kernel(const ToneParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float red[32];
    __shared__ unsigned s_last;
    const ToneFbArgs& a = p.a;
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int max_t = a.max_t, max_u = a.max_u;
    const int K = KT > 0 ? KT : a.tone_class_size;
    int T = a.t_len ? a.t_len[b] : max_t;
    int U = a.u_len ? a.u_len[b] : max_u;
    T = min(max(T, 0), max_t);
    U = min(max(U, 0), max_u);
    const size_t slab = (size_t)max_t * max_u * K;
    const float* le = a.log_emit + (size_t)b * slab;
    const float* ls = a.log_shift + (size_t)b * slab;
    const float* lt = a.log_tone + (size_t)b * max_u * K;
    float* ge = a.grad_emit + (size_t)b * slab;
    float* gs = a.grad_shift + (size_t)b * slab;
    float* gt = a.grad_tone + (size_t)b * max_u * K;
    float* scr = p.scratch + (size_t)b * slab;
    float* offs = p.offs + (size_t)b * max_t;

    const bool infeasible = T <= 0 || U <= 0 || U > T;
    float llt = 0.0f, offA_last = 0.0f;
    bool dead = true;
    if (p.only && p.only[b] == 0) {
        // nothing to redo for this utterance; it still takes part in the loss reduction below
    } else if (infeasible) {
        if (p.only && tid == 0) atomicAdd(p.fallbacks, 1u);
        for (size_t i = tid; i < slab; i += nt) { ge[i] = 0.0f; gs[i] = 0.0f; }
        for (int i = tid; i < max_u * K; i += nt) gt[i] = 0.0f;
        if (tid == 0) a.log_likelihood[b] = -INFINITY;
    } else {
        if (p.only && tid == 0) atomicAdd(p.fallbacks, 1u);
        // shared layout: cur[(U+2)*K], nxt[(U+2)*K], srow[U+2], tone2[U*K], gacc[U*K]
        const int RW = (max_u + 2) * K;
        float* cur = sm;
        float* nxt = cur + RW;
        float* srow = nxt + RW;           // S(t,u) or Bm(t+1,u), index u+1
        float* tone2 = srow + (max_u + 2);
        float* gacc = tone2 + max_u * K;
        // [kPF + 1][3][max_u * K]: le | ls | stored alpha of a row; 16-byte aligned for cp.async
        float* soffs = gacc + max_u * K;  // [max_t] integer row offsets of the stored alpha rows
        // (aligned by index arithmetic on the __shared__ array, not by pointer casts: a pointer that went
        // through uintptr_t is a GENERIC pointer to the compiler and every ring read became an LD.E)
        float* ring = sm + (((2 * RW + (max_u + 2) + 2 * max_u * K + max_t) + 3) & ~3);
        const int RS = 3 * max_u * K;
        const bool vec16 = (K % 4 == 0) && ((max_u * K) % 4 == 0) && ((reinterpret_cast<uintptr_t>(le) | reinterpret_cast<uintptr_t>(ls) |
                                             reinterpret_cast<uintptr_t>(scr)) % 16 == 0);
        // copies this thread's tokens of row t (le, ls and, when with_alpha, the stored alpha row) into
        // ring slot t % (kPF+1); always commits a group (possibly empty) so that group counts line up
        auto issue_row = [&](int t, bool with_alpha) {
            if (t >= 0 && t < T) {
                float* dst = ring + (size_t)(t % (kPF + 1)) * RS;
                const size_t ro = (size_t)t * max_u * K;
                for (int u = tid; u < U; u += nt) {
                    const int o = u * K;
                    if (vec16) {
#pragma unroll
                        for (int k = 0; k < K; k += 4) {
                            cp_async16(dst + o + k, le + ro + o + k);
                            cp_async16(dst + max_u * K + o + k, ls + ro + o + k);
                            if (with_alpha) cp_async16(dst + 2 * max_u * K + o + k, scr + ro + o + k);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            cp_async4(dst + o + k, le + ro + o + k);
                            cp_async4(dst + max_u * K + o + k, ls + ro + o + k);
                            if (with_alpha) cp_async4(dst + 2 * max_u * K + o + k, scr + ro + o + k);
                        }
                    }
                }
            }
            cp_async_commit();
        };
        for (int i = tid; i < RW; i += nt) { cur[i] = kNeg; nxt[i] = kNeg; }
        for (int i = tid; i < max_u + 2; i += nt) srow[i] = kNeg;
        for (int i = tid; i < max_u * K; i += nt) {
            tone2[i] = (i / K) < U ? to_log2(lt[i]) : kNeg;
            gacc[i] = 0.0f;
        }
        __syncthreads();
        for (int k = tid; k < K; k += nt) cur[K + k] = tone2[k];  // alpha(0,0,k) = lt(0,k)
        __syncthreads();
        float off = 0.0f;
        // ------------------------------ forward ------------------------------
        for (int t = 0; t < kPF; ++t) issue_row(t, false);
        for (int t = 0; t < T; ++t) {
            issue_row(t + kPF, false);
            cp_async_wait_pf();  // row t has landed (this thread's part, which is all it reads)
            const float* rle = ring + (size_t)(t % (kPF + 1)) * RS;
            const float* rls = rle + max_u * K;
            if ((t & 15) == 15) {
                float mx = kNeg;
                for (int i = tid; i < U * K; i += nt) mx = fmaxf(mx, cur[K + i]);
                mx = block_max(mx, red);
                const float c = mx > kNegTest ? rintf(mx) : 0.0f;
                for (int i = tid; i < U * K; i += nt) cur[K + i] = fmaxf(cur[K + i] - c, kNeg);
                off += c;
                __syncthreads();
            }
            for (int i = tid; i < U * K; i += nt) scr[(size_t)t * max_u * K + i] = cur[K + i];
            if (tid == 0) soffs[t] = off;
            if (t < T - 1) {
                for (int u = tid; u < U; u += nt) {
                    // LSE over the K tones as max + one LG2 of a sum of EX2s: the K exponentials are
                    // independent (a chain of K pairwise log-add-exps costs K dependent EX2+LG2 pairs)
                    float acc = kNeg;
                    if (u < U - 1) {
                        float mx = kNeg;
#pragma unroll
                        for (int k = 0; k < K; ++k) mx = fmaxf(mx, cur[(u + 1) * K + k] + to_log2(rls[u * K + k]));
                        if (mx > kNegTest) {
                            float sum = 0.0f;
#pragma unroll
                            for (int k = 0; k < K; ++k) sum += ex2((cur[(u + 1) * K + k] + to_log2(rls[u * K + k])) - mx);
                            acc = mx + lg2(sum);
                        }
                    }
                    srow[u + 1] = acc;
                }
                __syncthreads();
                for (int u = tid; u < U; u += nt)
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const float stay = cur[(u + 1) * K + k] + to_log2(rle[u * K + k]);
                        const float sh = u > 0 ? tone2[u * K + k] + srow[u] : kNeg;
                        nxt[(u + 1) * K + k] = lae2(stay, sh);
                    }
                __syncthreads();
                float* tmp = cur; cur = nxt; nxt = tmp;
            }
        }
        cp_async_wait_all();
        __syncthreads();  // also: every stored alpha row is written before the backward sweep prefetches it
        __threadfence_block();
        {
            float acc = kNeg;
#pragma unroll
            for (int k = 0; k < K; ++k)
                acc = lae2(acc, cur[U * K + k] + to_log2(le[((size_t)(T - 1) * max_u + U - 1) * K + k]));
            llt = acc;
        }
        offA_last = off;
        dead = !(llt > kNegTest);
        if (tid == 0) {
            const double ll2 = (double)llt + (double)offA_last;
            a.log_likelihood[b] = dead ? -INFINITY : (float)(ll2 * kLn2);
        }
        __syncthreads();
        // ------------------------------ backward + gradients ------------------------------
        for (int i = tid; i < RW; i += nt) { cur[i] = kNeg; nxt[i] = kNeg; }
        __syncthreads();
        for (int k = tid; k < K; k += nt) cur[U * K + k] = 0.0f;  // virtual beta(T, U-1, k) = 0
        __syncthreads();
        float offB = 0.0f;
        float* bm = srow;                  // Bm(t+1, u), index u+1
        for (int j = 0; j < kPF; ++j) issue_row(T - 1 - j, true);
        for (int t = T - 1; t >= 0; --t) {
            issue_row(t - kPF, true);
            cp_async_wait_pf();
            const float* rle = ring + (size_t)(t % (kPF + 1)) * RS;
            const float* rls = rle + max_u * K;
            const float* ral = rls + max_u * K;
            const float kt = ((soffs[t] - offA_last) + offB) - llt;
            // Bm(t+1,u) = LSE_k(lt(u,k) + beta(t+1,u,k))
            for (int u = tid; u < max_u + 1; u += nt) {
                float acc = kNeg;
                if (u < U) {
                    float mx = kNeg;
#pragma unroll
                    for (int k = 0; k < K; ++k) mx = fmaxf(mx, tone2[u * K + k] + cur[(u + 1) * K + k]);
                    if (mx > kNegTest) {
                        float sum = 0.0f;
#pragma unroll
                        for (int k = 0; k < K; ++k) sum += ex2((tone2[u * K + k] + cur[(u + 1) * K + k]) - mx);
                        acc = mx + lg2(sum);
                    }
                }
                bm[u + 1] = acc;
            }
            __syncthreads();
            for (int u = tid; u < max_u; u += nt) {
                // S(t, u) = LSE_k(alpha(t,u,k) + ls(t,u,k)) of this thread's own token (used by token u+1's
                // tone gradient): maximum first, the sum of exponentials rides along the main loop
                float s_max = kNeg, s_sum = 0.0f;
                if (u < U && !(t == T - 1 || u == U - 1))
#pragma unroll
                    for (int k = 0; k < K; ++k) s_max = fmaxf(s_max, ral[u * K + k] + to_log2(rls[u * K + k]));
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const size_t o = ((size_t)t * max_u + u) * K + k;
                    float g1 = 0.0f, g2 = 0.0f;
                    if (u < U) {
                        const float e = to_log2(rle[u * K + k]);
                        const float s = (t == T - 1 || u == U - 1) ? kNeg : to_log2(rls[u * K + k]);
                        const float av = ral[u * K + k];
                        const float x = e + cur[(u + 1) * K + k];
                        const float y = s + bm[u + 2];
                        if (!dead) {
                            g1 = ex2((av + x) + kt);
                            g2 = ex2((av + y) + kt);
                        }
                        nxt[(u + 1) * K + k] = lae2(x, y);
                        if (s_max > kNegTest) s_sum += ex2((av + s) - s_max);
                    }
                    ge[o] = g1;
                    gs[o] = g2;
                }
                const float s_prev = s_max > kNegTest ? s_max + lg2(s_sum) : kNeg;
                // entering token u+1 at frame t+1 with tone k: S(t,u) + lt(u+1,k) + beta(t+1,u+1,k)
                if (u + 1 < U && !dead)
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        gacc[(u + 1) * K + k] +=
                            ex2(((s_prev + tone2[(u + 1) * K + k]) + cur[(u + 2) * K + k]) + kt);
            }
            __syncthreads();
            float* tmp = cur; cur = nxt; nxt = tmp;
            if ((t & 15) == 0 && t > 0) {
                float mx = kNeg;
                for (int i = tid; i < U * K; i += nt) mx = fmaxf(mx, cur[K + i]);
                mx = block_max(mx, red);
                const float c = mx > kNegTest ? rintf(mx) : 0.0f;
                for (int i = tid; i < U * K; i += nt) cur[K + i] = fmaxf(cur[K + i] - c, kNeg);
                offB += c;
                __syncthreads();
            }
        }
        // token 0 draws its tone at the start: exp(lt(0,k) + beta(0,0,k) - LL)
        const float k0 = (offB - offA_last) - llt;
        for (int i = tid; i < max_u * K; i += nt) {
            float g = gacc[i];
            if (i < K) g = dead ? 0.0f : ex2((tone2[i] + cur[K + i]) + k0);
            gt[i] = (i / K) < U ? g : 0.0f;
        }
        for (size_t i = (size_t)T * max_u * K + tid; i < slab; i += nt) { ge[i] = 0.0f; gs[i] = 0.0f; }
    }
    // deterministic loss: last CTA sums ll in index order
    __threadfence();
    if (tid == 0) s_last = (atomicAdd(p.counter, 1u) == (unsigned)(a.batch_size - 1)) ? 1u : 0u;
    __syncthreads();
    if (s_last && tid < 32) {
        __threadfence();
        double acc = 0.0;
        for (int i = tid; i < a.batch_size; i += 32) acc -= (double)__ldcg(a.log_likelihood + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
        if (tid == 0) {
            if (a.loss) *a.loss = (float)acc;
            *p.counter = 0u;
        }
        if (a.xchg && a.loss) loss_exchange_publish(a.xchg, (float)acc, tid);  // multi-GPU, as in lattice::finish_loss
    }
}

}  // namespace

size_t tone_fb_workspace_bytes(int B, int max_t, int max_u, int K) {
    if (B <= 0 || max_t <= 0 || max_u <= 0 || K <= 0) return 256;
    size_t n = ((size_t)B * max_t * max_u * K + (size_t)B * max_t) * sizeof(float);
    n = (n + 255) & ~(size_t)255;
    const size_t bf = (tone_bf_workspace_bytes(B, max_t, max_u, K) + 255) & ~(size_t)255;
    return n + bf + tone_ws_workspace_bytes(B, max_t, max_u, K);
}

void launch_tone_forward_backward(const ToneFbArgs& a_in, cudaStream_t stream) {
    ToneFbArgs a = a_in;
    if (a.batch_size <= 0) {
        if (a.loss) SSNT_CUDA(cudaMemsetAsync(a.loss, 0, sizeof(float), stream));
        return;
    }
    SSNT_ASSERT(a.tone_class_size > 0, "tone_class_size must be positive");
    SSNT_ASSERT(a.max_t > 0 && a.max_u > 0, "tone_latent_forward_backward: empty lattice");
    void* ws = a.workspace;
    const size_t need = tone_fb_workspace_bytes(a.batch_size, a.max_t, a.max_u, a.tone_class_size);
    if (ws) {
        SSNT_ASSERT(a.workspace_bytes >= need, "tone_latent_forward_backward: workspace too small");
    } else {
        ws = device_scratch(1, need);
    }
    const int K = a.tone_class_size;
    const size_t log_bytes = (((size_t)a.batch_size * a.max_t * a.max_u * K + (size_t)a.batch_size * a.max_t) * sizeof(float) + 255) & ~(size_t)255;
    // A block-float path first; the log-domain kernel then redoes only what it flagged.  Warp-serial kernels (tone_ws.cu)
    // once there are at least two utterances per SM, and for every shape the split-role kernel (tone_bf.cu: K = 4,
    // max_u in {32,64,128}, a cluster of four CTAs per utterance) does not take.
    const unsigned* only = nullptr;
    const bool aligned = (reinterpret_cast<uintptr_t>(ws) & 15u) == 0;
    const bool can_bf = tone_bf_supported(a) && aligned, can_ws = tone_ws_supported(a) && aligned;
    int kind = tone_forced_kernel_kind();
    if (kind < 0) {
        if (can_ws && (!can_bf || (size_t)a.batch_size >= (size_t)2 * sm_count())) kind = 2;
        else kind = can_bf ? 1 : 0;
    }
    if ((kind == 1 && !can_bf) || (kind >= 2 && !can_ws)) kind = 0;
    tone_note_kernel_kind(kind);
    const size_t bf_bytes = (tone_bf_workspace_bytes(a.batch_size, a.max_t, a.max_u, K) + 255) & ~(size_t)255;
    if (kind == 1) only = launch_tone_bf(a, (char*)ws + log_bytes, done_counter_for(ws), stream);
    else if (kind >= 2) only = launch_tone_ws(a, (char*)ws + log_bytes + bf_bytes, kind == 3, stream);
    a.xchg = loss_exchange_device();  // only this kernel's reduction is exchanged (the block-float kernel's is provisional)
    ToneParams p;
    p.a = a;
    p.scratch = (float*)ws;
    p.offs = (float*)ws + (size_t)a.batch_size * a.max_t * a.max_u * a.tone_class_size;
    p.counter = done_counter_for(ws);  // (the block-float kernel above hands it back zeroed before this one starts)
    p.only = only;
    p.fallbacks = device_fallback_counter();
    int threads = ((a.max_u + 31) / 32) * 32;
    threads = threads > 1024 ? 1024 : threads;
    const size_t smem = ((size_t)2 * (a.max_u + 2) * K + (a.max_u + 2) + (size_t)2 * a.max_u * K +
                         (size_t)(kPF + 1) * 3 * a.max_u * K + a.max_t) * sizeof(float) + 16;
    SSNT_ASSERT(smem <= 227 * 1024, "tone_latent_forward_backward: max_u * K too large for shared memory");
    auto launch = [&](auto kernel) {
        if (smem > 48 * 1024)
            SSNT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<a.batch_size, threads, smem, stream>>>(p);
        SSNT_CUDA(cudaGetLastError());
    };
    switch (K) {
        case 1: launch(tone_fb_kernel<1>); break;
        case 2: launch(tone_fb_kernel<2>); break;
        case 4: launch(tone_fb_kernel<4>); break;
        case 8: launch(tone_fb_kernel<8>); break;
        default: launch(tone_fb_kernel<0>); break;
    }
}

}  // namespace ssnt
