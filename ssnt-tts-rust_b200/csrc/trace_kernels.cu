// Back-trace and duration upsampling.
//
//  * backtrace_kernel: src/util.rs:20-33 (one final branch + t_history gather) and
//    src/v2_util.rs:6-36 (every final branch of every batch entry).  The reference walks the
//    (T, W) parent table serially from the last row, a T-long pointer chase.  Here the table is
//    staged in shared memory and the chase is cut into C chunks: (1) every (chunk, w) composes
//    its chunk's parent maps, (2) each final branch hops over the C chunk maps, (3) every
//    (chunk, final) re-walks its chunk from the now-known entry point and writes its rows —
//    dependent depth ~2·T/C + C instead of T.
//  * upsample_kernel: src/v2_util.rs:39-66, an inclusive scan of the durations followed by a
//    binary search per output slot (coalesced writes; slots >= output_length stay untouched,
//    as the caller pre-filled them, upsample_source_indexes_op.cc:75).
#include "ssnt_common.cuh"

namespace ssnt {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kThreads = 256;

struct TraceParams {
    const int* final_branch;  // [B, F]; null → final_scalar (the reference passes one i32 by value)
    int final_scalar;
    const int* table;         // [B, T, W]
    const int* t_history;     // [B, T, W] or null
    int B, F, T, W;
    int* out_branch;          // [B, F, T]
    int* out_t;               // [B, F, T] or null
    unsigned* err;
    int staged;               // table fits shared memory
};

__global__ void __launch_bounds__(kThreads) backtrace_kernel(const TraceParams p) {
    extern __shared__ int sm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int T = p.T, W = p.W, F = p.F;
    const int* gtable = p.table + (size_t)b * T * W;
    // chunking: C chunks of L rows
    int C = kThreads / (W > 0 ? W : 1);
    C = C < 1 ? 1 : (C > T ? T : C);
    const int L = (T + C - 1) / C;
    C = (T + L - 1) / L;
    int* maps = sm;                 // [C][W]   entry (row hi) → exit (below row lo) of a chunk
    int* entry = maps + C * W;      // [C][F]   branch at the top row of chunk c for final f
    int* stab = entry + C * F;      // [T][W]   staged table
    const int* table = gtable;
    if (p.staged) {
        for (int i = tid; i < T * W; i += kThreads) stab[i] = gtable[i];
        table = stab;
    }
    __syncthreads();
    bool bad = false;
    // (1) chunk maps: rows [lo, hi] walked downwards from hi
    for (int idx = tid; idx < C * W; idx += kThreads) {
        const int c = idx / W, w = idx - c * W;
        const int lo = c * L, hi = min(T, lo + L) - 1;
        int cur = w;
        for (int r = hi; r >= lo; --r) {
            cur = table[r * W + cur];
            if ((unsigned)cur >= (unsigned)W) { bad = true; cur = 0; }
        }
        maps[idx] = cur;
    }
    __syncthreads();
    // (2) hop over chunks from the last one
    for (int f = tid; f < F; f += kThreads) {
        int cur = p.final_branch ? p.final_branch[(size_t)b * F + f] : p.final_scalar;
        if ((unsigned)cur >= (unsigned)W) { bad = true; cur = 0; }
        for (int c = C - 1; c >= 0; --c) {
            entry[c * F + f] = cur;
            cur = maps[c * W + cur];
        }
    }
    __syncthreads();
    // (3) re-walk and write
    for (int idx = tid; idx < C * F; idx += kThreads) {
        const int c = idx / F, f = idx - c * F;
        const int lo = c * L, hi = min(T, lo + L) - 1;
        int cur = entry[idx];
        int* ob = p.out_branch + ((size_t)b * F + f) * T;
        int* ot = p.out_t ? p.out_t + ((size_t)b * F + f) * T : nullptr;
        const int* th = p.t_history ? p.t_history + (size_t)b * T * W : nullptr;
        for (int r = hi; r >= lo; --r) {
            ob[r] = cur;
            if (ot) ot[r] = th[r * W + cur];
            cur = table[r * W + cur];
            if ((unsigned)cur >= (unsigned)W) cur = 0;
        }
    }
    if (bad) atomicOr(p.err, kErrBadIndex);
}

struct UpsampleParams {
    const int* duration;       // [R, T]   R = B*W rows
    const int* output_length;  // [R]
    int R, T, max_u;
    int* out;                  // [R, max_u]
    unsigned* err;
};

__global__ void __launch_bounds__(kThreads) upsample_kernel(const UpsampleParams p) {
    extern __shared__ int prefix[];  // inclusive scan of the durations
    __shared__ int warp_tot[kThreads / 32];
    __shared__ int s_base, s_bad;
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int T = p.T;
    const int* d = p.duration + (size_t)row * T;
    if (tid == 0) { s_base = 0; s_bad = 0; }
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += kThreads) {
        const int t = t0 + tid;
        int v = t < T ? d[t] : 0;
        if (v < 0) { s_bad = 1; v = 0; }  // `*d as usize` of a negative number panics in the reference
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(kFull, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        int add = s_base;
        for (int w = 0; w < wid; ++w) add += warp_tot[w];
        if (t < T) prefix[t] = x + add;
        __syncthreads();
        if (tid == kThreads - 1) s_base = x + add;
        __syncthreads();
    }
    const int total = s_base;
    const int len = p.output_length[row];
    if (s_bad || total != len) {  // assert_eq!(upsampled.len(), output_length[0]) src/v2_util.rs:58
        if (tid == 0) atomicOr(p.err, kErrUpsampleLength);
        return;
    }
    const int lim = min(len, p.max_u);
    int* o = p.out + (size_t)row * p.max_u;
    for (int pos = tid; pos < lim; pos += kThreads) {
        // smallest t with prefix[t] > pos
        int lo = 0, hi = T - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (prefix[mid] > pos) hi = mid; else lo = mid + 1;
        }
        o[pos] = lo;
    }
}

void launch_backtrace(TraceParams p, cudaStream_t stream) {
    if (p.B <= 0 || p.T <= 0 || p.F <= 0) return;
    SSNT_ASSERT(p.W > 0, "beam_width must be positive");
    int C = kThreads / p.W;
    C = C < 1 ? 1 : (C > p.T ? p.T : C);
    size_t base = ((size_t)C * p.W + (size_t)C * p.F) * sizeof(int);
    size_t staged = base + (size_t)p.T * p.W * sizeof(int);
    p.staged = staged <= 200 * 1024;
    size_t smem = p.staged ? staged : base;
    SSNT_ASSERT(smem <= 227 * 1024, "back-trace: beam too wide for shared memory");
    if (smem > 48 * 1024)
        SSNT_CUDA(cudaFuncSetAttribute(backtrace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    backtrace_kernel<<<p.B, kThreads, smem, stream>>>(p);
    SSNT_CUDA(cudaGetLastError());
}

}  // namespace

void extract_best_beam_branch(int best_final_branch, const int* beam_branch,
                              const int* t_history, int beam_width, int max_u,
                              int* best_beam_branch, int* best_t_history, cudaStream_t stream) {
    TraceParams p{};
    p.final_branch = nullptr;
    p.final_scalar = best_final_branch;
    p.table = beam_branch;
    p.t_history = t_history;
    p.B = 1; p.F = 1; p.T = max_u; p.W = beam_width;
    p.out_branch = best_beam_branch;
    p.out_t = best_t_history;
    p.err = device_error_flag();
    launch_backtrace(p, stream);
}

void order_beam_branch(const int* final_branch, const int* beam_branch, int batch_size,
                       int beam_width, int max_t, int* ordered, cudaStream_t stream) {
    TraceParams p{};
    p.final_branch = final_branch;
    p.table = beam_branch;
    p.t_history = nullptr;
    p.B = batch_size; p.F = beam_width; p.T = max_t; p.W = beam_width;
    p.out_branch = ordered;
    p.out_t = nullptr;
    p.err = device_error_flag();
    launch_backtrace(p, stream);
}

void upsample_source_indexes(const int* duration, const int* output_length, int batch_size,
                             int beam_width, int max_t, int max_u, int* out, cudaStream_t stream) {
    const long long rows = (long long)batch_size * beam_width;
    if (rows <= 0) return;
    UpsampleParams p{};
    p.duration = duration; p.output_length = output_length;
    p.R = (int)rows; p.T = max_t; p.max_u = max_u; p.out = out; p.err = device_error_flag();
    const size_t smem = (size_t)(max_t > 0 ? max_t : 1) * sizeof(int);
    SSNT_ASSERT(smem <= 227 * 1024, "upsample: max_t too large for shared memory");
    if (smem > 48 * 1024)
        SSNT_CUDA(cudaFuncSetAttribute(upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upsample_kernel<<<(unsigned)rows, kThreads, smem, stream>>>(p);
    SSNT_CUDA(cudaGetLastError());
}

// Device-side pre-fill for the DEVICE_GPU op wrappers (the reference's ops pre-fill their outputs on the host:
// ssnt_tts_beam_search_decode_op.cc:91, ssnt_tts_v2_beam_search_decode_op.cc:212, upsample_source_indexes_op.cc:75).
namespace {
__global__ void fill_i32_kernel(int* dst, size_t n, int value) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = value;
}
}  // namespace
void device_fill_i32(int* dst, size_t n, int value, cudaStream_t stream) {
    if (n == 0) return;
    const unsigned blocks = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    fill_i32_kernel<<<blocks, 256, 0, stream>>>(dst, n, value);
    SSNT_CUDA(cudaGetLastError());
}

}  // namespace ssnt
