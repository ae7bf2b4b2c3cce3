// Batched unit-cost Levenshtein distance, int32, bit-exact with src/edit_distance.rs:6-60
// (Kaldi's two-row recurrence: e[n] = min(e'[n-1] + delta, e'[n] + 1, e[n-1] + 1)).
//
// One CTA per (a, b) pair, sliced to its true lengths first as the reference does
// (src/edit_distance.rs:19-20).  The M x N table is swept as an anti-diagonal wavefront:
// thread n owns column n, on diagonal d it fills cell (m = d - n, n) from the two previous
// diagonals, which rotate through three shared-memory rows — M + N dependent steps instead of
// M * N.  Both sequences are staged in shared memory.  Integer min/add only, so the result is
// the same number the serial recurrence produces.
#include "ssnt_common.cuh"

namespace ssnt {
namespace {

struct EditParams {
    const int* a;
    const int* b;
    const int* a_len;
    const int* b_len;
    int B, max_length;
    int* distance;
};

__global__ void edit_distance_kernel(const EditParams p) {
    extern __shared__ int sm[];
    const int pair = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int L = p.max_length;
    int M = p.a_len[pair], N = p.b_len[pair];
    M = min(max(M, 0), L);
    N = min(max(N, 0), L);
    int* sa = sm;               // [L]
    int* sb = sa + L;           // [L]
    int* diag = sb + L;         // 3 x [L + 1], indexed by column n
    const int* ga = p.a + (size_t)pair * L;
    const int* gb = p.b + (size_t)pair * L;
    for (int i = tid; i < M; i += nt) sa[i] = ga[i];
    for (int i = tid; i < N; i += nt) sb[i] = gb[i];
    __syncthreads();
    const int stride = L + 1;
    // diagonal d holds cells (m, n) with m + n = d; column n ranges max(0, d-M) .. min(N, d)
    for (int d = 0; d <= M + N; ++d) {
        int* cur = diag + (d % 3) * stride;
        const int* p1 = diag + ((d + 2) % 3) * stride;  // diagonal d-1
        const int* p2 = diag + ((d + 1) % 3) * stride;  // diagonal d-2
        const int nlo = max(0, d - M), nhi = min(N, d);
        for (int n = nlo + tid; n <= nhi; n += nt) {
            const int m = d - n;
            int v;
            if (n == 0) v = m;        // e_tmp[0] = e[0] + 1 accumulated over m rows
            else if (m == 0) v = n;   // e = (0..=N)
            else {
                const int term1 = p2[n - 1] + (sa[m - 1] == sb[n - 1] ? 0 : 1);
                const int term2 = p1[n] + 1;       // E(m-1, n) + 1
                const int term3 = p1[n - 1] + 1;   // E(m, n-1) + 1
                v = min(term1, min(term2, term3));
            }
            cur[n] = v;
        }
        __syncthreads();
    }
    if (tid == 0) p.distance[pair] = diag[((M + N) % 3) * stride + N];
}

}  // namespace

void levenshtein_edit_distance(const int* a, const int* b, const int* a_len, const int* b_len,
                               int batch_size, int max_length, int* distance, cudaStream_t stream) {
    if (batch_size <= 0) return;
    EditParams p{a, b, a_len, b_len, batch_size, max_length, distance};
    int threads = ((max_length + 1 + 31) / 32) * 32;
    threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
    const size_t smem = ((size_t)2 * max_length + 3 * (size_t)(max_length + 1)) * sizeof(int);
    SSNT_ASSERT(smem <= 227 * 1024, "edit distance: max_length too large for shared memory");
    if (smem > 48 * 1024)
        SSNT_CUDA(cudaFuncSetAttribute(edit_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edit_distance_kernel<<<batch_size, threads, smem, stream>>>(p);
    SSNT_CUDA(cudaGetLastError());
}

}  // namespace ssnt
