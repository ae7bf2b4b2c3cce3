// Batched unit-cost Levenshtein distance, int32, bit-exact with src/edit_distance.rs:6-60
// (Kaldi's two-row recurrence: e[n] = min(e'[n-1] + delta, e'[n] + 1, e[n-1] + 1)).
// One CTA per (a, b) pair, sliced to its true lengths first as the reference does (src/edit_distance.rs:19-20).
//
// edit_distance_bitpar_kernel (max_length <= 1024): the bit-vector form of the same table (Myers 1999 / Hyyro 2003).
// Column n of the table is held as its vertical differences E(m, n) - E(m-1, n) in {-1, 0, +1}: two bit masks (plus,
// minus) of M bits, cut into 32-row blocks, one block per lane of ONE warp.  A column step of a block is ~20 integer
// instructions on (Pv, Mv), the match mask Eq(m) = [a[m] == b[n]] and the horizontal difference entering from the block
// above (+1 at the top edge: E(0, n) = n); it returns the horizontal difference leaving its last row.  Blocks are
// pipelined along the lanes with a skew of TWO columns (lane k works on column s - 2k at step s), so the shuffle that
// carries a block's outgoing difference to the lane below has a whole step to land and the step's dependency chain is
// the block's own (Pv, Mv) update only: N + 2(M/32 - 1) steps instead of the M + N barrier-separated anti-diagonals of
// the wavefront kernel.  All match masks are computed first by every warp of the CTA (each lane keeps its block's 32
// symbols of `a` in registers; 32 compares per mask) into shared memory, [N][M/32] words.  E(M, N) = M + the sum of the
// last block's outgoing differences.  Integer-exact, so the result is the number the serial recurrence produces.
//
// edit_distance_kernel (any longer max_length): anti-diagonal wavefront: thread n owns column n, on diagonal d it fills
// cell (m = d - n, n) from the two previous diagonals, which rotate through three shared-memory rows.
#include <cstdlib>
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


// ---- bit-parallel kernel ---------------------------------------------------------------------------------
constexpr int kBitparMaxLen = 1024;  // 32 lanes x 32 rows

__global__ void __launch_bounds__(1024) edit_distance_bitpar_kernel(const EditParams p) {
    extern __shared__ int sm[];
    const int pair = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int L = p.max_length;
    int M = p.a_len[pair], N = p.b_len[pair];
    M = min(max(M, 0), L);
    N = min(max(N, 0), L);
    if (M == 0 || N == 0) {  // uniform over the CTA
        if (tid == 0) p.distance[pair] = M + N;
        return;
    }
    const int W = (M + 31) >> 5;                       // blocks of 32 rows
    int* sa = sm;                                      // [33 * 32]  a, padded: symbol i at i + i / 32 (conflict-free column reads)
    int* sb = sa + 33 * 32;                            // [L]
    unsigned* eq = reinterpret_cast<unsigned*>(sb + L);  // [N][W] match masks
    const int* ga = p.a + (size_t)pair * L;
    const int* gb = p.b + (size_t)pair * L;
    for (int i = tid; i < M; i += nt) sa[i + (i >> 5)] = ga[i];
    for (int i = tid; i < N; i += nt) sb[i] = gb[i];
    __syncthreads();
    if (lane < W) {
        // rows >= M of the last block compare against stale shared memory: their mask bits are garbage, and no bit of a
        // block's update ever depends on a higher one (carries and shifts only move upwards), so rows < M are unaffected
        int pa[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) pa[i] = sa[lane * 33 + i];
        for (int n = warp; n < N; n += nwarps) {
            const int c = sb[n];
            unsigned w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                w0 |= (pa[i + 0] == c ? 1u : 0u) << (i + 0);
                w1 |= (pa[i + 1] == c ? 1u : 0u) << (i + 1);
                w2 |= (pa[i + 2] == c ? 1u : 0u) << (i + 2);
                w3 |= (pa[i + 3] == c ? 1u : 0u) << (i + 3);
            }
            eq[n * W + lane] = (w0 | w1) | (w2 | w3);
        }
    }
    __syncthreads();
    if (warp != 0) return;

    const bool active = lane < W;
    const int top = (lane == W - 1) ? ((M - 1) & 31) : 31;  // the block's last real row
    unsigned Pv = 0xffffffffu, Mv = 0u;                     // column 0: E(m, 0) = m, every vertical difference is +1
    int score = 0;
    int q0 = 0, q1 = 0;                                     // differences received one and two steps ago
    const int steps = N + 2 * (W - 1);
    int n = -2 * lane;                                      // this lane's column at step s
    unsigned Eq = (active && n == 0) ? eq[lane] : 0u;
#pragma unroll 2
    for (int s = 0; s < steps; ++s, ++n) {
        const bool valid = active && n >= 0 && n < N;
        const bool nvalid = active && n + 1 >= 0 && n + 1 < N;
        const unsigned Eq_next = nvalid ? eq[(n + 1) * W + lane] : 0u;  // consumed one step later
        const int hin = lane == 0 ? 1 : q0;
        const unsigned hneg = hin < 0 ? 1u : 0u, hpos = hin > 0 ? 1u : 0u;
        const unsigned Xv = Eq | Mv;
        const unsigned E2 = Eq | hneg;
        const unsigned Xh = (((E2 & Pv) + Pv) ^ Pv) | E2;
        unsigned Ph = Mv | ~(Xh | Pv);
        unsigned Mh = Pv & Xh;
        const int hout = (int)((Ph >> top) & 1u) - (int)((Mh >> top) & 1u);
        Ph = (Ph << 1) | hpos;
        Mh = (Mh << 1) | hneg;
        if (valid) {
            Pv = Mh | ~(Xv | Ph);
            Mv = Ph & Xv;
            score += hout;
        }
        q0 = q1;
        q1 = __shfl_up_sync(0xffffffffu, hout, 1);  // used by the lane below two steps from now, on the same column
        Eq = Eq_next;
    }
    if (lane == W - 1) p.distance[pair] = M + score;
}

}  // namespace

void levenshtein_edit_distance(const int* a, const int* b, const int* a_len, const int* b_len,
                               int batch_size, int max_length, int* distance, cudaStream_t stream) {
    if (batch_size <= 0) return;
    EditParams p{a, b, a_len, b_len, batch_size, max_length, distance};
    static const int force_wavefront = [] { const char* e = std::getenv("SSNT_EDIT_WAVEFRONT"); return e ? std::atoi(e) : 0; }();  // A/B aid
    if (max_length <= kBitparMaxLen && !force_wavefront) {
        const int words = (max_length + 31) / 32;
        const size_t smem = ((size_t)33 * 32 + (size_t)max_length + (size_t)words * max_length) * sizeof(int);
        const int threads = max_length > 512 ? 1024 : (max_length > 128 ? 512 : 128);
        if (smem > 48 * 1024)
            SSNT_CUDA(cudaFuncSetAttribute(edit_distance_bitpar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        edit_distance_bitpar_kernel<<<batch_size, threads, smem, stream>>>(p);
        SSNT_CUDA(cudaGetLastError());
        return;
    }
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
