// Shared declarations of the sm_100a backend behind the ssnt_tts_c C-ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace ssnt {

// ---- error handling ---------------------------------------------------------------------
// The reference reports every precondition failure by panicking (assert!/assert_eq! in
// ssnt_tts_c/src/lib.rs and src/*.rs), i.e. the process aborts.  We keep that behaviour.
[[noreturn]] inline void panic(const char* what, const char* file, int line) {
    std::fprintf(stderr, "ssnt_tts_c: panic: %s (%s:%d)\n", what, file, line);
    std::fflush(stderr);
    std::abort();
}
#define SSNT_ASSERT(cond, msg)                                   \
    do {                                                         \
        if (!(cond)) ::ssnt::panic(msg, __FILE__, __LINE__);      \
    } while (0)
#define SSNT_CUDA(call)                                                            \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) ::ssnt::panic(cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Device-side conditions that are a panic in the reference are recorded as bits of a flag word
// (mapped pinned host memory) and turned into the panic at the next synchronising call.
enum ErrorBits : unsigned {
    kErrV2EmptyBeam = 1u,        // src/v2.rs:292 assert_ne!(n_results, 0)
    kErrUpsampleLength = 2u,     // src/v2_util.rs:58 assert_eq!(upsampled.len(), output_length)
    kErrToneEmptyBeam = 4u,      // src/tone_latent.rs:199 `i % n_results` with n_results == 0
    kErrBadIndex = 8u,           // out-of-range parent index in a back-trace table
    kErrLossExchange = 16u,      // a peer rank never delivered its loss entry (multi-GPU loss exchange)
};

// ---- runtime (runtime.cu) ------------------------------------------------------------------
cudaStream_t current_stream();             // thread-local side channel, see ssnt_tts_set_stream
void set_stream(cudaStream_t s);
void set_space(int space);
// Completion counter of the kernels that run on `workspace`: zero whenever none of them is in flight (the last
// CTA hands it back zeroed).  One per workspace address, so a replayed graph never shares one with another call.
unsigned* done_counter_for(const void* workspace);
int device_ordinal();                      // cudaGetDevice(), checked
unsigned* device_error_flag();             // device pointer to the flag word
unsigned read_and_clear_error_flag();      // host side; caller must have synchronised
unsigned* device_fallback_counter();       // utterances the block-float kernel re-ran in the log domain
unsigned read_fallback_counter();          // host side, cumulative
void check_error_flag_or_panic();          // panics with the reference's message if a bit is set

enum MemSpace { kAuto = 0, kHost = 1, kDevice = 2 };
bool is_device_pointer(const void* p);     // honours ssnt_tts_set_memory_space; panics on a pointer of another GPU
bool is_pinned_host_pointer(const void* p);  // page-locked host memory (cudaHostAlloc / cudaHostRegister)?

// ---- loss exchange over NVLink peer memory (runtime.cu, lattice_common.cuh::finish_loss) ---------------------------
constexpr int kLossMaxWorld = 16;  // ranks of one node
constexpr int kLossRing = 1024;    // calls whose entries are kept (ranks may drift apart by that many calls)
struct LossExchange {
    int rank, world;
    unsigned seq;                                 // calls completed so far (device side; advances under graph replay)
    unsigned long long* peers[kLossMaxWorld];     // every rank's slot buffer [kLossRing][kLossMaxWorld] of {loss bits, seq}
};
#ifdef __CUDACC__
// Called by the 32 lanes of the warp that has just reduced a call's loss (every lane holds `loss`): lane r posts
// {loss bits, call number} into rank r's slot buffer with one plain 64-bit store over NVLink.  Posted stores: the
// warp does not wait for a round trip; the reader polls for the call number (loss_allreduce_kernel).
__device__ __forceinline__ void loss_exchange_publish(LossExchange* x, float loss, int lane) {
    unsigned seq = 0;
    if (lane == 0) seq = ++x->seq;
    seq = __shfl_sync(0xffffffffu, seq, 0);
    if (lane < x->world) {
        const unsigned long long entry = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(loss);
        unsigned long long* dst = x->peers[lane] + (size_t)(seq % (unsigned)kLossRing) * kLossMaxWorld + (size_t)x->rank;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(entry) : "memory");
    }
}
#endif
void loss_exchange_export(int world, unsigned char handle_out[64]);
void loss_exchange_connect(int rank, int world, const unsigned char* handles);
void loss_exchange_disconnect();
LossExchange* loss_exchange_device();      // null until connected
void launch_loss_allreduce(float* out_device, cudaStream_t stream);

// Host-to-host copies spread over the library's copy threads for the duration of one call (host_copy.cu).
class HostCopier {
public:
    HostCopier();
    ~HostCopier();
    void copy(void* dst, const void* src, size_t bytes);
private:
    bool pooled_;
};
int host_copy_threads();

// Grow-only per-thread device scratch (workspace the caller did not provide, staging of host
// buffers).  Slots are independent so one call can hold several live buffers.
void* device_scratch(int slot, size_t bytes);
void* pinned_scratch(int slot, size_t bytes);
int sm_count();

// RAII staging of one host array on the device for the host-pointer flavour of the C-ABI.
struct Staged {
    void* dev = nullptr;
    void* host = nullptr;
    size_t bytes = 0;
    bool copy_back = false;
};

// ---- launchers (one per kernel family; all asynchronous on current_stream()) -----------------
struct FbArgs {
    const float* log_emit;
    const float* log_shift;
    const int* t_len;   // may be null → max_t
    const int* u_len;   // may be null → max_u
    int batch_size, max_t, max_u;
    float* log_likelihood;  // [B]
    float* loss;            // [1], may be null
    float* grad_emit;       // [B,T,U]
    float* grad_shift;      // [B,T,U]
    void* workspace;
    size_t workspace_bytes;
    struct LossExchange* xchg = nullptr;  // set by the launcher when a loss exchange is connected (multi-GPU)
    // Raw-logit mode (ssnt_tts_forward_backward_logits): log_emit = log sigmoid(z), log_shift = log sigmoid(-z) are
    // formed on the fly and the gradient is chained through them, dLL/dz = grad_emit * sigmoid(-z) - grad_shift * sigmoid(z).
    // When `logits` is set, log_emit / log_shift / grad_emit / grad_shift are null.
    const float* logits = nullptr;   // [B,T,U]
    float* grad_logits = nullptr;    // [B,T,U]
};
size_t fb_workspace_bytes(int batch_size, int max_t, int max_u);
size_t fb_logits_workspace_bytes(int batch_size, int max_t, int max_u);  // raw-logit mode (adds the unfused path's buffers where needed)
void launch_forward_backward(const FbArgs& a, cudaStream_t stream);
// Which kernel family the last launch_forward_backward on this thread used (1 = warp/TMA
// lattice kernel, 0 = generic block kernel); for tests and the bench's launch accounting.
int fb_last_kernel_kind();
void fb_force_kernel_kind(int kind);  // -1 auto, 0 generic, 1 log-warp, 2 block-float, 3 = 2 + forced re-run
void fb_set_stats_buffer(long long* dev);
long long* fb_get_stats_buffer();  // profiling aid for the block-float kernel (null = off)

struct ToneFbArgs {
    const float* log_emit;   // [B,T,U,K]
    const float* log_shift;  // [B,T,U,K]
    const float* log_tone;   // [B,U,K]
    const int* t_len;
    const int* u_len;
    int batch_size, max_t, max_u, tone_class_size;
    float* log_likelihood;
    float* loss;
    float* grad_emit;
    float* grad_shift;
    float* grad_tone;
    void* workspace;
    size_t workspace_bytes;
    struct LossExchange* xchg = nullptr;
};
size_t tone_fb_workspace_bytes(int batch_size, int max_t, int max_u, int tone_class_size);
// block-float split-role tone kernel (tone_bf.cu): supported shapes, extra workspace, launch (returns the
// device [B] status words: non-zero = re-run that utterance in the log domain)
bool tone_bf_supported(const ToneFbArgs& a);
size_t tone_bf_workspace_bytes(int batch_size, int max_t, int max_u, int tone_class_size);
unsigned* launch_tone_bf(const ToneFbArgs& a, void* workspace, unsigned* counter, cudaStream_t stream);
// warp-serial block-float tone kernels (tone_ws.cu): the large-batch path and every K in {2,4,8} / max_u in {32..256}
bool tone_ws_supported(const ToneFbArgs& a);
size_t tone_ws_workspace_bytes(int batch_size, int max_t, int max_u, int tone_class_size);
unsigned* launch_tone_ws(const ToneFbArgs& a, void* workspace, int force_fallback, cudaStream_t stream);
// tone kernel selection for tests/benchmarks: -1 auto, 0 log-domain only, 1 split-role block-float, 2 warp-serial
// block-float, 3 = 2 with every utterance re-run in the log domain
void tone_force_kernel_kind(int kind);
int tone_forced_kernel_kind();
int tone_last_kernel_kind();
void tone_note_kernel_kind(int kind);
void launch_tone_forward_backward(const ToneFbArgs& a, cudaStream_t stream);

}  // namespace ssnt
