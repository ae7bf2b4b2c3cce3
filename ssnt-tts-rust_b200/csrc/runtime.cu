// Runtime plumbing behind the C-ABI: stream side channel, device error flag, pointer-space
// detection and grow-only scratch.  The reference's functions are stateless and re-entrant
// (SURVEY.md §8b); all state here is thread-local or immutable after first use so that the
// same holds for this library.
#include <atomic>
#include <mutex>
#include <vector>

#include "ssnt_common.cuh"

namespace ssnt {

namespace {
thread_local cudaStream_t tls_stream = nullptr;
thread_local int tls_space = kAuto;

struct Scratch {
    void* ptr = nullptr;
    size_t cap = 0;
};
constexpr int kSlots = 40;
thread_local Scratch tls_dev[kSlots];
thread_local Scratch tls_pin[kSlots];

std::once_flag g_flag_once;
unsigned* g_flag_host = nullptr;  // mapped pinned word
unsigned* g_flag_dev = nullptr;  // [0] error bits, [1] count of utterances re-run in the log domain
int g_sm_count = 0;
constexpr int kCounters = 1024;
unsigned* g_counters = nullptr;  // device, zeroed once; each user resets its ticket to 0
std::atomic<unsigned> g_next_counter{0};

void init_flag() {
    SSNT_CUDA(cudaHostAlloc((void**)&g_flag_host, 4 * sizeof(unsigned), cudaHostAllocMapped));
    g_flag_host[0] = g_flag_host[1] = g_flag_host[2] = g_flag_host[3] = 0;
    SSNT_CUDA(cudaHostGetDevicePointer((void**)&g_flag_dev, g_flag_host, 0));
    int dev = 0;
    SSNT_CUDA(cudaGetDevice(&dev));
    SSNT_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    SSNT_CUDA(cudaMalloc((void**)&g_counters, kCounters * 32));
    SSNT_CUDA(cudaMemset(g_counters, 0, kCounters * 32));
}
}  // namespace

cudaStream_t current_stream() { return tls_stream; }
void set_stream(cudaStream_t s) { tls_stream = s; }
void set_space(int s) { tls_space = s; }

unsigned* device_error_flag() {
    std::call_once(g_flag_once, init_flag);
    return g_flag_dev;
}

unsigned* next_done_counter() {
    std::call_once(g_flag_once, init_flag);
    unsigned i = g_next_counter.fetch_add(1) % kCounters;
    return g_counters + (size_t)i * 8;  // one 32-byte sector per ticket
}

int sm_count() {
    std::call_once(g_flag_once, init_flag);
    return g_sm_count;
}

unsigned* device_fallback_counter() {
    std::call_once(g_flag_once, init_flag);
    return g_flag_dev + 1;
}

unsigned read_fallback_counter() {
    std::call_once(g_flag_once, init_flag);
    return ((volatile unsigned*)g_flag_host)[1];
}

unsigned read_and_clear_error_flag() {
    std::call_once(g_flag_once, init_flag);
    unsigned v = *(volatile unsigned*)g_flag_host;
    if (v) *(volatile unsigned*)g_flag_host = 0;
    return v;
}

void check_error_flag_or_panic() {
    unsigned v = read_and_clear_error_flag();
    if (!v) return;
    if (v & kErrV2EmptyBeam)
        panic("Beam search could not find a duration sequence with compatible output length. "
              "Please increase duration class size and beam width. (src/v2.rs:292)",
              __FILE__, __LINE__);
    if (v & kErrUpsampleLength)
        panic("upsample_source_indexes: sum(duration) != output_length (src/v2_util.rs:58)",
              __FILE__, __LINE__);
    if (v & kErrToneEmptyBeam)
        panic("tone_latent beam search: empty candidate set (src/tone_latent.rs:199)", __FILE__, __LINE__);
    panic("back-trace: parent index out of range (slice index panic)", __FILE__, __LINE__);
}

bool is_device_pointer(const void* p) {
    if (tls_space == kHost) return false;
    if (tls_space == kDevice) return true;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

void* device_scratch(int slot, size_t bytes) {
    SSNT_ASSERT(slot >= 0 && slot < kSlots, "scratch slot");
    Scratch& s = tls_dev[slot];
    if (bytes > s.cap) {
        if (s.ptr) {
            // Work that still uses the old block may be in flight on the current stream.
            SSNT_CUDA(cudaStreamSynchronize(tls_stream));
            SSNT_CUDA(cudaFree(s.ptr));
        }
        size_t cap = bytes + bytes / 4 + 256;
        SSNT_CUDA(cudaMalloc(&s.ptr, cap));
        s.cap = cap;
    }
    return s.ptr;
}

void* pinned_scratch(int slot, size_t bytes) {
    SSNT_ASSERT(slot >= 0 && slot < kSlots, "scratch slot");
    Scratch& s = tls_pin[slot];
    if (bytes > s.cap) {
        if (s.ptr) {
            SSNT_CUDA(cudaStreamSynchronize(tls_stream));
            SSNT_CUDA(cudaFreeHost(s.ptr));
        }
        size_t cap = bytes + bytes / 4 + 256;
        SSNT_CUDA(cudaHostAlloc(&s.ptr, cap, cudaHostAllocDefault));
        s.cap = cap;
    }
    return s.ptr;
}

}  // namespace ssnt
