// Runtime plumbing behind the C-ABI: stream side channel, device error flag, pointer-space detection,
// grow-only scratch, completion counters and the NVLink loss exchange.  The reference's functions are
// stateless and re-entrant; all state here is thread-local, per device, or immutable after
// first use so that the same holds for this library.
#include <atomic>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "ssnt_common.cuh"

namespace ssnt {

namespace {
constexpr int kMaxDevices = 64;
constexpr int kSlots = 40;

thread_local cudaStream_t tls_stream = nullptr;
thread_local int tls_space = kAuto;

struct Scratch {
    void* ptr = nullptr;
    size_t cap = 0;
};
// per host thread AND per device: a thread that switches devices gets separate blocks
struct ThreadScratch {
    Scratch dev[kSlots];
    Scratch pin[kSlots];
};
thread_local ThreadScratch* tls_scratch[kMaxDevices] = {};

// State of one device, created on first use with that device current.
struct DeviceState {
    std::once_flag once;
    unsigned* flag_host = nullptr;  // mapped pinned words: [0] error bits, [1] utterances re-run in the log domain
    unsigned* flag_dev = nullptr;
    int sm_count = 0;
    // completion counters ("last CTA reduces the loss"): one 32-byte sector per workspace address, zero whenever no
    // kernel using it is in flight (the last CTA hands it back zeroed).  Keyed by workspace so that a captured graph,
    // which replays with its workspace, can never share a counter with an eager call on another buffer.
    std::mutex mu;
    std::unordered_map<const void*, unsigned*> counters;
    std::vector<unsigned*> blocks;
    size_t used_in_block = 0;
    static constexpr size_t kPerBlock = 4096;
    LossExchange* xchg_dev = nullptr;  // device copy of the exchange descriptor (null until connected)
    LossExchange xchg_host{};
    unsigned long long* xchg_local = nullptr;
};
DeviceState g_dev[kMaxDevices];

int current_device() {
    int dev = 0;
    SSNT_CUDA(cudaGetDevice(&dev));
    SSNT_ASSERT(dev >= 0 && dev < kMaxDevices, "device ordinal out of range");
    return dev;
}

DeviceState& state() {
    const int dev = current_device();
    DeviceState& s = g_dev[dev];
    std::call_once(s.once, [&]() {
        SSNT_CUDA(cudaHostAlloc((void**)&s.flag_host, 4 * sizeof(unsigned), cudaHostAllocMapped | cudaHostAllocPortable));
        s.flag_host[0] = s.flag_host[1] = s.flag_host[2] = s.flag_host[3] = 0;
        SSNT_CUDA(cudaHostGetDevicePointer((void**)&s.flag_dev, s.flag_host, 0));
        SSNT_CUDA(cudaDeviceGetAttribute(&s.sm_count, cudaDevAttrMultiProcessorCount, dev));
    });
    return s;
}

ThreadScratch& scratch() {
    const int dev = current_device();
    if (!tls_scratch[dev]) tls_scratch[dev] = new ThreadScratch();
    return *tls_scratch[dev];
}
}  // namespace

cudaStream_t current_stream() { return tls_stream; }
void set_stream(cudaStream_t s) { tls_stream = s; }
void set_space(int s) { tls_space = s; }

unsigned* device_error_flag() { return state().flag_dev; }
int sm_count() { return state().sm_count; }
int device_ordinal() { return current_device(); }
unsigned* device_fallback_counter() { return state().flag_dev + 1; }
unsigned read_fallback_counter() { return ((volatile unsigned*)state().flag_host)[1]; }

unsigned* done_counter_for(const void* workspace) {
    DeviceState& s = state();
    std::lock_guard<std::mutex> lock(s.mu);
    auto it = s.counters.find(workspace);
    if (it != s.counters.end()) return it->second;
    if (s.blocks.empty() || s.used_in_block == DeviceState::kPerBlock) {
        unsigned* blk = nullptr;
        SSNT_CUDA(cudaMalloc((void**)&blk, DeviceState::kPerBlock * 32));
        SSNT_CUDA(cudaMemset(blk, 0, DeviceState::kPerBlock * 32));
        s.blocks.push_back(blk);
        s.used_in_block = 0;
    }
    unsigned* c = s.blocks.back() + (s.used_in_block++) * 8;  // one 32-byte sector each
    s.counters.emplace(workspace, c);
    return c;
}

unsigned read_and_clear_error_flag() {
    DeviceState& s = state();
    unsigned v = *(volatile unsigned*)s.flag_host;
    if (v) *(volatile unsigned*)s.flag_host = 0;
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
    if (v & kErrLossExchange)
        panic("loss exchange: a peer rank did not deliver its loss (did every rank make the same calls?)", __FILE__, __LINE__);
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
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) {
        // one process may drive several GPUs, but a call's buffers must live on the current device
        SSNT_ASSERT(at.device == current_device(),
                    "device pointer belongs to another GPU than the current one (cudaSetDevice before the call)");
        return true;
    }
    return false;
}

bool is_pinned_host_pointer(const void* p) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

void* device_scratch(int slot, size_t bytes) {
    SSNT_ASSERT(slot >= 0 && slot < kSlots, "scratch slot");
    Scratch& s = scratch().dev[slot];
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
    Scratch& s = scratch().pin[slot];
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

// ---- loss exchange over NVLink peer memory ---------------------------------------------------------------------
// Every rank owns a slot buffer [kLossRing][world] of 8-byte entries {loss bits, sequence number}.  The kernel that
// reduces a call's loss stores its entry into EVERY rank's buffer (its own included) with one 64-bit store per peer;
// the sum of a call's `world` entries is the all-reduced loss.  Nothing is launched for it and the host issues no
// collective: the exchange is replayed with the CUDA graph that holds the call.
void loss_exchange_export(int world, unsigned char handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    SSNT_ASSERT(world >= 1 && world <= kLossMaxWorld, "loss exchange: world size out of range");
    DeviceState& s = state();
    if (!s.xchg_local) {
        const size_t bytes = (size_t)kLossRing * kLossMaxWorld * sizeof(unsigned long long);
        SSNT_CUDA(cudaMalloc((void**)&s.xchg_local, bytes));
        SSNT_CUDA(cudaMemset(s.xchg_local, 0, bytes));
        SSNT_CUDA(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    SSNT_CUDA(cudaIpcGetMemHandle(&h, s.xchg_local));
    std::memcpy(handle_out, &h, 64);
}

void loss_exchange_connect(int rank, int world, const unsigned char* handles) {
    SSNT_ASSERT(world >= 1 && world <= kLossMaxWorld && rank >= 0 && rank < world, "loss exchange: bad rank / world size");
    DeviceState& s = state();
    SSNT_ASSERT(s.xchg_local != nullptr, "loss exchange: call ssnt_tts_loss_exchange_export first");
    LossExchange x{};
    x.rank = rank;
    x.world = world;
    x.seq = 0;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            x.peers[r] = s.xchg_local;
        } else {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handles + (size_t)r * 64, 64);
            void* p = nullptr;
            SSNT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            x.peers[r] = (unsigned long long*)p;
        }
    }
    if (!s.xchg_dev) SSNT_CUDA(cudaMalloc((void**)&s.xchg_dev, sizeof(LossExchange)));
    SSNT_CUDA(cudaMemcpy(s.xchg_dev, &x, sizeof(x), cudaMemcpyHostToDevice));
    s.xchg_host = x;
}

void loss_exchange_disconnect() {
    DeviceState& s = state();
    if (!s.xchg_dev) return;
    SSNT_CUDA(cudaDeviceSynchronize());
    for (int r = 0; r < s.xchg_host.world; ++r)
        if (r != s.xchg_host.rank && s.xchg_host.peers[r]) cudaIpcCloseMemHandle(s.xchg_host.peers[r]);
    SSNT_CUDA(cudaFree(s.xchg_dev));
    s.xchg_dev = nullptr;
    s.xchg_host = LossExchange{};
}

LossExchange* loss_exchange_device() { return state().xchg_dev; }

}  // namespace ssnt
