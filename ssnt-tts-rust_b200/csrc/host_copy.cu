// Parallel host-to-host copies for the host-pointer lattice calls (c_api.cu).
//
// The reference's DEVICE_CPU callers hold ordinary pageable buffers
// (ssnt-tts-tensorflow/src/ssnt_tts_v2_beam_search_decode_op.cc:146-177 allocate_output).  A cudaMemcpyAsync on
// pageable memory is staged by the driver on one thread and overlaps with nothing, so the library stages such
// buffers itself: chunk by chunk through its own page-locked scratch, with the host-side copy spread over a small
// pool of threads while the DMA engines and the kernels work on the neighbouring chunks.
//
// The pool is per process and lazily created.  Between calls its workers sleep on a condition variable; during a
// call (begin() .. end()) they spin on one atomic word, so handing them a copy costs about a microsecond.  Pieces
// are claimed dynamically, the calling thread takes pieces too, and a second host thread that finds the pool busy
// simply copies by itself: correctness never depends on a worker being scheduled.
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <immintrin.h>

#include "ssnt_common.cuh"

namespace ssnt {

namespace {

constexpr size_t kPiece = (size_t)256 << 10;  // bytes per claimed piece

// One piece.  The destination is written once and next read by a DMA engine (staging in) or by the caller much
// later (staging out), so the stores bypass the cache: no read-for-ownership of the destination lines, a third less
// memory traffic than memcpy below glibc's own non-temporal threshold.  SSNT_COPY_NT=0 selects plain memcpy.
bool use_streaming_stores() {
    static const bool nt = [] { const char* e = std::getenv("SSNT_COPY_NT"); return !e || std::atoi(e) != 0; }();
    return nt;
}
void copy_piece(char* dst, const char* src, size_t n) {
    if (!use_streaming_stores() || n < 4096) {
        std::memcpy(dst, src, n);
        return;
    }
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (head) {
        std::memcpy(dst, src, head);
        dst += head; src += head; n -= head;
    }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i));
        const __m128i b = _mm_loadu_si128((const __m128i*)(src + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32));
        const __m128i d = _mm_loadu_si128((const __m128i*)(src + i + 48));
        _mm_stream_si128((__m128i*)(dst + i), a);
        _mm_stream_si128((__m128i*)(dst + i + 16), b);
        _mm_stream_si128((__m128i*)(dst + i + 32), c);
        _mm_stream_si128((__m128i*)(dst + i + 48), d);
    }
    _mm_sfence();
    if (i < n) std::memcpy(dst + i, src + i, n - i);
}

class CopyPool {
public:
    static CopyPool& get() {
        static CopyPool* p = new CopyPool();  // leaked at exit: workers must not be joined from a static destructor
        return *p;
    }

    bool try_begin() {
        if (workers_.empty() || !busy_.try_lock()) return false;
        {
            std::lock_guard<std::mutex> l(mu_);
            active_.store(true, std::memory_order_release);
        }
        cv_.notify_all();
        return true;
    }
    void end() {
        active_.store(false, std::memory_order_release);
        busy_.unlock();
    }

    // Only between try_begin() and end().
    void copy(void* dst, const void* src, size_t bytes) {
        const unsigned npieces = (unsigned)((bytes + kPiece - 1) / kPiece);
        if (npieces <= 1) {
            copy_piece((char*)dst, (const char*)src, bytes);
            return;
        }
        unsigned long long e = ((next_.load(std::memory_order_relaxed) >> 32) + 1) & 0xffffffffull;
        if (e == 0) e = 2;  // epoch 0 means "no job yet"; keep the parity alternating
        Job& j = jobs_[e & 1];
        j.dst = (char*)dst;
        j.src = (const char*)src;
        j.bytes = bytes;
        j.npieces.store(npieces, std::memory_order_relaxed);
        done_.store(0, std::memory_order_relaxed);
        next_.store(e << 32, std::memory_order_release);  // publishes the job
        work(e);
        while (done_.load(std::memory_order_acquire) != npieces) _mm_pause();
    }

    int threads() const { return (int)workers_.size() + 1; }

private:
    struct Job {
        char* dst = nullptr;
        const char* src = nullptr;
        size_t bytes = 0;
        std::atomic<unsigned> npieces{0};
    };

    CopyPool() {
        int n = 0;
        if (const char* e = std::getenv("SSNT_COPY_THREADS")) {
            n = std::atoi(e) - 1;  // the calling thread counts
        } else {
            const unsigned hw = std::thread::hardware_concurrency();
            n = (int)(hw / 2) - 1;  // measured at cfg2 on a 16-core host: 1 thread 4.8 ms, 2: 2.9, 4: 2.0, 8: 1.45 per call
            n = n > 11 ? 11 : (n < 0 ? 0 : n);
        }
        for (int k = 0; k < n; ++k) workers_.emplace_back([this] { loop(); });
        for (auto& t : workers_) t.detach();
    }

    // Claims and copies pieces of job `e` until none is left (or, e == 0: of whatever job is current).
    bool work(unsigned long long e) {
        bool did = false;
        for (;;) {
            unsigned long long cur = next_.load(std::memory_order_acquire);
            const unsigned long long ce = cur >> 32;
            if (ce == 0 || (e != 0 && ce != e)) return did;
            Job& j = jobs_[ce & 1];
            const unsigned idx = (unsigned)cur, np = j.npieces.load(std::memory_order_relaxed);
            if (idx >= np) return did;
            if (!next_.compare_exchange_weak(cur, cur + 1, std::memory_order_acq_rel)) continue;
            // the claim succeeded on epoch ce: its descriptor is stable until every piece is done
            const size_t off = (size_t)idx * kPiece;
            const size_t n = j.bytes - off < kPiece ? j.bytes - off : kPiece;
            copy_piece(j.dst + off, j.src + off, n);
            done_.fetch_add(1, std::memory_order_release);
            did = true;
        }
    }

    void loop() {
        for (;;) {
            if (!active_.load(std::memory_order_acquire)) {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [this] { return active_.load(std::memory_order_acquire); });
            }
            unsigned idle = 0;
            while (active_.load(std::memory_order_acquire)) {
                if (work(0)) {
                    idle = 0;
                } else if (++idle < 4096) {
                    _mm_pause();
                } else {
                    std::this_thread::yield();
                }
            }
        }
    }

    std::vector<std::thread> workers_;
    std::mutex busy_;  // one call at a time owns the workers
    std::mutex mu_;
    std::condition_variable cv_;
    std::atomic<bool> active_{false};
    Job jobs_[2];
    std::atomic<unsigned long long> next_{0};  // (epoch << 32) | next piece to claim
    std::atomic<unsigned> done_{0};
};

}  // namespace

HostCopier::HostCopier() : pooled_(CopyPool::get().try_begin()) {}
HostCopier::~HostCopier() {
    if (pooled_) CopyPool::get().end();
}
void HostCopier::copy(void* dst, const void* src, size_t bytes) {
    if (!bytes) return;
    if (pooled_) CopyPool::get().copy(dst, src, bytes);
    else copy_piece((char*)dst, (const char*)src, bytes);
}
int host_copy_threads() { return CopyPool::get().threads(); }

}  // namespace ssnt
