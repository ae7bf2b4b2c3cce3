// extern "C" surface (include/ssnt_tts_c.h).  Mirrors ssnt_tts_c/src/lib.rs: every entry point
// null-checks its pointers (assert!(!p.is_null()) there → abort here), derives the array lengths
// from the scalar arguments exactly as the `from_raw_parts` calls do, and forwards to the compute
// layer.  The one addition is the memory-space dispatch: device pointers are passed straight to
// the kernels (asynchronous on the caller's stream); host pointers — what the reference's
// DEVICE_CPU TensorFlow ops hand over — are staged to the GPU, the same kernels run, results are
// copied back and the call returns once the host buffers are complete.  There is no CPU compute
// path in this library.
#include <vector>

#include "../../include/ssnt_tts_c.h"
#include "ssnt_common.cuh"

namespace ssnt {
// compute layer (device pointers), defined in the kernel translation units
void v1_beam_search_decode(const float*, const float*, const bool*, const int*, const int*, int, int, int,
                           int*, float*, int*, int*, bool*, int*, cudaStream_t);
void v2_beam_search_decode(const float*, const float*, const bool*, const int*, const int*, const int*,
                           const int*, const int*, const int*, int, int, int, int, bool, bool, int*,
                           float*, int*, int*, bool*, int*, int*, cudaStream_t);
void tone_beam_search_decode(const float*, const float*, const bool*, const int*, const int*, const int*,
                             int, int, int, int, int*, float*, int*, int*, bool*, int*, cudaStream_t);
void extract_best_beam_branch(int, const int*, const int*, int, int, int*, int*, cudaStream_t);
void order_beam_branch(const int*, const int*, int, int, int, int*, cudaStream_t);
void upsample_source_indexes(const int*, const int*, int, int, int, int, int*, cudaStream_t);
void levenshtein_edit_distance(const int*, const int*, const int*, const int*, int, int, int*, cudaStream_t);
void device_fill_i32(int*, size_t, int, cudaStream_t);
void v2_decode_loop(const float*, const int*, const int*, const int*, const float*, const bool*, const int*, const int*,
                    const int*, int, int, int, int, int, bool, bool, int, int*, int*, float*, int*, int*, bool*, int*, int*,
                    int*, int*, cudaStream_t);
void tone_decode_loop(const float*, const int*, const float*, const bool*, const int*, const int*, int, int, int, int, int,
                      int*, int*, float*, int*, int*, bool*, int*, int*, cudaStream_t);

namespace {

#define NOT_NULL(p) SSNT_ASSERT((p) != nullptr, "assertion failed: !" #p ".is_null()")

inline size_t n3(int a, int b, int c) { return (size_t)(a > 0 ? a : 0) * (b > 0 ? b : 0) * (c > 0 ? c : 0); }
inline size_t n2(int a, int b) { return (size_t)(a > 0 ? a : 0) * (b > 0 ? b : 0); }

// Host-pointer flavour: per-call staging through the thread's grow-only device scratch.
class HostCall {
public:
    HostCall() : stream_(current_stream()) {}
    template <typename T>
    const T* in(const T* host, size_t n) {
        if (!host) return nullptr;  // optional argument
        T* d = (T*)device_scratch(slot_++, n * sizeof(T) + 16);
        if (n) SSNT_CUDA(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, stream_));
        return d;
    }
    // preload = the callee leaves some slots untouched, so the caller's pre-fill must survive
    template <typename T>
    T* out(T* host, size_t n, bool preload = false) {
        if (!host) return nullptr;  // optional argument
        T* d = (T*)device_scratch(slot_++, n * sizeof(T) + 16);
        if (preload && n) SSNT_CUDA(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, stream_));
        outs_.push_back({d, host, n * sizeof(T)});
        return d;
    }
    cudaStream_t stream() const { return stream_; }
    // Copies results back; aborts (like the reference's panic) if a device-side assert fired, in
    // which case the host outputs are left untouched.
    void finish() {
        SSNT_CUDA(cudaStreamSynchronize(stream_));
        check_error_flag_or_panic();
        for (auto& o : outs_)
            if (o.bytes) SSNT_CUDA(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, stream_));
        SSNT_CUDA(cudaStreamSynchronize(stream_));
    }

private:
    struct Out { void* dev; void* host; size_t bytes; };
    cudaStream_t stream_;
    int slot_ = 4;  // slots 0..3 are workspaces
    std::vector<Out> outs_;
};

// ---- host-pointer lattice calls: chunked staging pipeline -----------------------------------------------------
// The batch is cut into chunks; each chunk runs H2D, kernels, D2H in order on its own stream, so chunk c+1 uploads
// while chunk c computes and chunk c-1 downloads.  Host buffers of the reference's DEVICE_CPU callers are ordinary
// pageable allocations (ssnt-tts-tensorflow/src/ssnt_tts_v2_beam_search_decode_op.cc:146-177 allocate_output): those
// are staged through the thread's page-locked scratch, the host-side copies spread over the library's copy threads
// (host_copy.cu) and overlapped with the DMA of the neighbouring chunks.  Buffers the caller page-locked itself are
// handed to the DMA engines directly.  A call larger than kStageBytes is processed in several passes so that neither
// the page-locked nor the device scratch grows with the batch.
constexpr int kMaxChunks = 16;
constexpr int kComputeStreams = 8;
constexpr size_t kStageBytes = (size_t)1 << 30;  // host bytes (inputs + outputs) of one pass
struct AuxStreams {
    cudaStream_t k[kComputeStreams] = {};
    cudaEvent_t start = nullptr, done[kMaxChunks] = {};
    void init() {
        if (start) return;
        for (int i = 0; i < kComputeStreams; ++i) SSNT_CUDA(cudaStreamCreateWithFlags(&k[i], cudaStreamNonBlocking));
        for (int i = 0; i < kMaxChunks; ++i) SSNT_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        SSNT_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    }
};
thread_local AuxStreams tls_aux[64];  // per device (a host thread may drive several GPUs)

// One host array of a lattice call, cut along the batch axis.
struct HostArray {
    const void* host_in;  // inputs
    void* host_out;       // outputs
    size_t per_b;         // bytes per utterance
    int slot;             // scratch slot (device and page-locked)
    char* dev = nullptr;
    char* stage = nullptr;  // page-locked staging copy, or the caller's own buffer if that is page-locked
    bool direct = false;
};

// launch(b0, nb, chunk, stream): enqueues the kernels of utterances [b0, b0 + nb) (device arrays via .dev).
template <class Launch>
void staged_lattice_pass(std::vector<HostArray>& ins, std::vector<HostArray>& outs, int B, int nchunks, Launch&& launch) {
    HostCopier copier;
    for (auto* v : {&ins, &outs})
        for (HostArray& a : *v) {
            const void* user = a.host_in ? a.host_in : a.host_out;
            a.dev = (char*)device_scratch(a.slot, a.per_b * B + 16);
            a.direct = is_pinned_host_pointer(user);
            a.stage = a.direct ? (char*)const_cast<void*>(user) : (char*)pinned_scratch(a.slot, a.per_b * B + 16);
        }
    AuxStreams& aux = tls_aux[device_ordinal()];
    aux.init();
    SSNT_CUDA(cudaEventRecord(aux.start, current_stream()));
    const int per = (B + nchunks - 1) / nchunks;
    int issued = 0, drained = 0;
    auto drain = [&](int c) {  // chunk c has landed in the staging buffers: hand it to the caller
        const int b0 = c * per, nb = (b0 + per <= B ? per : B - b0);
        for (HostArray& a : outs)
            if (!a.direct) copier.copy((char*)a.host_out + (size_t)b0 * a.per_b, a.stage + (size_t)b0 * a.per_b, (size_t)nb * a.per_b);
    };
    for (int c = 0; c < nchunks; ++c) {
        const int b0 = c * per, nb = (b0 + per <= B ? per : B - b0);
        if (nb <= 0) break;
        cudaStream_t s = aux.k[c % kComputeStreams];
        if (c < kComputeStreams) SSNT_CUDA(cudaStreamWaitEvent(s, aux.start, 0));
        for (HostArray& a : ins) {
            const size_t o = (size_t)b0 * a.per_b, n = (size_t)nb * a.per_b;
            if (!a.direct) copier.copy(a.stage + o, (const char*)a.host_in + o, n);
            SSNT_CUDA(cudaMemcpyAsync(a.dev + o, a.stage + o, n, cudaMemcpyHostToDevice, s));
        }
        launch(b0, nb, c, s);
        for (HostArray& a : outs) {
            const size_t o = (size_t)b0 * a.per_b, n = (size_t)nb * a.per_b;
            SSNT_CUDA(cudaMemcpyAsync(a.stage + o, a.dev + o, n, cudaMemcpyDeviceToHost, s));
        }
        SSNT_CUDA(cudaEventRecord(aux.done[c], s));
        issued = c + 1;
        // whatever has completed meanwhile goes out now, while the later chunks are in flight
        while (drained < issued - 1 && cudaEventQuery(aux.done[drained]) == cudaSuccess) drain(drained++);
    }
    for (; drained < issued; ++drained) {
        SSNT_CUDA(cudaEventSynchronize(aux.done[drained]));
        drain(drained);
    }
    cudaGetLastError();  // cudaEventQuery's cudaErrorNotReady is not an error
}

int lattice_chunks(int B, size_t in_bytes) {
    // measured at cfg2 (26 MB in, 26 MB out) — see DESIGN.md §4.4; small calls are not worth cutting
    int n = 1;
    // cfg2, pageable, 12 copy threads: 2 chunks 1.24 ms, 3: 1.12, 4: 1.11, 5: 1.07-1.18, 6: 1.13-1.21, 8: 1.20-1.26, 10: 1.19-1.25
    if (B >= 2 && in_bytes >= ((size_t)2 << 20)) n = (int)(in_bytes / ((size_t)6 << 20));
    n = n < 2 ? (B >= 2 && in_bytes >= ((size_t)2 << 20) ? 2 : 1) : n;
    n = n > 8 ? 8 : n;
    static const int env_chunks = [] { const char* e = std::getenv("SSNT_FB_CHUNKS"); return e ? std::atoi(e) : 0; }();  // tuning aid
    if (env_chunks > 0) n = env_chunks > kMaxChunks ? kMaxChunks : env_chunks;
    if (n > B) n = B > 0 ? B : 1;
    return n;
}

}  // namespace
}  // namespace ssnt

using namespace ssnt;

#pragma GCC visibility push(default)
extern "C" {

// ssnt_tts_c/src/lib.rs:10-83 — "Restricted to single batch."
void ssnt_tts_beam_search_decode(const float* h, const float* log_prob_history, const bool* is_finished,
                                 const int* t, const int* u, int max_t, int beam_width, int* prediction,
                                 float* log_probs, int* next_t, int* next_u, bool* next_is_finished,
                                 int* beam_branch) {
    NOT_NULL(h); NOT_NULL(log_prob_history); NOT_NULL(is_finished); NOT_NULL(t); NOT_NULL(u);
    NOT_NULL(prediction); NOT_NULL(log_probs); NOT_NULL(next_t); NOT_NULL(next_u);
    NOT_NULL(next_is_finished); NOT_NULL(beam_branch);
    const int batch_size = 1;
    const size_t W = n2(batch_size, beam_width);
    if (is_device_pointer(h)) {
        v1_beam_search_decode(h, log_prob_history, is_finished, t, u, batch_size, max_t, beam_width,
                              prediction, log_probs, next_t, next_u, next_is_finished, beam_branch,
                              current_stream());
        return;
    }
    HostCall c;
    auto dh = c.in(h, W * 2); auto dl = c.in(log_prob_history, W); auto df = c.in(is_finished, W);
    auto dt = c.in(t, W); auto du = c.in(u, W);
    auto op = c.out(prediction, W); auto ol = c.out(log_probs, W); auto ot = c.out(next_t, W);
    auto ou = c.out(next_u, W); auto of = c.out(next_is_finished, W); auto ob = c.out(beam_branch, W);
    v1_beam_search_decode(dh, dl, df, dt, du, batch_size, max_t, beam_width, op, ol, ot, ou, of, ob, c.stream());
    c.finish();
}

// ssnt_tts_c/src/lib.rs:86-116
void ssnt_extract_best_beam_branch(int best_final_branch, const int* beam_branch, const int* t_history,
                                   int beam_width, int max_u, int* best_beam_branch, int* best_t_history) {
    NOT_NULL(beam_branch); NOT_NULL(t_history); NOT_NULL(best_beam_branch); NOT_NULL(best_t_history);
    const size_t n = n2(max_u, beam_width), m = (size_t)(max_u > 0 ? max_u : 0);
    if (is_device_pointer(beam_branch)) {
        extract_best_beam_branch(best_final_branch, beam_branch, t_history, beam_width, max_u,
                                 best_beam_branch, best_t_history, current_stream());
        return;
    }
    HostCall c;
    auto db = c.in(beam_branch, n); auto dt = c.in(t_history, n);
    auto ob = c.out(best_beam_branch, m); auto ot = c.out(best_t_history, m);
    extract_best_beam_branch(best_final_branch, db, dt, beam_width, max_u, ob, ot, c.stream());
    c.finish();
}

// ssnt_tts_c/src/lib.rs:118-218
void ssnt_tts_v2_beam_search_decode(const float* h, const float* log_prob_history, const bool* is_finished,
                                    const int* total_duration, const int* duration_table, const int* t,
                                    const int* u, const int* input_length, const int* output_length,
                                    int batch_size, int beam_width, int duration_class_size,
                                    int zero_duration_id, bool allow_skip, bool test_mode, int* prediction,
                                    float* log_probs, int* next_t, int* next_u, bool* next_is_finished,
                                    int* next_total_duration, int* beam_branch) {
    NOT_NULL(h); NOT_NULL(log_prob_history); NOT_NULL(is_finished); NOT_NULL(total_duration);
    NOT_NULL(duration_table); NOT_NULL(t); NOT_NULL(u); NOT_NULL(input_length); NOT_NULL(output_length);
    NOT_NULL(prediction); NOT_NULL(log_probs); NOT_NULL(next_t); NOT_NULL(next_u);
    NOT_NULL(next_is_finished); NOT_NULL(next_total_duration); NOT_NULL(beam_branch);
    const size_t BW = n2(batch_size, beam_width), B = (size_t)(batch_size > 0 ? batch_size : 0);
    if (is_device_pointer(h)) {
        v2_beam_search_decode(h, log_prob_history, is_finished, total_duration, duration_table, t, u,
                              input_length, output_length, batch_size, beam_width, duration_class_size,
                              zero_duration_id, allow_skip, test_mode, prediction, log_probs, next_t, next_u,
                              next_is_finished, next_total_duration, beam_branch, current_stream());
        return;
    }
    HostCall c;
    auto dh = c.in(h, n3(batch_size, beam_width, duration_class_size));
    auto dl = c.in(log_prob_history, BW); auto df = c.in(is_finished, BW);
    auto dtd = c.in(total_duration, BW);
    auto dtab = c.in(duration_table, (size_t)(duration_class_size > 0 ? duration_class_size : 0));
    auto dt = c.in(t, BW); auto du = c.in(u, BW);
    auto dil = c.in(input_length, B); auto dol = c.in(output_length, B);
    auto op = c.out(prediction, BW); auto ol = c.out(log_probs, BW); auto ot = c.out(next_t, BW);
    auto ou = c.out(next_u, BW); auto of = c.out(next_is_finished, BW);
    auto otd = c.out(next_total_duration, BW); auto ob = c.out(beam_branch, BW);
    v2_beam_search_decode(dh, dl, df, dtd, dtab, dt, du, dil, dol, batch_size, beam_width,
                          duration_class_size, zero_duration_id, allow_skip, test_mode, op, ol, ot, ou, of,
                          otd, ob, c.stream());
    c.finish();
}

// ssnt_tts_c/src/lib.rs:220-241
void ssnt_order_beam_branch(const int* final_branch, const int* beam_branch, int batch_size, int beam_width,
                            int max_t, int* ordered_beam_branch) {
    NOT_NULL(final_branch); NOT_NULL(beam_branch); NOT_NULL(ordered_beam_branch);
    if (is_device_pointer(final_branch)) {
        order_beam_branch(final_branch, beam_branch, batch_size, beam_width, max_t, ordered_beam_branch,
                          current_stream());
        return;
    }
    HostCall c;
    auto df = c.in(final_branch, n2(batch_size, beam_width));
    auto db = c.in(beam_branch, n3(batch_size, max_t, beam_width));
    auto oo = c.out(ordered_beam_branch, n3(batch_size, beam_width, max_t));
    order_beam_branch(df, db, batch_size, beam_width, max_t, oo, c.stream());
    c.finish();
}

// ssnt_tts_c/src/lib.rs:244-265
void ssnt_upsample_source_indexes(const int* duration, const int* output_length, int batch_size,
                                  int beam_width, int max_t, int max_u, int* upsampled_source_indexes) {
    NOT_NULL(duration); NOT_NULL(output_length); NOT_NULL(upsampled_source_indexes);
    if (is_device_pointer(duration)) {
        upsample_source_indexes(duration, output_length, batch_size, beam_width, max_t, max_u,
                                upsampled_source_indexes, current_stream());
        return;
    }
    HostCall c;
    auto dd = c.in(duration, n3(batch_size, beam_width, max_t));
    auto dl = c.in(output_length, n2(batch_size, beam_width));
    auto oo = c.out(upsampled_source_indexes, n3(batch_size, beam_width, max_u), /*preload=*/true);
    upsample_source_indexes(dd, dl, batch_size, beam_width, max_t, max_u, oo, c.stream());
    c.finish();
}

// ssnt_tts_c/src/lib.rs:267-343
void tone_latent_beam_search_decode(const float* h, const float* log_prob_history, const bool* is_finished,
                                    const int* t, const int* u, const int* input_length, int batch_size,
                                    int beam_width, int tone_class_size, int empty_tone_id, int* prediction,
                                    float* log_probs, int* next_t, int* next_u, bool* next_is_finished,
                                    int* beam_branch) {
    NOT_NULL(h); NOT_NULL(log_prob_history); NOT_NULL(is_finished); NOT_NULL(t); NOT_NULL(u);
    NOT_NULL(input_length); NOT_NULL(prediction); NOT_NULL(log_probs); NOT_NULL(next_t); NOT_NULL(next_u);
    NOT_NULL(next_is_finished); NOT_NULL(beam_branch);
    const size_t BW = n2(batch_size, beam_width);
    if (is_device_pointer(h)) {
        tone_beam_search_decode(h, log_prob_history, is_finished, t, u, input_length, batch_size, beam_width,
                                tone_class_size, empty_tone_id, prediction, log_probs, next_t, next_u,
                                next_is_finished, beam_branch, current_stream());
        return;
    }
    HostCall c;
    auto dh = c.in(h, n3(batch_size, beam_width, tone_class_size));
    auto dl = c.in(log_prob_history, BW); auto df = c.in(is_finished, BW);
    auto dt = c.in(t, BW); auto du = c.in(u, BW);
    auto dil = c.in(input_length, (size_t)(batch_size > 0 ? batch_size : 0));
    auto op = c.out(prediction, BW); auto ol = c.out(log_probs, BW); auto ot = c.out(next_t, BW);
    auto ou = c.out(next_u, BW); auto of = c.out(next_is_finished, BW); auto ob = c.out(beam_branch, BW);
    tone_beam_search_decode(dh, dl, df, dt, du, dil, batch_size, beam_width, tone_class_size, empty_tone_id,
                            op, ol, ot, ou, of, ob, c.stream());
    c.finish();
}

// ssnt_tts_c/src/lib.rs:346-381
void tone_latent_levenshtein_edit_distance(const int* a, const int* b, const int* a_lengths,
                                           const int* b_lengths, int batch_size, int max_length,
                                           int* distance) {
    NOT_NULL(a); NOT_NULL(b); NOT_NULL(a_lengths); NOT_NULL(b_lengths); NOT_NULL(distance);
    const size_t B = (size_t)(batch_size > 0 ? batch_size : 0);
    if (is_device_pointer(a)) {
        levenshtein_edit_distance(a, b, a_lengths, b_lengths, batch_size, max_length, distance, current_stream());
        return;
    }
    HostCall c;
    auto da = c.in(a, n2(batch_size, max_length)); auto db = c.in(b, n2(batch_size, max_length));
    auto dal = c.in(a_lengths, B); auto dbl = c.in(b_lengths, B);
    auto od = c.out(distance, B);
    levenshtein_edit_distance(da, db, dal, dbl, batch_size, max_length, od, c.stream());
    c.finish();
}

// ---- whole-loop decoding (no counterpart in the reference's ABI: SURVEY.md §8 f2) ------------------------------
void ssnt_tts_v2_decode_loop(const float* h, const int* duration_table, const int* input_length, const int* output_length,
                             const float* log_prob_history, const bool* is_finished, const int* total_duration,
                             const int* t, const int* u, int batch_size, int steps, int beam_width,
                             int duration_class_size, int zero_duration_id, bool allow_skip, bool test_mode, int max_u,
                             int* prediction_history, int* beam_branch_history, float* log_probs, int* final_t,
                             int* final_u, bool* final_is_finished, int* final_total_duration, int* ordered_beam_branch,
                             int* duration, int* upsampled_source_indexes) {
    NOT_NULL(h); NOT_NULL(duration_table); NOT_NULL(input_length); NOT_NULL(output_length);
    NOT_NULL(prediction_history); NOT_NULL(beam_branch_history); NOT_NULL(log_probs); NOT_NULL(final_t); NOT_NULL(final_u);
    NOT_NULL(final_is_finished); NOT_NULL(final_total_duration); NOT_NULL(ordered_beam_branch); NOT_NULL(duration);
    const size_t BW = n2(batch_size, beam_width), B = (size_t)(batch_size > 0 ? batch_size : 0);
    const size_t BSW = BW * (size_t)(steps > 0 ? steps : 0);
    if (is_device_pointer(h)) {
        v2_decode_loop(h, duration_table, input_length, output_length, log_prob_history, is_finished, total_duration, t, u,
                       batch_size, steps, beam_width, duration_class_size, zero_duration_id, allow_skip, test_mode, max_u,
                       prediction_history, beam_branch_history, log_probs, final_t, final_u, final_is_finished,
                       final_total_duration, ordered_beam_branch, duration, upsampled_source_indexes, current_stream());
        return;
    }
    HostCall c;
    auto dh = c.in(h, BSW * (size_t)(duration_class_size > 0 ? duration_class_size : 0));
    auto dtab = c.in(duration_table, (size_t)(duration_class_size > 0 ? duration_class_size : 0));
    auto dil = c.in(input_length, B); auto dol = c.in(output_length, B);
    auto dl = c.in(log_prob_history, BW); auto df = c.in(is_finished, BW); auto dtd = c.in(total_duration, BW);
    auto dt = c.in(t, BW); auto du = c.in(u, BW);
    auto oph = c.out(prediction_history, BSW); auto obh = c.out(beam_branch_history, BSW);
    auto ol = c.out(log_probs, BW); auto ot = c.out(final_t, BW); auto ou = c.out(final_u, BW);
    auto of = c.out(final_is_finished, BW); auto otd = c.out(final_total_duration, BW);
    auto oo = c.out(ordered_beam_branch, BSW); auto od = c.out(duration, BSW);
    auto oup = c.out(upsampled_source_indexes, BW * (size_t)(max_u > 0 ? max_u : 0), /*preload=*/true);
    v2_decode_loop(dh, dtab, dil, dol, dl, df, dtd, dt, du, batch_size, steps, beam_width, duration_class_size,
                   zero_duration_id, allow_skip, test_mode, max_u, oph, obh, ol, ot, ou, of, otd, oo, od, oup, c.stream());
    c.finish();
}

void tone_latent_decode_loop(const float* h, const int* input_length, const float* log_prob_history, const bool* is_finished,
                             const int* t, const int* u, int batch_size, int steps, int beam_width, int tone_class_size,
                             int empty_tone_id, int* prediction_history, int* beam_branch_history, float* log_probs,
                             int* final_t, int* final_u, bool* final_is_finished, int* ordered_beam_branch,
                             int* ordered_tone) {
    NOT_NULL(h); NOT_NULL(input_length); NOT_NULL(prediction_history); NOT_NULL(beam_branch_history); NOT_NULL(log_probs);
    NOT_NULL(final_t); NOT_NULL(final_u); NOT_NULL(final_is_finished); NOT_NULL(ordered_beam_branch); NOT_NULL(ordered_tone);
    const size_t BW = n2(batch_size, beam_width), B = (size_t)(batch_size > 0 ? batch_size : 0);
    const size_t BSW = BW * (size_t)(steps > 0 ? steps : 0);
    if (is_device_pointer(h)) {
        tone_decode_loop(h, input_length, log_prob_history, is_finished, t, u, batch_size, steps, beam_width, tone_class_size,
                         empty_tone_id, prediction_history, beam_branch_history, log_probs, final_t, final_u,
                         final_is_finished, ordered_beam_branch, ordered_tone, current_stream());
        return;
    }
    HostCall c;
    auto dh = c.in(h, BSW * (size_t)(tone_class_size > 0 ? tone_class_size : 0));
    auto dil = c.in(input_length, B);
    auto dl = c.in(log_prob_history, BW); auto df = c.in(is_finished, BW); auto dt = c.in(t, BW); auto du = c.in(u, BW);
    auto oph = c.out(prediction_history, BSW); auto obh = c.out(beam_branch_history, BSW);
    auto ol = c.out(log_probs, BW); auto ot = c.out(final_t, BW); auto ou = c.out(final_u, BW);
    auto of = c.out(final_is_finished, BW); auto oo = c.out(ordered_beam_branch, BSW); auto od = c.out(ordered_tone, BSW);
    tone_decode_loop(dh, dil, dl, df, dt, du, batch_size, steps, beam_width, tone_class_size, empty_tone_id, oph, obh, ol, ot,
                     ou, of, oo, od, c.stream());
    c.finish();
}

// ---- lattice forward-backward ----------------------------------------------------------------
size_t ssnt_tts_forward_backward_workspace_bytes(int batch_size, int max_t, int max_u) {
    return fb_workspace_bytes(batch_size, max_t, max_u);
}

void ssnt_tts_forward_backward(const float* log_emit, const float* log_shift, const int* t_len,
                               const int* u_len, int batch_size, int max_t, int max_u,
                               float* log_likelihood, float* loss, float* grad_emit, float* grad_shift,
                               void* workspace, size_t workspace_bytes) {
    NOT_NULL(log_emit); NOT_NULL(log_shift); NOT_NULL(log_likelihood); NOT_NULL(grad_emit); NOT_NULL(grad_shift);
    if (is_device_pointer(log_emit)) {
        FbArgs a{log_emit, log_shift, t_len, u_len, batch_size, max_t, max_u, log_likelihood, loss,
                 grad_emit, grad_shift, workspace, workspace_bytes};
        launch_forward_backward(a, current_stream());
        return;
    }
    // Host buffers: the chunked staging pipeline above, in passes of at most kStageBytes.
    const size_t slab = n2(max_t, max_u) * sizeof(float);
    const int B = batch_size > 0 ? batch_size : 0;
    const size_t per_b = 4 * slab + 3 * sizeof(int);
    const int pass_b = per_b * B <= kStageBytes ? B : (int)(kStageBytes / per_b > 0 ? kStageBytes / per_b : 1);
    for (int p0 = 0; p0 < B; p0 += pass_b) {
        const int nb_pass = p0 + pass_b <= B ? pass_b : B - p0;
        const size_t o = (size_t)p0 * (slab / sizeof(float));
        std::vector<HostArray> ins = {{log_emit + o, nullptr, slab, 4}, {log_shift + o, nullptr, slab, 5}};
        if (t_len) ins.push_back({t_len + p0, nullptr, sizeof(int), 9});
        if (u_len) ins.push_back({u_len + p0, nullptr, sizeof(int), 10});
        std::vector<HostArray> outs = {{nullptr, grad_emit + o, slab, 6}, {nullptr, grad_shift + o, slab, 7},
                                       {nullptr, log_likelihood + p0, sizeof(float), 8}};
        const int nchunks = lattice_chunks(nb_pass, 2 * slab * nb_pass);
        const int per = (nb_pass + nchunks - 1) / nchunks;
        const size_t ws_each = (fb_workspace_bytes(per, max_t, max_u) + 255) & ~(size_t)255;
        char* ws = (char*)device_scratch(0, ws_each * nchunks);
        const int i_tl = t_len ? 2 : -1, i_ul = u_len ? (t_len ? 3 : 2) : -1;
        staged_lattice_pass(ins, outs, nb_pass, nchunks, [&](int b0, int nb, int c, cudaStream_t s) {
            const size_t e = (size_t)b0 * (slab / sizeof(float));
            FbArgs a{(const float*)ins[0].dev + e, (const float*)ins[1].dev + e,
                     i_tl >= 0 ? (const int*)ins[i_tl].dev + b0 : nullptr, i_ul >= 0 ? (const int*)ins[i_ul].dev + b0 : nullptr,
                     nb, max_t, max_u, (float*)outs[2].dev + b0, nullptr, (float*)outs[0].dev + e, (float*)outs[1].dev + e,
                     ws + (size_t)c * ws_each, ws_each};
            launch_forward_backward(a, s);
        });
    }
    check_error_flag_or_panic();
    if (loss) {
        // loss = -sum_b ll[b] in batch order, in double (what the kernel's own reduction computes)
        double acc = 0.0;
        for (int b2 = 0; b2 < B; ++b2) acc -= (double)log_likelihood[b2];
        *loss = (float)acc;
    }
}

// Raw-logit entry: log_emit = log sigmoid(z), log_shift = log sigmoid(-z) fused into the kernels' loads,
// the gradient chained through them.  12 bytes per lattice cell cross the memory bus instead of 16 (plus the upstream
// log_sigmoid kernels' own traffic, which disappears).
size_t ssnt_tts_forward_backward_logits_workspace_bytes(int batch_size, int max_t, int max_u) {
    return fb_logits_workspace_bytes(batch_size, max_t, max_u);
}

void ssnt_tts_forward_backward_logits(const float* logits, const int* t_len, const int* u_len, int batch_size, int max_t,
                                      int max_u, float* log_likelihood, float* loss, float* grad_logits, void* workspace,
                                      size_t workspace_bytes) {
    NOT_NULL(logits); NOT_NULL(log_likelihood); NOT_NULL(grad_logits);
    if (is_device_pointer(logits)) {
        FbArgs a{nullptr, nullptr, t_len, u_len, batch_size, max_t, max_u, log_likelihood, loss, nullptr, nullptr, workspace,
                 workspace_bytes};
        a.logits = logits;
        a.grad_logits = grad_logits;
        launch_forward_backward(a, current_stream());
        return;
    }
    // Host buffers: the chunked staging pipeline of ssnt_tts_forward_backward, one tensor each way.
    const size_t slab = n2(max_t, max_u) * sizeof(float);
    const int B = batch_size > 0 ? batch_size : 0;
    const size_t per_b = 2 * slab + 3 * sizeof(int);
    const int pass_b = per_b * B <= kStageBytes ? B : (int)(kStageBytes / per_b > 0 ? kStageBytes / per_b : 1);
    for (int p0 = 0; p0 < B; p0 += pass_b) {
        const int nb_pass = p0 + pass_b <= B ? pass_b : B - p0;
        const size_t o = (size_t)p0 * (slab / sizeof(float));
        std::vector<HostArray> ins = {{logits + o, nullptr, slab, 4}};
        if (t_len) ins.push_back({t_len + p0, nullptr, sizeof(int), 9});
        if (u_len) ins.push_back({u_len + p0, nullptr, sizeof(int), 10});
        std::vector<HostArray> outs = {{nullptr, grad_logits + o, slab, 6}, {nullptr, log_likelihood + p0, sizeof(float), 8}};
        const int nchunks = lattice_chunks(nb_pass, slab * nb_pass);
        const int per = (nb_pass + nchunks - 1) / nchunks;
        const size_t ws_each = (fb_logits_workspace_bytes(per, max_t, max_u) + 255) & ~(size_t)255;
        char* ws = (char*)device_scratch(0, ws_each * nchunks);
        const int i_tl = t_len ? 1 : -1, i_ul = u_len ? (t_len ? 2 : 1) : -1;
        staged_lattice_pass(ins, outs, nb_pass, nchunks, [&](int b0, int nb, int c, cudaStream_t s) {
            const size_t e = (size_t)b0 * (slab / sizeof(float));
            FbArgs a{nullptr, nullptr, i_tl >= 0 ? (const int*)ins[i_tl].dev + b0 : nullptr,
                     i_ul >= 0 ? (const int*)ins[i_ul].dev + b0 : nullptr, nb, max_t, max_u, (float*)outs[1].dev + b0, nullptr,
                     nullptr, nullptr, ws + (size_t)c * ws_each, ws_each};
            a.logits = (const float*)ins[0].dev + e;
            a.grad_logits = (float*)outs[0].dev + e;
            launch_forward_backward(a, s);
        });
    }
    check_error_flag_or_panic();
    if (loss) {
        double acc = 0.0;  // loss = -sum_b ll[b] in batch order, in double (as the kernel's own reduction)
        for (int b2 = 0; b2 < B; ++b2) acc -= (double)log_likelihood[b2];
        *loss = (float)acc;
    }
}

size_t tone_latent_forward_backward_workspace_bytes(int batch_size, int max_t, int max_u, int tone_class_size) {
    return tone_fb_workspace_bytes(batch_size, max_t, max_u, tone_class_size);
}

void tone_latent_forward_backward(const float* log_emit, const float* log_shift, const float* log_tone,
                                  const int* t_len, const int* u_len, int batch_size, int max_t, int max_u,
                                  int tone_class_size, float* log_likelihood, float* loss, float* grad_emit,
                                  float* grad_shift, float* grad_tone, void* workspace, size_t workspace_bytes) {
    NOT_NULL(log_emit); NOT_NULL(log_shift); NOT_NULL(log_tone); NOT_NULL(log_likelihood);
    NOT_NULL(grad_emit); NOT_NULL(grad_shift); NOT_NULL(grad_tone);
    if (is_device_pointer(log_emit)) {
        ToneFbArgs a{log_emit, log_shift, log_tone, t_len, u_len, batch_size, max_t, max_u, tone_class_size,
                     log_likelihood, loss, grad_emit, grad_shift, grad_tone, workspace, workspace_bytes};
        launch_tone_forward_backward(a, current_stream());
        return;
    }
    // Host buffers: the same chunked staging pipeline as ssnt_tts_forward_backward above.
    const int K = tone_class_size > 0 ? tone_class_size : 0;
    const int B = batch_size > 0 ? batch_size : 0;
    const size_t slab = n2(max_t, max_u) * (size_t)K * sizeof(float), tslab = n2(max_u, K) * sizeof(float);
    const size_t per_b = 4 * slab + 2 * tslab + 3 * sizeof(int);
    const int pass_b = per_b * B <= kStageBytes ? B : (int)(kStageBytes / per_b > 0 ? kStageBytes / per_b : 1);
    for (int p0 = 0; p0 < B; p0 += pass_b) {
        const int nb_pass = p0 + pass_b <= B ? pass_b : B - p0;
        const size_t o = (size_t)p0 * (slab / sizeof(float)), ot = (size_t)p0 * (tslab / sizeof(float));
        std::vector<HostArray> ins = {{log_tone + ot, nullptr, tslab, 11}, {log_emit + o, nullptr, slab, 4},
                                      {log_shift + o, nullptr, slab, 5}};
        if (t_len) ins.push_back({t_len + p0, nullptr, sizeof(int), 9});
        if (u_len) ins.push_back({u_len + p0, nullptr, sizeof(int), 10});
        std::vector<HostArray> outs = {{nullptr, grad_emit + o, slab, 6}, {nullptr, grad_shift + o, slab, 7},
                                       {nullptr, grad_tone + ot, tslab, 12}, {nullptr, log_likelihood + p0, sizeof(float), 8}};
        const int nchunks = lattice_chunks(nb_pass, 2 * slab * nb_pass);
        const int per = (nb_pass + nchunks - 1) / nchunks;
        // every chunk's workspace starts 256-byte aligned (the block-float kernel needs 16)
        const size_t ws_each = (tone_fb_workspace_bytes(per, max_t, max_u, K) + 255) & ~(size_t)255;
        char* ws = (char*)device_scratch(1, ws_each * nchunks);
        const int i_tl = t_len ? 3 : -1, i_ul = u_len ? (t_len ? 4 : 3) : -1;
        staged_lattice_pass(ins, outs, nb_pass, nchunks, [&](int b0, int nb, int c, cudaStream_t s) {
            const size_t e = (size_t)b0 * (slab / sizeof(float)), et = (size_t)b0 * (tslab / sizeof(float));
            ToneFbArgs a{(const float*)ins[1].dev + e, (const float*)ins[2].dev + e, (const float*)ins[0].dev + et,
                         i_tl >= 0 ? (const int*)ins[i_tl].dev + b0 : nullptr, i_ul >= 0 ? (const int*)ins[i_ul].dev + b0 : nullptr,
                         nb, max_t, max_u, tone_class_size, (float*)outs[3].dev + b0, nullptr, (float*)outs[0].dev + e,
                         (float*)outs[1].dev + e, (float*)outs[2].dev + et, ws + (size_t)c * ws_each, ws_each};
            launch_tone_forward_backward(a, s);
        });
    }
    check_error_flag_or_panic();
    if (loss) {
        double acc = 0.0;  // loss = -sum_b ll[b] in batch order, in double (as the kernel's own reduction)
        for (int b2 = 0; b2 < B; ++b2) acc -= (double)log_likelihood[b2];
        *loss = (float)acc;
    }
}

// ---- runtime side channel ----------------------------------------------------------------------
void ssnt_tts_set_stream(void* cuda_stream) { set_stream((cudaStream_t)cuda_stream); }
void* ssnt_tts_get_stream(void) { return (void*)current_stream(); }
void ssnt_tts_set_memory_space(int space) {
    SSNT_ASSERT(space >= 0 && space <= 2, "ssnt_tts_set_memory_space: 0 auto, 1 host, 2 device");
    set_space(space);
}
void ssnt_tts_synchronize(void) {
    SSNT_CUDA(cudaStreamSynchronize(current_stream()));
    check_error_flag_or_panic();
}
unsigned ssnt_tts_last_error(void) {
    SSNT_CUDA(cudaStreamSynchronize(current_stream()));
    return read_and_clear_error_flag();
}
void ssnt_tts_set_fb_kernel(int kind) { fb_force_kernel_kind(kind); }
void ssnt_tts_debug_set_fb_stats(void* dev_buffer) { fb_set_stats_buffer((long long*)dev_buffer); }
int ssnt_tts_get_fb_kernel_used(void) { return fb_last_kernel_kind(); }
void ssnt_tts_set_tone_kernel(int kind) { tone_force_kernel_kind(kind); }
int ssnt_tts_get_tone_kernel_used(void) { return tone_last_kernel_kind(); }
unsigned ssnt_tts_fb_fallback_count(void) {
    SSNT_CUDA(cudaStreamSynchronize(current_stream()));
    return read_fallback_counter();
}
const char* ssnt_tts_backend(void) { return "cuda-sm_100a"; }
void ssnt_tts_fill_i32(int* dst, size_t n, int value) {
    NOT_NULL(dst);
    if (is_device_pointer(dst)) {
        device_fill_i32(dst, n, value, current_stream());
    } else {
        for (size_t i = 0; i < n; ++i) dst[i] = value;
    }
}
int ssnt_tts_debug_host_copy(void* dst, const void* src, size_t bytes) {
    NOT_NULL(dst); NOT_NULL(src);
    HostCopier c;
    c.copy(dst, src, bytes);
    return host_copy_threads();
}

// ---- multi-GPU loss exchange (batch-sharded training) ---------------------------------------------
void ssnt_tts_loss_exchange_export(int world_size, unsigned char* handle_out) {
    NOT_NULL(handle_out);
    loss_exchange_export(world_size, handle_out);
}
void ssnt_tts_loss_exchange_connect(int rank, int world_size, const unsigned char* handles) {
    NOT_NULL(handles);
    loss_exchange_connect(rank, world_size, handles);
}
void ssnt_tts_loss_exchange_disconnect(void) { loss_exchange_disconnect(); }
void ssnt_tts_loss_allreduce(float* out) {
    NOT_NULL(out);
    if (is_device_pointer(out)) {
        launch_loss_allreduce(out, current_stream());
        return;
    }
    float* d = (float*)device_scratch(39, sizeof(float));
    launch_loss_allreduce(d, current_stream());
    SSNT_CUDA(cudaMemcpyAsync(out, d, sizeof(float), cudaMemcpyDeviceToHost, current_stream()));
    SSNT_CUDA(cudaStreamSynchronize(current_stream()));
    check_error_flag_or_panic();
}

}  // extern "C"
#pragma GCC visibility pop
