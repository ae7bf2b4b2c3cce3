/*
 * ssnt_tts_c.h — C-ABI of the B200 (sm_100a) backend for the ssnt-tts-rust hot path.
 *
 * Drop-in boundary: the seven functions of the first block are exactly the `#[no_mangle] pub
 * extern fn` symbols of the reference's `ssnt_tts_c` crate (ssnt_tts_c/src/lib.rs), with the
 * argument order, int32 sizes, 1-byte bool and caller-owned dense row-major buffers the
 * reference's TensorFlow ops bind (ssnt-tts-tensorflow/src/[name]_op.cc, cited per function).
 * libssnt_tts_c.so replaces the reference's libssnt_tts_c.a at link time.
 *
 * Memory spaces.  Every pointer argument of one call must live in the same memory space:
 *   - device pointers (cudaMalloc / framework GPU tensors): the call only enqueues kernels on
 *     the stream set with ssnt_tts_set_stream() and returns; nothing is copied.
 *   - host pointers (what the reference's DEVICE_CPU ops pass today): inputs are copied to the
 *     GPU, the same kernels run, outputs are copied back and the call returns when the host
 *     buffers are complete.  There is no CPU implementation behind this header.  The lattice
 *     calls cut the batch into chunks and overlap upload, kernels and download; ordinary
 *     (pageable) caller buffers go through the library's own page-locked staging buffers, the
 *     host-side copies spread over a few library threads (SSNT_COPY_THREADS, default
 *     min(12, cores/2) with the caller's); buffers the caller page-locked are used in place.
 *     The small decoding calls issue plain cudaMemcpyAsync on the caller's pointers.
 * The space is detected from the first pointer argument (cudaPointerGetAttributes) unless
 * ssnt_tts_set_memory_space() pins it.
 *
 * Errors.  The reference panics (process abort) on null pointers and on the data-dependent
 * asserts src/v2.rs:292 and src/v2_util.rs:58.  Null pointers abort here as well.  The
 * data-dependent asserts are detected on the device: host-pointer calls abort before
 * returning (same observable behaviour), device-pointer calls raise a flag that
 * ssnt_tts_synchronize() turns into the abort (ssnt_tts_last_error() reads it without
 * aborting).
 */
#ifndef SSNT_TTS_C_H_
#define SSNT_TTS_C_H_

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Block 1 — the reference's own symbols (signatures kept verbatim).
 * ---------------------------------------------------------------------------------------- */

/* v1 Emit/Shift beam step, single batch.  Replaces ssnt_tts_c/src/lib.rs:10-83 (→
 * src/lib.rs:121-230); bound by ssnt_tts_beam_search_decode_op.cc:5-8.
 * h[W,2] log_prob_history[W] is_finished[W] t[W] u[W] → all outputs [W]. */
void ssnt_tts_beam_search_decode(const float *h, const float *log_prob_history,
                                 const bool *is_finished, const int *t, const int *u, int max_t,
                                 int beam_width, int *prediction, float *log_probs, int *next_t,
                                 int *next_u, bool *next_is_finished, int *beam_branch);

/* Best-beam back-trace.  Replaces ssnt_tts_c/src/lib.rs:86-116 (→ src/util.rs:20-33); bound by
 * ssnt_extract_best_beam_branch_op.cc:6-8.  beam_branch,t_history [max_u,W] → [max_u] each. */
void ssnt_extract_best_beam_branch(int best_final_branch, const int *beam_branch,
                                   const int *t_history, int beam_width, int max_u,
                                   int *best_beam_branch, int *best_t_history);

/* v2 duration-class beam step.  Replaces ssnt_tts_c/src/lib.rs:118-218 (→ src/v2.rs:221-339);
 * bound by ssnt_tts_v2_beam_search_decode_op.cc:5-26.  h[B,W,D], duration_table[D],
 * input_length/output_length[B], everything else [B,W]. */
void ssnt_tts_v2_beam_search_decode(const float *h, const float *log_prob_history,
                                    const bool *is_finished, const int *total_duration,
                                    const int *duration_table, const int *t, const int *u,
                                    const int *input_length, const int *output_length,
                                    int batch_size, int beam_width, int duration_class_size,
                                    int zero_duration_id, bool allow_skip, bool test_mode,
                                    int *prediction, float *log_probs, int *next_t, int *next_u,
                                    bool *next_is_finished, int *next_total_duration,
                                    int *beam_branch);

/* Back-trace of every final beam.  Replaces ssnt_tts_c/src/lib.rs:220-241 (→
 * src/v2_util.rs:6-36); bound by ssnt_order_beam_branch_op.cc:6-11.
 * final_branch[B,W], beam_branch[B,T,W] → ordered_beam_branch[B,W,T]. */
void ssnt_order_beam_branch(const int *final_branch, const int *beam_branch, int batch_size,
                            int beam_width, int max_t, int *ordered_beam_branch);

/* Duration → source-index upsampling.  Replaces ssnt_tts_c/src/lib.rs:244-265 (→
 * src/v2_util.rs:39-66); bound by upsample_source_indexes_op.cc:6-12.  duration[B,W,T],
 * output_length[B,W] → upsampled_source_indexes[B,W,max_u]; only the first output_length
 * slots of each row are written (the caller pre-fills the rest, op.cc:75). */
void ssnt_upsample_source_indexes(const int *duration, const int *output_length, int batch_size,
                                  int beam_width, int max_t, int max_u,
                                  int *upsampled_source_indexes);

/* Tone-latent beam step.  Replaces ssnt_tts_c/src/lib.rs:267-343 (→
 * src/tone_latent.rs:144-234); bound by tone_latent_beam_search_decode_op.cc:5-20. */
void tone_latent_beam_search_decode(const float *h, const float *log_prob_history,
                                    const bool *is_finished, const int *t, const int *u,
                                    const int *input_length, int batch_size, int beam_width,
                                    int tone_class_size, int empty_tone_id, int *prediction,
                                    float *log_probs, int *next_t, int *next_u,
                                    bool *next_is_finished, int *beam_branch);

/* Batched Levenshtein distance, int32, bit-exact.  Replaces ssnt_tts_c/src/lib.rs:346-381 (→
 * src/edit_distance.rs:6-60); bound by ssnt_tts_edit_distance.cc:6-9.
 * a,b[B,max_length], a_lengths,b_lengths[B] → distance[B]. */
void tone_latent_levenshtein_edit_distance(const int *a, const int *b, const int *a_lengths,
                                           const int *b_lengths, int batch_size, int max_length,
                                           int *distance);

/* ------------------------------------------------------------------------------------------
 * Block 2 — lattice forward-backward (no counterpart in the reference, which has no loss /
 * gradient code; specification: DESIGN.md §2, SURVEY.md §8 a-FB / a-TL).  Same style: void,
 * int32 sizes, caller-owned buffers.
 *
 * T = output frames (serial axis), U = input tokens; tensors are [B, max_t, max_u] row-major
 * (U fastest).  t_len/u_len may be NULL (= full lengths).  Outputs:
 *   log_likelihood[B]   natural-log likelihood per utterance (-inf if no path exists)
 *   loss[1]             -sum_b log_likelihood[b]  (may be NULL)
 *   grad_emit/shift     d log_likelihood / d log-prob = posterior transition occupancy;
 *                       every element of the [B,max_t,max_u] tensors is written (padded,
 *                       unreachable and infeasible cells get exactly 0).
 * workspace: device memory of at least ssnt_tts_forward_backward_workspace_bytes(); NULL lets
 * the library use an internal grow-only buffer (per host thread).
 * ---------------------------------------------------------------------------------------- */
size_t ssnt_tts_forward_backward_workspace_bytes(int batch_size, int max_t, int max_u);
void ssnt_tts_forward_backward(const float *log_emit, const float *log_shift, const int *t_len,
                               const int *u_len, int batch_size, int max_t, int max_u,
                               float *log_likelihood, float *loss, float *grad_emit,
                               float *grad_shift, void *workspace, size_t workspace_bytes);

/* Raw-logit variant of ssnt_tts_forward_backward for models whose emit/shift scores are the two-way split of ONE logit,
 * log_emit = log sigmoid(z), log_shift = log sigmoid(-z) (the reference's rows are such two-way distributions,
 * tests/test_decoding.rs:25-29): the log-sigmoids are formed inside the kernels' loads and the gradient is chained
 * through them, grad_logits = dLL/dz = grad_emit * sigmoid(-z) - grad_shift * sigmoid(z).  logits, grad_logits [B,T,U];
 * everything else as ssnt_tts_forward_backward.  Results equal that call on (log sigmoid(z), log sigmoid(-z)) to
 * rounding.  12 bytes per lattice cell cross the memory bus instead of 16. */
size_t ssnt_tts_forward_backward_logits_workspace_bytes(int batch_size, int max_t, int max_u);
void ssnt_tts_forward_backward_logits(const float *logits, const int *t_len, const int *u_len, int batch_size,
                                      int max_t, int max_u, float *log_likelihood, float *loss,
                                      float *grad_logits, void *workspace, size_t workspace_bytes);

/* Tone-latent marginalised lattice: log_emit/log_shift [B,max_t,max_u,K], log_tone [B,max_u,K]
 * (log-prior of each token's tone class; cf. tone_class_size of src/tone_latent.rs:79-95). */
size_t tone_latent_forward_backward_workspace_bytes(int batch_size, int max_t, int max_u,
                                                    int tone_class_size);
void tone_latent_forward_backward(const float *log_emit, const float *log_shift,
                                  const float *log_tone, const int *t_len, const int *u_len,
                                  int batch_size, int max_t, int max_u, int tone_class_size,
                                  float *log_likelihood, float *loss, float *grad_emit,
                                  float *grad_shift, float *grad_tone, void *workspace,
                                  size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------
 * Block 3 — runtime side channel (the reference ABI has no stream / device notion).
 * ---------------------------------------------------------------------------------------- */
/* cudaStream_t used by subsequent calls of the calling host thread (NULL = legacy stream). */
void ssnt_tts_set_stream(void *cuda_stream);
void *ssnt_tts_get_stream(void);
/* 0 = auto-detect per call (default), 1 = treat pointers as host, 2 = as device. */
void ssnt_tts_set_memory_space(int space);
/* Waits for the current stream and aborts if a device-side assert fired (see "Errors"). */
void ssnt_tts_synchronize(void);
/* Waits for the current stream, returns and clears the error bits without aborting:
 * 1 = v2 empty beam (src/v2.rs:292), 2 = upsample length mismatch (src/v2_util.rs:58),
 * 4 = tone-latent empty beam, 8 = back-trace index out of range, 16 = a peer never delivered
 * its loss entry (loss exchange). */
unsigned ssnt_tts_last_error(void);
/* Forward-backward kernel selection for tests/benchmarks: -1 auto, 0 generic, 1 log-domain
 * warp/TMA, 2 block-float fused (cluster of 2 CTAs per utterance), 3 = 2 with forced log re-run,
 * 4 block-float split-role (cluster of 4 CTAs per utterance), 5 = 4 with forced log re-run,
 * 6 time-parallel block-float (chunk transfer operators built concurrently, banded mat-vec sweep
 * over the chunk boundaries, chunk interiors filled independently; the auto choice whenever
 * max_u % 4 == 0 and max_u <= 256), 7 = 6 with forced log re-run, 8 warp-serial block-float (one warp
 * per utterance, alpha checkpoints forward, chunked beta + gradients backward: 25 bytes per cell of
 * memory traffic; the auto choice for large batches), 9 = 8 with forced log re-run, 10 = 6 with 8-frame
 * chunks (max_u > 128 only; what 6 picks by itself for wide lattices at medium batch sizes). */
void ssnt_tts_set_fb_kernel(int kind);
int ssnt_tts_get_fb_kernel_used(void);
/* Tone-latent lattice kernel selection for tests/benchmarks: -1 auto, 0 log domain only, 1 block-float split-role
 * (cluster of four CTAs per utterance; K = 4, max_u in {32,64,128}; the auto choice for small batches), 2 block-float
 * warp-serial (one warp per utterance, alpha checkpoints forward, chunked beta + gradients backward; K in {2,4,8},
 * max_u in {32,64,128,256} with 4 <= max_u*K/32 <= 32; the auto choice from two utterances per SM and for every shape
 * 1 does not take), 3 = 2 with every utterance re-run in the log domain. */
void ssnt_tts_set_tone_kernel(int kind);
int ssnt_tts_get_tone_kernel_used(void);
/* Cumulative number of utterances the block-float kernel had to re-run in the log domain
 * because their dynamic range did not fit (waits for the current stream). */
unsigned ssnt_tts_fb_fallback_count(void);
/* Profiling aid: device buffer of (2*B*8*16 + 4096) int64 that the block-float kernel fills with per-warp
 * cycle counters (total, and cycles blocked on each hand-off barrier); NULL switches it off. */
void ssnt_tts_debug_set_fb_stats(void *dev_buffer);
const char *ssnt_tts_backend(void); /* "cuda-sm_100a" */
/* Pre-fill of an int32 output in the memory space of `dst` (device: one small kernel on the current stream).  The
 * reference's ops pre-fill on the host (ssnt_tts_beam_search_decode_op.cc:91, ssnt_tts_v2_beam_search_decode_op.cc:212,
 * tone_latent_beam_search_decode_op.cc:168, upsample_source_indexes_op.cc:75); DEVICE_GPU registrations call this. */
void ssnt_tts_fill_i32(int *dst, size_t n, int value);
/* The host-to-host copy the host-pointer lattice calls stage pageable buffers with (spread over the library's copy
 * threads; SSNT_COPY_THREADS sets their number, the calling thread included).  Returns the thread count.  Test aid. */
int ssnt_tts_debug_host_copy(void *dst, const void *src, size_t bytes);

/* ------------------------------------------------------------------------------------------
 * Block 2b — whole-loop decoding (no counterpart in the reference's ABI).
 * The reference decodes with one call per output step (ssnt-tts-tensorflow/ssnt_tts_tensorflow/__init__.py:33-73 is
 * meant for a tf.while_loop), then back-traces (src/v2_util.rs:6-36) and upsamples (src/v2_util.rs:39-66).  When the
 * per-step scores are known up front, h[B, steps, W, classes], one launch does all of it and the results equal the
 * per-step calls bit for bit:
 *   every step s = 0..steps-1: ssnt_tts_v2_beam_search_decode (src/v2.rs:221-339) on h[:, s] and the running state;
 *     prediction_history, beam_branch_history [B, steps, W] record each step's prediction and beam_branch
 *   log_probs / final_* [B, W]: the state after the last step
 *   ordered_beam_branch [B, W, steps] = ssnt_order_beam_branch(final_branch = 0..W-1, beam_branch_history)
 *   duration [B, W, steps] = duration_table[prediction_history gathered along ordered_beam_branch]
 *   upsampled_source_indexes [B, W, max_u] (may be NULL; caller pre-filled) =
 *     ssnt_upsample_source_indexes(duration, output_length = final_total_duration)
 * The initial state pointers (log_prob_history, is_finished, total_duration, t, u; [B, W]) may be NULL = zeros / false.
 * tone_latent_decode_loop is the same over tone_latent_beam_search_decode (src/tone_latent.rs:144-234);
 * ordered_tone [B, W, steps] = prediction_history gathered along ordered_beam_branch.
 * ---------------------------------------------------------------------------------------- */
void ssnt_tts_v2_decode_loop(const float *h, const int *duration_table, const int *input_length,
                             const int *output_length, const float *log_prob_history,
                             const bool *is_finished, const int *total_duration, const int *t, const int *u,
                             int batch_size, int steps, int beam_width, int duration_class_size,
                             int zero_duration_id, bool allow_skip, bool test_mode, int max_u,
                             int *prediction_history, int *beam_branch_history, float *log_probs, int *final_t,
                             int *final_u, bool *final_is_finished, int *final_total_duration,
                             int *ordered_beam_branch, int *duration, int *upsampled_source_indexes);
void tone_latent_decode_loop(const float *h, const int *input_length, const float *log_prob_history,
                             const bool *is_finished, const int *t, const int *u, int batch_size, int steps,
                             int beam_width, int tone_class_size, int empty_tone_id, int *prediction_history,
                             int *beam_branch_history, float *log_probs, int *final_t, int *final_u,
                             bool *final_is_finished, int *ordered_beam_branch, int *ordered_tone);

/* ------------------------------------------------------------------------------------------
 * Block 4 — multi-GPU loss exchange (no counterpart in the reference, which is single-process).
 * Utterances are independent, so N GPUs each take a batch shard (the reference's par_chunks over
 * the batch, src/v2.rs:227); the one value that crosses GPUs is the scalar loss.  Once connected,
 * the kernel that reduces a forward_backward call's loss also stores {loss, call number} into
 * every rank's slot buffer with one 64-bit NVLink store per peer: nothing extra is launched, the
 * host issues no collective, and the exchange replays with a CUDA graph that captured the call.
 * Every rank must make the same sequence of loss-producing calls (device pointers, loss != NULL).
 *   export   allocates this rank's slot buffer and returns its 64-byte CUDA IPC handle
 *   connect  handles[world_size][64] gathered from all ranks (any transport), own included
 *   allreduce  out[0] = sum over ranks of the latest call's loss (same bits on every rank);
 *              waits, on the stream, for the peers' entries; host or device pointer
 * ---------------------------------------------------------------------------------------- */
void ssnt_tts_loss_exchange_export(int world_size, unsigned char *handle_out /* [64] */);
void ssnt_tts_loss_exchange_connect(int rank, int world_size, const unsigned char *handles);
void ssnt_tts_loss_exchange_disconnect(void);
void ssnt_tts_loss_allreduce(float *out);

#ifdef __cplusplus
}
#endif
#endif /* SSNT_TTS_C_H_ */
