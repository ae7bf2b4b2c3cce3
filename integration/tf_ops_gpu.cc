// DEVICE_GPU registrations of the reference's seven TensorFlow ops over libssnt_tts_c.so, plus the two lattice ops
// the reference does not have.  Drop this file next to ssnt-tts-tensorflow/src/*.cc (whose
// REGISTER_OP definitions and DEVICE_CPU kernels stay as they are) and link libssnt_tts_c.so instead of the Rust
// crate's libssnt_tts_c.a: the same seven C symbols then receive device pointers here and host pointers there.
//
// What each kernel checks mirrors the corresponding DEVICE_CPU Compute() (shape ranks, beam width, batch sizes; cited
// per op).  Differences that the device forces:
//   * scalar inputs (max_t, best_final_branch, max_u, out_of_range_source_index) are pinned to host memory;
//   * outputs the reference pre-fills on the host are pre-filled on the device (ssnt_tts_fill_i32);
//   * the library enqueues on the op's stream (ssnt_tts_set_stream) and returns; data-dependent panics of the
//     reference (src/v2.rs:292, src/v2_util.rs:58) surface at ssnt_tts_synchronize() / ssnt_tts_last_error().
// TensorFlow is not available in this repository's image: the file is type-checked against integration/tf_mock
// (tests/test_tf_ops_syntax.py), not built or run.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"

#include "ssnt_tts_c.h"

namespace tf = tensorflow;

namespace ssnt_gpu {

// ---- small helpers: fetch an input and check its rank, allocate an output and return its device pointer ----------
struct In {
    const tf::Tensor* t = nullptr;
    bool get(tf::OpKernelContext* ctx, const char* name, int rank) {
        if (!ctx->input(name, &t).ok() || t == nullptr) {
            ctx->CtxFailure(tf::errors::InvalidArgument("missing input ", name));
            return false;
        }
        if (t->dims() != rank) {
            ctx->CtxFailure(tf::errors::InvalidArgument(name, " is not a ", rank, "D-Tensor"));
            return false;
        }
        return true;
    }
    tf::int64 dim(int i) const { return t->dim_size(i); }
    template <typename T> const T* ptr() const { return t->flat<T>().data(); }
    template <typename T> T host_scalar() const { return t->scalar<T>()(); }
};

template <typename T>
T* out(tf::OpKernelContext* ctx, const char* name, const tf::TensorShape& shape) {
    tf::Tensor* o = nullptr;
    if (!ctx->allocate_output(name, shape, &o).ok() || o == nullptr) {
        ctx->CtxFailure(tf::errors::InvalidArgument("cannot allocate output ", name));
        return nullptr;
    }
    return o->flat<T>().data();
}

inline void bind_stream(tf::OpKernelContext* ctx) { ssnt_tts_set_stream((void*)ctx->eigen_gpu_device().stream()); }

#define SSNT_CHECK(ctx, cond, ...) OP_REQUIRES(ctx, cond, tf::errors::InvalidArgument(__VA_ARGS__))

// ---- SSNTBeamSearchDecode (ssnt_tts_beam_search_decode_op.cc:28-139): v1 Emit/Shift step, single batch ----------------
class BeamSearchDecodeGpu : public tf::OpKernel {
public:
    explicit BeamSearchDecodeGpu(tf::OpKernelConstruction* c) : tf::OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("beam_width", &w_)); }
    void Compute(tf::OpKernelContext* ctx) override {
        In h, lph, fin, t, u, max_t;
        if (!h.get(ctx, "h", 2) || !lph.get(ctx, "log_prob_history", 1) || !fin.get(ctx, "is_finished", 1) ||
            !t.get(ctx, "t", 1) || !u.get(ctx, "u", 1) || !max_t.get(ctx, "max_t", 0))
            return;
        SSNT_CHECK(ctx, h.dim(0) == w_ && lph.dim(0) == w_ && fin.dim(0) == w_ && t.dim(0) == w_ && u.dim(0) == w_,
                   "Incompatible beam width");
        SSNT_CHECK(ctx, h.dim(1) == 2, "h must hold the Emit and Shift scores: [beam_width, 2]");
        const tf::TensorShape s({(tf::int64)w_});
        int* pred = out<int>(ctx, "prediction", s);
        float* lp = out<float>(ctx, "log_prob", s);
        int* nt = out<int>(ctx, "next_t", s);
        int* nu = out<int>(ctx, "next_u", s);
        bool* nf = out<bool>(ctx, "next_is_finished", s);
        int* bb = out<int>(ctx, "beam_branch", s);
        if (!pred || !lp || !nt || !nu || !nf || !bb) return;
        bind_stream(ctx);
        ssnt_tts_fill_i32(pred, (size_t)w_, -1);  // _op.cc:91
        ssnt_tts_beam_search_decode(h.ptr<float>(), lph.ptr<float>(), fin.ptr<bool>(), t.ptr<int>(), u.ptr<int>(),
                                    max_t.host_scalar<int>(), w_, pred, lp, nt, nu, nf, bb);
    }
private:
    int w_ = 0;
};
REGISTER_KERNEL_BUILDER(Name("SSNTBeamSearchDecode").Device(tf::DEVICE_GPU).HostMemory("max_t"), BeamSearchDecodeGpu);

// ---- SSNTExtractBestBeamBranch (ssnt_extract_best_beam_branch_op.cc:22-81) --------------------------------------------
class ExtractBestBeamBranchGpu : public tf::OpKernel {
public:
    explicit ExtractBestBeamBranchGpu(tf::OpKernelConstruction* c) : tf::OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("beam_width", &w_)); }
    void Compute(tf::OpKernelContext* ctx) override {
        In best, bb, th;
        if (!best.get(ctx, "best_final_branch", 0) || !bb.get(ctx, "beam_branch", 2) || !th.get(ctx, "t_history", 2)) return;
        SSNT_CHECK(ctx, bb.dim(1) == w_ && th.dim(1) == w_, "Incompatible beam width");
        SSNT_CHECK(ctx, bb.dim(0) == th.dim(0), "beam_branch and t_history differ in length");
        const int max_u = (int)bb.dim(0);
        int* ob = out<int>(ctx, "best_beam_branch", tf::TensorShape({(tf::int64)max_u}));
        int* ot = out<int>(ctx, "best_t_history", tf::TensorShape({(tf::int64)max_u}));
        if (!ob || !ot) return;
        bind_stream(ctx);
        ssnt_extract_best_beam_branch(best.host_scalar<int>(), bb.ptr<int>(), th.ptr<int>(), w_, max_u, ob, ot);
    }
private:
    int w_ = 0;
};
REGISTER_KERNEL_BUILDER(Name("SSNTExtractBestBeamBranch").Device(tf::DEVICE_GPU).HostMemory("best_final_branch"),
                        ExtractBestBeamBranchGpu);

// ---- SSNTV2BeamSearchDecode (ssnt_tts_v2_beam_search_decode_op.cc:54-216) ---------------------------------------------
class V2BeamSearchDecodeGpu : public tf::OpKernel {
public:
    explicit V2BeamSearchDecodeGpu(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
        OP_REQUIRES_OK(c, c->GetAttr("beam_width", &w_));
        OP_REQUIRES_OK(c, c->GetAttr("duration_class_size", &d_));
        OP_REQUIRES_OK(c, c->GetAttr("zero_duration_id", &zero_));
        OP_REQUIRES_OK(c, c->GetAttr("allow_skip", &skip_));
        OP_REQUIRES_OK(c, c->GetAttr("test_mode", &test_));
    }
    void Compute(tf::OpKernelContext* ctx) override {
        In h, lph, fin, tot, tab, t, u, il, ol;
        if (!h.get(ctx, "h", 3) || !lph.get(ctx, "log_prob_history", 2) || !fin.get(ctx, "is_finished", 2) ||
            !tot.get(ctx, "total_duration", 2) || !tab.get(ctx, "duration_table", 1) || !t.get(ctx, "t", 2) ||
            !u.get(ctx, "u", 2) || !il.get(ctx, "input_length", 1) || !ol.get(ctx, "output_length", 1))
            return;
        const tf::int64 B = h.dim(0);
        SSNT_CHECK(ctx, h.dim(1) == w_ && lph.dim(1) == w_ && fin.dim(1) == w_ && tot.dim(1) == w_ && t.dim(1) == w_ && u.dim(1) == w_,
                   "Incompatible beam width");
        SSNT_CHECK(ctx, h.dim(2) == d_ && tab.dim(0) == d_, "Incompatible duration class size");
        SSNT_CHECK(ctx, lph.dim(0) == B && fin.dim(0) == B && tot.dim(0) == B && t.dim(0) == B && u.dim(0) == B && il.dim(0) == B && ol.dim(0) == B,
                   "Incompatible batch sizes");
        const tf::TensorShape s({B, (tf::int64)w_});
        int* pred = out<int>(ctx, "prediction", s);
        float* lp = out<float>(ctx, "log_prob", s);
        int* nt = out<int>(ctx, "next_t", s);
        int* nu = out<int>(ctx, "next_u", s);
        bool* nf = out<bool>(ctx, "next_is_finished", s);
        int* ntot = out<int>(ctx, "next_total_duration", s);
        int* bb = out<int>(ctx, "beam_branch", s);
        if (!pred || !lp || !nt || !nu || !nf || !ntot || !bb) return;
        bind_stream(ctx);
        ssnt_tts_fill_i32(pred, (size_t)(B * w_), zero_);  // _op.cc:149,212
        ssnt_tts_v2_beam_search_decode(h.ptr<float>(), lph.ptr<float>(), fin.ptr<bool>(), tot.ptr<int>(), tab.ptr<int>(),
                                       t.ptr<int>(), u.ptr<int>(), il.ptr<int>(), ol.ptr<int>(), (int)B, w_, d_, zero_, skip_,
                                       test_, pred, lp, nt, nu, nf, ntot, bb);
    }
private:
    int w_ = 0, d_ = 0, zero_ = 0;
    bool skip_ = false, test_ = false;
};
REGISTER_KERNEL_BUILDER(Name("SSNTV2BeamSearchDecode").Device(tf::DEVICE_GPU), V2BeamSearchDecodeGpu);

// ---- SSNTOrderBeamBranch (ssnt_order_beam_branch_op.cc:22-73) -----------------------------------------------------------
class OrderBeamBranchGpu : public tf::OpKernel {
public:
    explicit OrderBeamBranchGpu(tf::OpKernelConstruction* c) : tf::OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("beam_width", &w_)); }
    void Compute(tf::OpKernelContext* ctx) override {
        In fb, bb;
        if (!fb.get(ctx, "final_branch", 2) || !bb.get(ctx, "beam_branch", 3)) return;
        SSNT_CHECK(ctx, bb.dim(0) == fb.dim(0), "Incompatible batch sizes");
        SSNT_CHECK(ctx, bb.dim(2) == w_ && fb.dim(1) == w_, "Incompatible beam width");
        const tf::int64 B = bb.dim(0), T = bb.dim(1);
        int* o = out<int>(ctx, "ordered_beam_branch", tf::TensorShape({B, (tf::int64)w_, T}));
        if (!o) return;
        bind_stream(ctx);
        ssnt_order_beam_branch(fb.ptr<int>(), bb.ptr<int>(), (int)B, w_, (int)T, o);
    }
private:
    int w_ = 0;
};
REGISTER_KERNEL_BUILDER(Name("SSNTOrderBeamBranch").Device(tf::DEVICE_GPU), OrderBeamBranchGpu);

// ---- SSNTUpsampleSourceIndexes (upsample_source_indexes_op.cc:26-95) ------------------------------------------------------
class UpsampleSourceIndexesGpu : public tf::OpKernel {
public:
    explicit UpsampleSourceIndexesGpu(tf::OpKernelConstruction* c) : tf::OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("beam_width", &w_)); }
    void Compute(tf::OpKernelContext* ctx) override {
        In d, ol, mu, oor;
        if (!d.get(ctx, "duration", 3) || !ol.get(ctx, "output_length", 2) || !mu.get(ctx, "max_u", 0) ||
            !oor.get(ctx, "out_of_range_source_index", 0))
            return;
        SSNT_CHECK(ctx, ol.dim(0) == d.dim(0), "Incompatible batch sizes");
        SSNT_CHECK(ctx, d.dim(1) == w_ && ol.dim(1) == w_, "Incompatible beam width");
        const tf::int64 B = d.dim(0), T = d.dim(2);
        const int max_u = mu.host_scalar<int>();
        int* o = out<int>(ctx, "upsampled_source_indexes", tf::TensorShape({B, (tf::int64)w_, (tf::int64)max_u}));
        if (!o) return;
        bind_stream(ctx);
        ssnt_tts_fill_i32(o, (size_t)(B * w_ * max_u), oor.host_scalar<int>());  // _op.cc:75
        ssnt_upsample_source_indexes(d.ptr<int>(), ol.ptr<int>(), (int)B, w_, (int)T, max_u, o);
    }
private:
    int w_ = 0;
};
REGISTER_KERNEL_BUILDER(Name("SSNTUpsampleSourceIndexes").Device(tf::DEVICE_GPU).HostMemory("max_u").HostMemory("out_of_range_source_index"),
                        UpsampleSourceIndexesGpu);

// ---- ToneLatentBeamSearchDecode (tone_latent_beam_search_decode_op.cc) ----------------------------------------------------
class ToneLatentBeamSearchDecodeGpu : public tf::OpKernel {
public:
    explicit ToneLatentBeamSearchDecodeGpu(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
        OP_REQUIRES_OK(c, c->GetAttr("beam_width", &w_));
        OP_REQUIRES_OK(c, c->GetAttr("tone_class_size", &k_));
        OP_REQUIRES_OK(c, c->GetAttr("empty_tone_id", &empty_));
    }
    void Compute(tf::OpKernelContext* ctx) override {
        In h, lph, fin, t, u, il;
        if (!h.get(ctx, "h", 3) || !lph.get(ctx, "log_prob_history", 2) || !fin.get(ctx, "is_finished", 2) || !t.get(ctx, "t", 2) ||
            !u.get(ctx, "u", 2) || !il.get(ctx, "input_length", 1))
            return;
        const tf::int64 B = h.dim(0);
        SSNT_CHECK(ctx, h.dim(1) == w_ && lph.dim(1) == w_ && fin.dim(1) == w_ && t.dim(1) == w_ && u.dim(1) == w_, "Incompatible beam width");
        SSNT_CHECK(ctx, h.dim(2) == k_, "Incompatible tone class size");
        SSNT_CHECK(ctx, lph.dim(0) == B && fin.dim(0) == B && t.dim(0) == B && u.dim(0) == B && il.dim(0) == B, "Incompatible batch sizes");
        const tf::TensorShape s({B, (tf::int64)w_});
        int* pred = out<int>(ctx, "prediction", s);
        float* lp = out<float>(ctx, "log_prob", s);
        int* nt = out<int>(ctx, "next_t", s);
        int* nu = out<int>(ctx, "next_u", s);
        bool* nf = out<bool>(ctx, "next_is_finished", s);
        int* bb = out<int>(ctx, "beam_branch", s);
        if (!pred || !lp || !nt || !nu || !nf || !bb) return;
        bind_stream(ctx);
        ssnt_tts_fill_i32(pred, (size_t)(B * w_), empty_);  // _op.cc:168
        tone_latent_beam_search_decode(h.ptr<float>(), lph.ptr<float>(), fin.ptr<bool>(), t.ptr<int>(), u.ptr<int>(), il.ptr<int>(),
                                       (int)B, w_, k_, empty_, pred, lp, nt, nu, nf, bb);
    }
private:
    int w_ = 0, k_ = 0, empty_ = 0;
};
REGISTER_KERNEL_BUILDER(Name("ToneLatentBeamSearchDecode").Device(tf::DEVICE_GPU), ToneLatentBeamSearchDecodeGpu);

// ---- ToneLatentLevenshteinEditDistance (ssnt_tts_edit_distance.cc:24-78) -----------------------------------------------------
class LevenshteinEditDistanceGpu : public tf::OpKernel {
public:
    explicit LevenshteinEditDistanceGpu(tf::OpKernelConstruction* c) : tf::OpKernel(c) {}
    void Compute(tf::OpKernelContext* ctx) override {
        In a, b, al, bl;
        if (!a.get(ctx, "a", 2) || !b.get(ctx, "b", 2) || !al.get(ctx, "a_lengths", 1) || !bl.get(ctx, "b_lengths", 1)) return;
        SSNT_CHECK(ctx, a.dim(0) == b.dim(0) && a.dim(1) == b.dim(1), "a and b differ in shape");
        SSNT_CHECK(ctx, al.dim(0) == bl.dim(0) && al.dim(0) == a.dim(0), "Incompatible batch sizes");
        const tf::int64 B = a.dim(0);
        int* dist = out<int>(ctx, "distance", tf::TensorShape({B}));
        if (!dist) return;
        bind_stream(ctx);
        tone_latent_levenshtein_edit_distance(a.ptr<int>(), b.ptr<int>(), al.ptr<int>(), bl.ptr<int>(), (int)B, (int)a.dim(1), dist);
    }
};
REGISTER_KERNEL_BUILDER(Name("ToneLatentLevenshteinEditDistance").Device(tf::DEVICE_GPU), LevenshteinEditDistanceGpu);

// ---- new: the lattice loss and its gradients (no counterpart in the reference) ------------------------------------------------
REGISTER_OP("SSNTForwardBackward")
    .Input("log_emit: float32")
    .Input("log_shift: float32")
    .Input("t_len: int32")
    .Input("u_len: int32")
    .Output("log_likelihood: float32")
    .Output("loss: float32")
    .Output("grad_emit: float32")
    .Output("grad_shift: float32");
REGISTER_OP("SSNTForwardBackwardLogits")
    .Input("logits: float32")
    .Input("t_len: int32")
    .Input("u_len: int32")
    .Output("log_likelihood: float32")
    .Output("loss: float32")
    .Output("grad_logits: float32");

template <bool kLogits>
class ForwardBackwardGpu : public tf::OpKernel {
public:
    explicit ForwardBackwardGpu(tf::OpKernelConstruction* c) : tf::OpKernel(c) {}
    void Compute(tf::OpKernelContext* ctx) override {
        In x, y, tl, ul;
        if (!x.get(ctx, kLogits ? "logits" : "log_emit", 3) || (!kLogits && !y.get(ctx, "log_shift", 3)) || !tl.get(ctx, "t_len", 1) ||
            !ul.get(ctx, "u_len", 1))
            return;
        const tf::int64 B = x.dim(0), T = x.dim(1), U = x.dim(2);
        if (!kLogits) SSNT_CHECK(ctx, y.dim(0) == B && y.dim(1) == T && y.dim(2) == U, "log_emit and log_shift differ in shape");
        SSNT_CHECK(ctx, tl.dim(0) == B && ul.dim(0) == B, "Incompatible batch sizes");
        float* ll = out<float>(ctx, "log_likelihood", tf::TensorShape({B}));
        float* loss = out<float>(ctx, "loss", tf::TensorShape({}));
        float* g1 = out<float>(ctx, kLogits ? "grad_logits" : "grad_emit", tf::TensorShape({B, T, U}));
        float* g2 = kLogits ? nullptr : out<float>(ctx, "grad_shift", tf::TensorShape({B, T, U}));
        if (!ll || !loss || !g1 || (!kLogits && !g2)) return;
        // the workspace lives as long as the kernels: a temporary of the op's allocator (stream-ordered in TensorFlow)
        const size_t ws_bytes = kLogits ? ssnt_tts_forward_backward_logits_workspace_bytes((int)B, (int)T, (int)U)
                                        : ssnt_tts_forward_backward_workspace_bytes((int)B, (int)T, (int)U);
        tf::Tensor ws;
        OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({(tf::int64)ws_bytes}), &ws));
        bind_stream(ctx);
        if (kLogits)
            ssnt_tts_forward_backward_logits(x.ptr<float>(), tl.ptr<int>(), ul.ptr<int>(), (int)B, (int)T, (int)U, ll, loss, g1,
                                             ws.flat<unsigned char>().data(), ws_bytes);
        else
            ssnt_tts_forward_backward(x.ptr<float>(), y.ptr<float>(), tl.ptr<int>(), ul.ptr<int>(), (int)B, (int)T, (int)U, ll, loss,
                                      g1, g2, ws.flat<unsigned char>().data(), ws_bytes);
    }
};
REGISTER_KERNEL_BUILDER(Name("SSNTForwardBackward").Device(tf::DEVICE_GPU), ForwardBackwardGpu<false>);
REGISTER_KERNEL_BUILDER(Name("SSNTForwardBackwardLogits").Device(tf::DEVICE_GPU), ForwardBackwardGpu<true>);

}  // namespace ssnt_gpu
