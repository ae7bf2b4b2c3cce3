// Minimal stand-in for the few TensorFlow types integration/tf_ops_gpu.cc touches, so that the file can be
// syntax- and type-checked (g++ -fsyntax-only) in an image without TensorFlow.  NOT TensorFlow: shapes of the real
// API only (tensorflow/core/framework/op_kernel.h, tensor.h, tensor_shape.h, register_types.h).
#pragma once
#include <cstdint>
#include <initializer_list>
#include <string>
#include <vector>

typedef struct CUstream_st* cudaStream_t;

namespace tensorflow {

typedef long long int64;
constexpr const char* DEVICE_GPU = "GPU";
constexpr const char* DEVICE_CPU = "CPU";

class Status {
public:
    Status() = default;
    explicit Status(std::string m) : msg_(std::move(m)), ok_(false) {}
    bool ok() const { return ok_; }
private:
    std::string msg_;
    bool ok_ = true;
};
namespace errors {
template <typename... A>
Status InvalidArgument(A...) { return Status("invalid argument"); }
}  // namespace errors

class TensorShape {
public:
    TensorShape() = default;
    TensorShape(std::initializer_list<int64> d) : d_(d) {}
    int dims() const { return (int)d_.size(); }
    int64 dim_size(int i) const { return d_[(size_t)i]; }
    int64 num_elements() const { int64 n = 1; for (auto x : d_) n *= x; return n; }
private:
    std::vector<int64> d_;
};

template <typename T>
struct FlatView {
    T* p; int64 n;
    T* data() const { return p; }
    int64 size() const { return n; }
};
template <typename T>
struct ScalarView {
    T* p;
    T& operator()() const { return *p; }
};

class Tensor {
public:
    const TensorShape& shape() const { return shape_; }
    int dims() const { return shape_.dims(); }
    int64 dim_size(int i) const { return shape_.dim_size(i); }
    int64 NumElements() const { return shape_.num_elements(); }
    template <typename T> FlatView<T> flat() { return {static_cast<T*>(buf_), NumElements()}; }
    template <typename T> FlatView<const T> flat() const { return {static_cast<const T*>(buf_), NumElements()}; }
    template <typename T> ScalarView<const T> scalar() const { return {static_cast<const T*>(buf_)}; }
private:
    TensorShape shape_;
    void* buf_ = nullptr;
};

struct GpuDevice {
    cudaStream_t stream() const { return nullptr; }
};

class OpKernelConstruction {
public:
    template <typename T> Status GetAttr(const char*, T*) const { return Status(); }
    void CtxFailure(const Status&) {}
    void CtxFailureWithWarning(const Status&) {}
};

class OpKernelContext {
public:
    Status input(const char*, const Tensor**) { return Status(); }
    Status allocate_output(const char*, const TensorShape&, Tensor**) { return Status(); }
    Status allocate_temp(int /*dtype*/, const TensorShape&, Tensor*) { return Status(); }
    const GpuDevice& eigen_gpu_device() const { return dev_; }
    void CtxFailure(const Status&) {}
    void CtxFailureWithWarning(const Status&) {}
private:
    GpuDevice dev_;
};

class OpKernel {
public:
    explicit OpKernel(OpKernelConstruction*) {}
    virtual ~OpKernel() = default;
    virtual void Compute(OpKernelContext* ctx) = 0;
};

enum { DT_UINT8 = 4 };

namespace register_kernel {
struct Name {
    explicit Name(const char*) {}
    Name& Device(const char*) { return *this; }
    Name& HostMemory(const char*) { return *this; }
};
struct Registrar {
    template <typename F> Registrar(const Name&, F) {}
};
}  // namespace register_kernel

struct OpDefBuilderStub {
    explicit OpDefBuilderStub(const char*) {}
    OpDefBuilderStub& Input(const char*) { return *this; }
    OpDefBuilderStub& Output(const char*) { return *this; }
    OpDefBuilderStub& Attr(const char*) { return *this; }
};

}  // namespace tensorflow

#define SSNT_TF_CAT2(a, b) a##b
#define SSNT_TF_CAT(a, b) SSNT_TF_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::OpDefBuilderStub SSNT_TF_CAT(ssnt_op_def_, __COUNTER__) = ::tensorflow::OpDefBuilderStub(name)
#define REGISTER_KERNEL_BUILDER(builder, ...)                                                             \
    static ::tensorflow::register_kernel::Registrar SSNT_TF_CAT(ssnt_kernel_reg_, __COUNTER__)(            \
        ::tensorflow::register_kernel::builder, [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(c); })
#define OP_REQUIRES(ctx, cond, status)      \
    do {                                    \
        if (!(cond)) {                      \
            (ctx)->CtxFailure((status));    \
            return;                         \
        }                                   \
    } while (0)
#define OP_REQUIRES_OK(ctx, expr)           \
    do {                                    \
        ::tensorflow::Status s_ = (expr);   \
        if (!s_.ok()) {                     \
            (ctx)->CtxFailureWithWarning(s_); \
            return;                         \
        }                                   \
    } while (0)
