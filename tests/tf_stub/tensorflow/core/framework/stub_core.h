// Minimal FUNCTIONAL stand-in for the three TensorFlow headers tf_shim/dispu_tf_ops.cc includes (op.h, op_kernel.h,
// shape_inference.h).  Test infrastructure only: TensorFlow itself is not installable in the build container, so the
// shim is compiled against this stub to check that it is valid C++ against the API surface it uses, that it registers
// the reference's op signatures, and — on the GPU box — that each OpKernel::Compute validates shapes, allocates the
// right outputs and forwards to the C ABI correctly (tests/test_tf_shim.py drives it through stub_runtime.cc).
// It mirrors only the members the shim touches; semantics follow TF 1.x (Status, OP_REQUIRES, REGISTER_OP builder,
// REGISTER_KERNEL_BUILDER, InferenceContext).
#pragma once
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace tensorflow {

// ---- Status / errors -------------------------------------------------------------------------------
class Status {
 public:
  Status() : ok_(true) {}
  Status(int code, std::string msg) : ok_(false), code_(code), msg_(std::move(msg)) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string &error_message() const { return msg_; }
  int code() const { return code_; }
 private:
  bool ok_;
  int code_ = 0;
  std::string msg_;
};
namespace errors {
template <typename... A>
inline std::string StrCat(const A &...a) { std::ostringstream os; using x = int[]; (void)x{0, ((void)(os << a), 0)...}; return os.str(); }
template <typename... A> inline Status InvalidArgument(const A &...a) { return Status(3, StrCat(a...)); }
template <typename... A> inline Status Internal(const A &...a) { return Status(13, StrCat(a...)); }
}  // namespace errors
#define TF_RETURN_IF_ERROR(expr) do { ::tensorflow::Status _s = (expr); if (!_s.ok()) return _s; } while (0)

// ---- tensors ---------------------------------------------------------------------------------------
enum DataType { DT_FLOAT = 1, DT_INT32 = 3 };
class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64_t> d) : dims_(d) {}
  explicit TensorShape(std::vector<int64_t> d) : dims_(std::move(d)) {}
  int dims() const { return (int)dims_.size(); }
  int64_t dim_size(int i) const { return dims_[i]; }
  int64_t num_elements() const { int64_t n = 1; for (auto d : dims_) n *= d; return n; }
  bool operator==(const TensorShape &o) const { return dims_ == o.dims_; }
  const std::vector<int64_t> &vec() const { return dims_; }
 private:
  std::vector<int64_t> dims_;
};
template <typename T> struct FlatView { T *p; T *data() const { return p; } };
class Tensor {
 public:
  Tensor() {}
  Tensor(DataType dt, TensorShape s, void *p) : dt_(dt), shape_(std::move(s)), p_(p) {}
  int dims() const { return shape_.dims(); }
  int64_t dim_size(int i) const { return shape_.dim_size(i); }
  const TensorShape &shape() const { return shape_; }
  DataType dtype() const { return dt_; }
  template <typename T> FlatView<T> flat() const { return FlatView<T>{static_cast<T *>(p_)}; }
 private:
  DataType dt_ = DT_FLOAT;
  TensorShape shape_;
  void *p_ = nullptr;
};

// ---- shape inference -------------------------------------------------------------------------------
namespace shape_inference {
struct DimensionHandle { int64_t v = -1; DimensionHandle() {} DimensionHandle(int64_t x) : v(x) {} };
struct ShapeHandle { std::vector<int64_t> dims; bool known = false; };
class InferenceContext {
 public:
  std::vector<ShapeHandle> inputs, outputs;
  std::map<std::string, int> attrs;
  ShapeHandle input(int i) const { return inputs.at(i); }
  Status WithRank(const ShapeHandle &s, int rank, ShapeHandle *out) {
    if ((int)s.dims.size() != rank) return errors::InvalidArgument("Shape must be rank ", rank, " but is rank ", s.dims.size());
    *out = s; return Status::OK();
  }
  Status GetAttr(const std::string &name, int *v) {
    auto it = attrs.find(name);
    if (it == attrs.end()) return errors::InvalidArgument("no attr ", name);
    *v = it->second; return Status::OK();
  }
  DimensionHandle Dim(const ShapeHandle &s, int i) { return DimensionHandle(s.dims.at(i)); }
  ShapeHandle MakeShape(std::initializer_list<DimensionHandle> d) { ShapeHandle s; s.known = true; for (auto &x : d) s.dims.push_back(x.v); return s; }
  void set_output(int i, const ShapeHandle &s) { if ((int)outputs.size() <= i) outputs.resize(i + 1); outputs[i] = s; outputs[i].known = true; }
};
}  // namespace shape_inference

// ---- op registry -----------------------------------------------------------------------------------
struct StubOpDef {
  std::string name;
  std::vector<std::string> attrs, inputs, outputs;
  std::function<Status(shape_inference::InferenceContext *)> shape_fn;
};
std::vector<StubOpDef> &StubOpRegistry();
class OpDefBuilderWrapper {
 public:
  explicit OpDefBuilderWrapper(const char *name) { def_.name = name; }
  OpDefBuilderWrapper &Attr(std::string s) { def_.attrs.push_back(std::move(s)); return *this; }
  OpDefBuilderWrapper &Input(std::string s) { def_.inputs.push_back(std::move(s)); return *this; }
  OpDefBuilderWrapper &Output(std::string s) { def_.outputs.push_back(std::move(s)); return *this; }
  OpDefBuilderWrapper &SetShapeFn(std::function<Status(shape_inference::InferenceContext *)> f) { def_.shape_fn = std::move(f); return *this; }
  const StubOpDef &def() const { return def_; }
 private:
  StubOpDef def_;
};
struct OpRegistrar { OpRegistrar(const OpDefBuilderWrapper &w) { StubOpRegistry().push_back(w.def()); } };
#define TF_STUB_CAT2(a, b) a##b
#define TF_STUB_CAT(a, b) TF_STUB_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::OpRegistrar TF_STUB_CAT(_tf_stub_op_, __COUNTER__) = ::tensorflow::OpDefBuilderWrapper(name)

// ---- kernels ---------------------------------------------------------------------------------------
class OpKernelConstruction {
 public:
  std::map<std::string, int> attrs;
  Status status;
  Status GetAttr(const std::string &name, int *v) {
    auto it = attrs.find(name);
    if (it == attrs.end()) return errors::InvalidArgument("no attr ", name);
    *v = it->second; return Status::OK();
  }
  void CtxFailure(const Status &s) { status = s; }
};
// allocation goes back to the test driver: (output index or -1 for a temp, dtype, rank, dims) -> device pointer
using StubAlloc = void *(*)(int index, int dtype, int rank, const long long *dims);
class OpKernelContext {
 public:
  std::vector<Tensor> inputs;
  std::map<int, Tensor> outputs;
  std::vector<std::unique_ptr<Tensor>> owned;
  StubAlloc alloc = nullptr;
  Status status;
  std::vector<DataType> out_types;
  const Tensor &input(int i) const { return inputs.at(i); }
  Status allocate_output(int i, const TensorShape &s, Tensor **out) {
    std::vector<long long> d(s.vec().begin(), s.vec().end());
    DataType dt = i < (int)out_types.size() ? out_types[i] : DT_FLOAT;
    void *p = alloc(i, (int)dt, s.dims(), d.data());
    if (!p && s.num_elements() > 0) return errors::Internal("allocate_output failed");
    owned.emplace_back(new Tensor(dt, s, p));
    outputs[i] = *owned.back();
    *out = owned.back().get();
    return Status::OK();
  }
  Status allocate_temp(DataType dt, const TensorShape &s, Tensor *out) {
    std::vector<long long> d(s.vec().begin(), s.vec().end());
    void *p = alloc(-1, (int)dt, s.dims(), d.data());
    if (!p && s.num_elements() > 0) return errors::Internal("allocate_temp failed");
    *out = Tensor(dt, s, p);
    return Status::OK();
  }
  void CtxFailure(const Status &s) { status = s; }
};
class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction *) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext *ctx) = 0;
};
#define OP_REQUIRES(CTX, EXP, STATUS) do { if (!(EXP)) { (CTX)->CtxFailure((STATUS)); return; } } while (0)
#define OP_REQUIRES_OK(CTX, ...) do { ::tensorflow::Status _s(__VA_ARGS__); if (!_s.ok()) { (CTX)->CtxFailure(_s); return; } } while (0)

constexpr const char *DEVICE_GPU = "GPU";
constexpr const char *DEVICE_CPU = "CPU";
struct KernelDef { std::string op, device; };
class Name {
 public:
  explicit Name(const char *op) { def_.op = op; }
  Name &Device(const char *d) { def_.device = d; return *this; }
  const KernelDef &def() const { return def_; }
 private:
  KernelDef def_;
};
struct StubKernelEntry { KernelDef def; std::function<OpKernel *(OpKernelConstruction *)> make; };
std::vector<StubKernelEntry> &StubKernelRegistry();
struct KernelRegistrar {
  KernelRegistrar(const Name &n, std::function<OpKernel *(OpKernelConstruction *)> f) { StubKernelRegistry().push_back({n.def(), std::move(f)}); }
};
#define REGISTER_KERNEL_BUILDER(kernel_builder, ...)                                                                  \
  static ::tensorflow::KernelRegistrar TF_STUB_CAT(_tf_stub_kernel_, __COUNTER__)(                                   \
      ::tensorflow::kernel_builder, [](::tensorflow::OpKernelConstruction *c) -> ::tensorflow::OpKernel * { return new __VA_ARGS__(c); })

}  // namespace tensorflow
