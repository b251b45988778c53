// Minimal stand-ins for the TensorFlow C++ headers that
// pointnet2_tensorflow/tf_ops/interpolation_3d/tf_interpolate.cpp includes, so that the reference
// translation unit compiles UNMODIFIED without TensorFlow and its plain functions
// (threenn_cpu, threeinterpolate_cpu, threeinterpolate_grad_cpu; tf_interpolate.cpp:60-153)
// can be called as oracle/_ref.  Nothing here computes anything: op registration and the
// OpKernel classes compile to dead code.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <initializer_list>
#include <vector>

namespace tensorflow {

struct Status {
  static Status OK() { return Status(); }
  bool ok() const { return true; }
};

namespace errors {
template <class... A> inline Status InvalidArgument(A...) { return Status(); }
}  // namespace errors

namespace shape_inference {
struct ShapeHandle {};
struct DimensionHandle {};
struct InferenceContext {
  ShapeHandle input(int) { return ShapeHandle(); }
  void set_output(int, ShapeHandle) {}
  Status WithRank(ShapeHandle, int, ShapeHandle *) { return Status(); }
  DimensionHandle Dim(ShapeHandle, int) { return DimensionHandle(); }
  ShapeHandle MakeShape(std::initializer_list<DimensionHandle>) { return ShapeHandle(); }
};
}  // namespace shape_inference

struct OpDefBuilderStub {
  explicit OpDefBuilderStub(const char *) {}
  OpDefBuilderStub &Input(const char *) { return *this; }
  OpDefBuilderStub &Output(const char *) { return *this; }
  OpDefBuilderStub &Attr(const char *) { return *this; }
  template <class F> OpDefBuilderStub &SetShapeFn(F) { return *this; }
};

struct TensorShape {
  std::vector<long long> d;
  TensorShape() {}
  TensorShape(std::initializer_list<int> l) : d(l.begin(), l.end()) {}
  long long dim_size(int i) const { return d[i]; }
  int dims() const { return (int)d.size(); }
};

template <class T> struct FlatStub {
  T *p;
  T &operator()(size_t i) const { return p[i]; }
};

struct Tensor {
  TensorShape s;
  void *buf = nullptr;
  int dims() const { return s.dims(); }
  const TensorShape &shape() const { return s; }
  template <class T> FlatStub<T> flat() const { return FlatStub<T>{static_cast<T *>(buf)}; }
};

struct OpKernelConstruction {
  template <class T> Status GetAttr(const char *, T *) { return Status(); }
};

struct OpKernelContext {
  Tensor dummy;
  const Tensor &input(int) { return dummy; }
  Status allocate_output(int, TensorShape, Tensor **t) { *t = &dummy; return Status(); }
};

struct OpKernel {
  explicit OpKernel(OpKernelConstruction *) {}
  virtual void Compute(OpKernelContext *) = 0;
  virtual ~OpKernel() {}
};

struct KernelDefBuilderStub {
  KernelDefBuilderStub &Device(const char *) { return *this; }
};
inline KernelDefBuilderStub Name(const char *) { return KernelDefBuilderStub(); }

}  // namespace tensorflow

#define TFSTUB_CAT2(a, b) a##b
#define TFSTUB_CAT(a, b) TFSTUB_CAT2(a, b)
#define REGISTER_OP(name) \
  static ::tensorflow::OpDefBuilderStub TFSTUB_CAT(tfstub_op_, __COUNTER__) = ::tensorflow::OpDefBuilderStub(name)
#define DEVICE_CPU "CPU"
#define DEVICE_GPU "GPU"
#define REGISTER_KERNEL_BUILDER(builder, cls) \
  static int TFSTUB_CAT(tfstub_k_, __COUNTER__) = ((void)(builder), (int)sizeof(cls))
#define OP_REQUIRES(ctx, cond, status) \
  do { if (!(cond)) { (void)(status); return; } } while (0)
#define OP_REQUIRES_OK(ctx, status) \
  do { (void)(status); } while (0)
