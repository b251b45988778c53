// Stand-in declarations of the TensorFlow C++ API surface the shim uses, ONLY so that `check.sh` can run
// `g++ -fsyntax-only` on the shim where TensorFlow is not installed.  Declarations only; nothing here runs.
#pragma once
#include <cstddef>
#include <initializer_list>
#include <vector>
namespace Eigen { struct GpuDevice { void *stream() const; }; }
namespace tensorflow {
typedef signed char int8;
enum DataType { DT_INT8, DT_FLOAT, DT_INT32 };
struct Status { static Status OK(); bool ok() const; };
namespace errors {
template <class... A> Status InvalidArgument(A...);
template <class... A> Status Internal(A...);
}
namespace shape_inference {
struct ShapeHandle {};
struct DimensionHandle { DimensionHandle(); DimensionHandle(int); };
struct InferenceContext {
  ShapeHandle input(int);
  void set_output(int, ShapeHandle);
  Status WithRank(ShapeHandle, int, ShapeHandle *);
  DimensionHandle Dim(ShapeHandle, int);
  ShapeHandle MakeShape(std::initializer_list<DimensionHandle>);
  template <class T> Status GetAttr(const char *, T *);
};
}
struct OpDefBuilderStub {
  explicit OpDefBuilderStub(const char *);
  OpDefBuilderStub &Input(const char *);
  OpDefBuilderStub &Output(const char *);
  OpDefBuilderStub &Attr(const char *);
  template <class F> OpDefBuilderStub &SetShapeFn(F);
};
struct TensorShape {
  TensorShape();
  TensorShape(std::initializer_list<long long>);
  long long dim_size(int) const;
  int dims() const;
};
template <class T> struct FlatStub { T *data() const; };
struct Tensor {
  int dims() const;
  const TensorShape &shape() const;
  template <class T> FlatStub<T> flat() const;
};
struct OpKernelConstruction { template <class T> Status GetAttr(const char *, T *); };
struct OpKernelContext {
  const Tensor &input(int);
  Status allocate_output(int, TensorShape, Tensor **);
  Status allocate_temp(DataType, TensorShape, Tensor *);
  template <class D> const D &eigen_device() const;
};
struct OpKernel { explicit OpKernel(OpKernelConstruction *); virtual void Compute(OpKernelContext *) = 0; virtual ~OpKernel(); };
struct KernelDefBuilderStub { KernelDefBuilderStub &Device(const char *); };
KernelDefBuilderStub Name(const char *);
}  // namespace tensorflow
#define TFS_CAT2(a, b) a##b
#define TFS_CAT(a, b) TFS_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::OpDefBuilderStub TFS_CAT(tfs_op_, __COUNTER__) = ::tensorflow::OpDefBuilderStub(name)
#define DEVICE_CPU "CPU"
#define DEVICE_GPU "GPU"
#define REGISTER_KERNEL_BUILDER(builder, cls) static int TFS_CAT(tfs_k_, __COUNTER__) = ((void)(builder), (int)sizeof(cls))
#define OP_REQUIRES(ctx, cond, status) do { if (!(cond)) { (void)(status); return; } } while (0)
#define OP_REQUIRES_OK(ctx, status) do { if (!(status).ok()) return; } while (0)
#define TF_RETURN_IF_ERROR(expr) do { ::tensorflow::Status s__ = (expr); if (!s__.ok()) return s__; } while (0)
