// Shared helpers of the TensorFlow op shim over libpcops.so.
// The shim re-registers the reference's custom ops under identical names / attrs / shapes so that the reference's
// Python (tf_sampling.py, tf_grouping.py, tf_interpolate.py, pointnet_util.py, the models) runs unchanged.
#pragma once
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/framework/common_shape_fns.h"

#include "pcops.h"

namespace pcshim {
using namespace tensorflow;  // NOLINT
using shape_inference::InferenceContext;
using shape_inference::ShapeHandle;

// The CUDA stream TensorFlow runs this op on (the reference launches on the legacy default stream instead).
#ifndef PCSHIM_STREAM
#define PCSHIM_STREAM(ctx) ((pc_stream_t)(ctx)->eigen_device<Eigen::GpuDevice>().stream())
#endif

inline const float *F(const Tensor &t) { return t.flat<float>().data(); }
inline const int *I(const Tensor &t) { return t.flat<int>().data(); }
inline float *F(Tensor *t) { return t->flat<float>().data(); }
inline int *I(Tensor *t) { return t->flat<int>().data(); }
inline int dim(const Tensor &t, int i) { return static_cast<int>(t.shape().dim_size(i)); }

// Caller-owned scratch for the deterministic gradient ops: a temp int8 tensor of `bytes` bytes (may be 0).
inline Status scratch(OpKernelContext *ctx, size_t bytes, Tensor *t, void **ptr) {
  *ptr = nullptr;
  if (bytes == 0) return Status::OK();
  Status s = ctx->allocate_temp(DT_INT8, TensorShape({static_cast<long long>(bytes)}), t);
  if (s.ok()) *ptr = t->flat<int8>().data();
  return s;
}

#define PCSHIM_CHECK_RC(ctx, rc, what) \
  OP_REQUIRES(ctx, (rc) == PC_OK, errors::Internal(what, ": ", pc_error_string(rc)))
}  // namespace pcshim
