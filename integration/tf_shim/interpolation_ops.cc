// tf_interpolate_so.so: ThreeNN, ThreeInterpolate, ThreeInterpolateGrad over libpcops.so.
// Registry restates pointnet2_tensorflow/tf_ops/interpolation_3d/tf_interpolate.cpp:12-46,157-262.  The reference
// registers these three ops for DEVICE_CPU only (:187,:222,:262); here they are DEVICE_GPU kernels, so TensorFlow no
// longer bounces the tensors through the host around every feature-propagation level.
#include "shim_common.h"

namespace pcshim {

REGISTER_OP("ThreeNN")
    .Input("xyz1: float32")
    .Input("xyz2: float32")
    .Output("dist: float32")
    .Output("idx: int32")
    .SetShapeFn([](InferenceContext *c) {  // the reference copies input 0's shape (b,n,3) to both outputs
      c->set_output(0, c->input(0));
      c->set_output(1, c->input(0));
      return Status::OK();
    });

REGISTER_OP("ThreeInterpolate")
    .Input("points: float32")
    .Input("idx: int32")
    .Input("weight: float32")
    .Output("out: float32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle pts, ix;  // (b,m,c), (b,n,3)
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &pts));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 3, &ix));
      c->set_output(0, c->MakeShape({c->Dim(pts, 0), c->Dim(ix, 1), c->Dim(pts, 2)}));
      return Status::OK();
    });

REGISTER_OP("ThreeInterpolateGrad")
    .Input("points: float32")
    .Input("idx: int32")
    .Input("weight: float32")
    .Input("grad_out: float32")
    .Output("grad_points: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->input(0));
      return Status::OK();
    });

class ThreeNNGpuOp : public OpKernel {
 public:
  explicit ThreeNNGpuOp(OpKernelConstruction *c) : OpKernel(c) {}
  void Compute(OpKernelContext *ctx) override {
    const Tensor &xyz1 = ctx->input(0), &xyz2 = ctx->input(1);
    OP_REQUIRES(ctx, xyz1.dims() == 3 && dim(xyz1, 2) == 3, errors::InvalidArgument("ThreeNN expects (b,n,3) xyz1 shape."));
    const int b = dim(xyz1, 0), n = dim(xyz1, 1);
    OP_REQUIRES(ctx, xyz2.dims() == 3 && dim(xyz2, 2) == 3, errors::InvalidArgument("ThreeNN expects (b,m,3) xyz2 shape."));
    const int m = dim(xyz2, 1);
    Tensor *dist = nullptr, *idx = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, n, 3}, &dist));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape{b, n, 3}, &idx));
    Tensor ws_t;  // per-scene cell grid over the known cloud (same outputs as pc_three_nn)
    void *ws = nullptr;
    OP_REQUIRES_OK(ctx, scratch(ctx, pc_three_nn_grid_workspace_bytes(b, n, m), &ws_t, &ws));
    PCSHIM_CHECK_RC(ctx, pc_three_nn_grid(b, n, m, F(xyz1), F(xyz2), F(dist), I(idx), ws, PCSHIM_STREAM(ctx)),
                    "pc_three_nn_grid");
  }
};
REGISTER_KERNEL_BUILDER(Name("ThreeNN").Device(DEVICE_GPU), ThreeNNGpuOp);

class ThreeInterpolateGpuOp : public OpKernel {
 public:
  explicit ThreeInterpolateGpuOp(OpKernelConstruction *c) : OpKernel(c) {}
  void Compute(OpKernelContext *ctx) override {
    const Tensor &points = ctx->input(0), &idx = ctx->input(1), &weight = ctx->input(2);
    OP_REQUIRES(ctx, points.dims() == 3, errors::InvalidArgument("ThreeInterpolate expects (b,m,c) points shape"));
    const int b = dim(points, 0), m = dim(points, 1), c = dim(points, 2);
    OP_REQUIRES(ctx, idx.dims() == 3 && dim(idx, 0) == b && dim(idx, 2) == 3,
                errors::InvalidArgument("ThreeInterpolate expects (b,n,3) idx shape"));
    const int n = dim(idx, 1);
    OP_REQUIRES(ctx, weight.dims() == 3 && dim(weight, 0) == b && dim(weight, 1) == n && dim(weight, 2) == 3,
                errors::InvalidArgument("ThreeInterpolate expects (b,n,3) weight shape"));
    Tensor *out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, n, c}, &out));
    PCSHIM_CHECK_RC(ctx, pc_three_interpolate(b, m, c, n, F(points), I(idx), F(weight), F(out), PCSHIM_STREAM(ctx)),
                    "pc_three_interpolate");
  }
};
REGISTER_KERNEL_BUILDER(Name("ThreeInterpolate").Device(DEVICE_GPU), ThreeInterpolateGpuOp);

class ThreeInterpolateGradGpuOp : public OpKernel {
 public:
  explicit ThreeInterpolateGradGpuOp(OpKernelConstruction *c) : OpKernel(c) {}
  void Compute(OpKernelContext *ctx) override {
    const Tensor &points = ctx->input(0), &idx = ctx->input(1), &weight = ctx->input(2), &grad_out = ctx->input(3);
    OP_REQUIRES(ctx, points.dims() == 3, errors::InvalidArgument("ThreeInterpolateGrad expects (b,m,c) points shape"));
    const int b = dim(points, 0), m = dim(points, 1), c = dim(points, 2);
    OP_REQUIRES(ctx, idx.dims() == 3 && dim(idx, 0) == b,
                errors::InvalidArgument("ThreeInterpolateGrad expects (b,n,3) idx shape"));
    const int n = dim(idx, 1);
    OP_REQUIRES(ctx, weight.dims() == 3 && dim(weight, 0) == b && dim(weight, 1) == n && dim(weight, 2) == 3,
                errors::InvalidArgument("ThreeInterpolateGrad expects (b,n,3) weight shape"));
    OP_REQUIRES(ctx, grad_out.dims() == 3 && dim(grad_out, 0) == b && dim(grad_out, 1) == n && dim(grad_out, 2) == c,
                errors::InvalidArgument("ThreeInterpolateGrad expects (b,n,c) grad_out shape"));
    Tensor *grad_points = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, m, c}, &grad_points));
    Tensor ws_t;
    void *ws = nullptr;
    OP_REQUIRES_OK(ctx, scratch(ctx, pc_three_interpolate_grad_workspace_bytes(b, n, c, m), &ws_t, &ws));
    PCSHIM_CHECK_RC(ctx,
                    pc_three_interpolate_grad(b, n, c, m, F(grad_out), I(idx), F(weight), F(grad_points), ws,
                                              PCSHIM_STREAM(ctx)),
                    "pc_three_interpolate_grad");
  }
};
REGISTER_KERNEL_BUILDER(Name("ThreeInterpolateGrad").Device(DEVICE_GPU), ThreeInterpolateGradGpuOp);

}  // namespace pcshim
