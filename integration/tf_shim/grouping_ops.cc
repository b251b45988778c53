// tf_grouping_so.so: QueryBallPoint, SelectionSort, GroupPoint, GroupPointGrad over libpcops.so.
// Registry restates pointnet2_tensorflow/tf_ops/grouping/tf_grouping.cpp:13-63,67-208.
#include "shim_common.h"

namespace pcshim {

REGISTER_OP("QueryBallPoint")
    .Attr("radius: float")
    .Attr("nsample: int")
    .Input("xyz1: float32")
    .Input("xyz2: float32")
    .Output("idx: int32")
    .Output("pts_cnt: int32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle q;  // (batch, npoint, 3)
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 3, &q));
      int nsample;
      TF_RETURN_IF_ERROR(c->GetAttr("nsample", &nsample));
      c->set_output(0, c->MakeShape({c->Dim(q, 0), c->Dim(q, 1), nsample}));
      c->set_output(1, c->MakeShape({c->Dim(q, 0), c->Dim(q, 1)}));
      return Status::OK();
    });

REGISTER_OP("SelectionSort")
    .Attr("k: int")
    .Input("dist: float32")
    .Output("outi: int32")
    .Output("out: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->input(0));
      c->set_output(1, c->input(0));
      return Status::OK();
    });

REGISTER_OP("GroupPoint")
    .Input("points: float32")
    .Input("idx: int32")
    .Output("out: float32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle pts, ix;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &pts));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 3, &ix));
      c->set_output(0, c->MakeShape({c->Dim(ix, 0), c->Dim(ix, 1), c->Dim(ix, 2), c->Dim(pts, 2)}));
      return Status::OK();
    });

REGISTER_OP("GroupPointGrad")
    .Input("points: float32")
    .Input("idx: int32")
    .Input("grad_out: float32")
    .Output("grad_points: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->input(0));
      return Status::OK();
    });

class QueryBallPointGpuOp : public OpKernel {
 public:
  explicit QueryBallPointGpuOp(OpKernelConstruction *c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("radius", &radius_));
    OP_REQUIRES(c, radius_ > 0, errors::InvalidArgument("QueryBallPoint expects positive radius"));
    OP_REQUIRES_OK(c, c->GetAttr("nsample", &nsample_));
    OP_REQUIRES(c, nsample_ > 0, errors::InvalidArgument("QueryBallPoint expects positive nsample"));
  }
  void Compute(OpKernelContext *ctx) override {
    const Tensor &xyz1 = ctx->input(0), &xyz2 = ctx->input(1);
    OP_REQUIRES(ctx, xyz1.dims() == 3 && dim(xyz1, 2) == 3,
                errors::InvalidArgument("QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape."));
    const int b = dim(xyz1, 0), n = dim(xyz1, 1);
    OP_REQUIRES(ctx, xyz2.dims() == 3 && dim(xyz2, 2) == 3,
                errors::InvalidArgument("QueryBallPoint expects (batch_size, npoint, 3) xyz2 shape."));
    const int m = dim(xyz2, 1);
    Tensor *idx = nullptr, *cnt = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, m, nsample_}, &idx));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape{b, m}, &cnt));
    Tensor ws_t;  // per-scene cell grid (same outputs as pc_query_ball, far fewer pair tests)
    void *ws = nullptr;
    OP_REQUIRES_OK(ctx, scratch(ctx, pc_query_ball_grid_workspace_bytes(b, n, m), &ws_t, &ws));
    PCSHIM_CHECK_RC(ctx,
                    pc_query_ball_grid(b, n, m, radius_, nsample_, F(xyz1), F(xyz2), I(idx), I(cnt), ws, PCSHIM_STREAM(ctx)),
                    "pc_query_ball_grid");
  }

 private:
  float radius_;
  int nsample_;
};
REGISTER_KERNEL_BUILDER(Name("QueryBallPoint").Device(DEVICE_GPU), QueryBallPointGpuOp);

class SelectionSortGpuOp : public OpKernel {
 public:
  explicit SelectionSortGpuOp(OpKernelConstruction *c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("k", &k_));
    OP_REQUIRES(c, k_ > 0, errors::InvalidArgument("SelectionSort expects positive k"));
  }
  void Compute(OpKernelContext *ctx) override {
    const Tensor &dist = ctx->input(0);
    OP_REQUIRES(ctx, dist.dims() == 3, errors::InvalidArgument("SelectionSort expects (b,m,n) dist shape."));
    const int b = dim(dist, 0), m = dim(dist, 1), n = dim(dist, 2);
    Tensor *outi = nullptr, *out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, m, n}, &outi));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape{b, m, n}, &out));
    PCSHIM_CHECK_RC(ctx, pc_selection_sort(b, n, m, k_, F(dist), I(outi), F(out), PCSHIM_STREAM(ctx)),
                    "pc_selection_sort");
  }

 private:
  int k_;
};
REGISTER_KERNEL_BUILDER(Name("SelectionSort").Device(DEVICE_GPU), SelectionSortGpuOp);

class GroupPointGpuOp : public OpKernel {
 public:
  explicit GroupPointGpuOp(OpKernelConstruction *c) : OpKernel(c) {}
  void Compute(OpKernelContext *ctx) override {
    const Tensor &points = ctx->input(0), &idx = ctx->input(1);
    OP_REQUIRES(ctx, points.dims() == 3,
                errors::InvalidArgument("GroupPoint expects (batch_size, num_points, channel) points shape"));
    const int b = dim(points, 0), n = dim(points, 1), c = dim(points, 2);
    OP_REQUIRES(ctx, idx.dims() == 3 && dim(idx, 0) == b,
                errors::InvalidArgument("GroupPoint expects (batch_size, npoints, nsample) idx shape"));
    const int m = dim(idx, 1), nsample = dim(idx, 2);
    Tensor *out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, m, nsample, c}, &out));
    PCSHIM_CHECK_RC(ctx, pc_group_point(b, n, c, m, nsample, F(points), I(idx), F(out), PCSHIM_STREAM(ctx)),
                    "pc_group_point");
  }
};
REGISTER_KERNEL_BUILDER(Name("GroupPoint").Device(DEVICE_GPU), GroupPointGpuOp);

class GroupPointGradGpuOp : public OpKernel {
 public:
  explicit GroupPointGradGpuOp(OpKernelConstruction *c) : OpKernel(c) {}
  void Compute(OpKernelContext *ctx) override {
    const Tensor &points = ctx->input(0), &idx = ctx->input(1), &grad_out = ctx->input(2);
    OP_REQUIRES(ctx, points.dims() == 3,
                errors::InvalidArgument("GroupPointGrad expects (batch_size, num_points, channel) points shape"));
    const int b = dim(points, 0), n = dim(points, 1), c = dim(points, 2);
    OP_REQUIRES(ctx, idx.dims() == 3 && dim(idx, 0) == b,
                errors::InvalidArgument("GroupPointGrad expects (batch_size, npoints, nsample) idx shape"));
    const int m = dim(idx, 1), nsample = dim(idx, 2);
    OP_REQUIRES(ctx,
                grad_out.dims() == 4 && dim(grad_out, 0) == b && dim(grad_out, 1) == m && dim(grad_out, 2) == nsample &&
                    dim(grad_out, 3) == c,
                errors::InvalidArgument("GroupPointGrad expects (batch_size, npoints, nsample, channel) grad_out shape"));
    Tensor *grad_points = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, n, c}, &grad_points));
    Tensor ws_t;
    void *ws = nullptr;
    OP_REQUIRES_OK(ctx, scratch(ctx, pc_group_point_grad_workspace_bytes(b, n, c, m, nsample), &ws_t, &ws));
    PCSHIM_CHECK_RC(ctx,
                    pc_group_point_grad(b, n, c, m, nsample, F(grad_out), I(idx), F(grad_points), ws, PCSHIM_STREAM(ctx)),
                    "pc_group_point_grad");
  }
};
REGISTER_KERNEL_BUILDER(Name("GroupPointGrad").Device(DEVICE_GPU), GroupPointGradGpuOp);

}  // namespace pcshim
