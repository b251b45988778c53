// tf_sampling_so.so: ProbSample, FarthestPointSample, GatherPoint, GatherPointGrad over libpcops.so.
// Registry (names, attrs, dtypes, shape functions, messages) restates the reference's
// pointnet2_tensorflow/tf_ops/sampling/tf_sampling.cpp:14-63,66-178; the kernels are pc_prob_sample / pc_fps /
// pc_gather_point / pc_gather_point_grad.
#include "shim_common.h"

namespace pcshim {

REGISTER_OP("ProbSample")
    .Input("inp: float32")
    .Input("inpr: float32")
    .Output("out: int32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle p, r;  // (batch, ncategory), (batch, npoints)
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 2, &p));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 2, &r));
      c->set_output(0, c->MakeShape({c->Dim(r, 0), c->Dim(r, 1)}));
      return Status::OK();
    });

REGISTER_OP("FarthestPointSample")
    .Attr("npoint: int")
    .Input("inp: float32")
    .Output("out: int32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle in;  // (batch, ndataset, 3)
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &in));
      int npoint;
      TF_RETURN_IF_ERROR(c->GetAttr("npoint", &npoint));
      c->set_output(0, c->MakeShape({c->Dim(in, 0), npoint}));
      return Status::OK();
    });

REGISTER_OP("GatherPoint")
    .Input("inp: float32")
    .Input("idx: int32")
    .Output("out: float32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle in, ix;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &in));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 2, &ix));
      c->set_output(0, c->MakeShape({c->Dim(in, 0), c->Dim(ix, 1), c->Dim(in, 2)}));
      return Status::OK();
    });

REGISTER_OP("GatherPointGrad")
    .Input("inp: float32")
    .Input("idx: int32")
    .Input("out_g: float32")
    .Output("inp_g: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->input(0));
      return Status::OK();
    });

class ProbSampleGpuOp : public OpKernel {
 public:
  explicit ProbSampleGpuOp(OpKernelConstruction *c) : OpKernel(c) {}
  void Compute(OpKernelContext *ctx) override {
    const Tensor &inp = ctx->input(0), &inpr = ctx->input(1);
    OP_REQUIRES(ctx, inp.dims() == 2, errors::InvalidArgument("ProbSample expects (batch_size,num_choices) inp shape"));
    const int b = dim(inp, 0), n = dim(inp, 1);
    OP_REQUIRES(ctx, inpr.dims() == 2 && dim(inpr, 0) == b,
                errors::InvalidArgument("ProbSample expects (batch_size,num_points) inpr shape"));
    const int m = dim(inpr, 1);
    Tensor *out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, m}, &out));
    Tensor temp;  // cumulative sums, tf_sampling.cpp:85-88
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_FLOAT, TensorShape{b, n}, &temp));
    PCSHIM_CHECK_RC(ctx, pc_prob_sample(b, n, m, F(inp), F(inpr), temp.flat<float>().data(), I(out), PCSHIM_STREAM(ctx)),
                    "pc_prob_sample");
  }
};
REGISTER_KERNEL_BUILDER(Name("ProbSample").Device(DEVICE_GPU), ProbSampleGpuOp);

class FarthestPointSampleGpuOp : public OpKernel {
 public:
  explicit FarthestPointSampleGpuOp(OpKernelConstruction *c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("npoint", &npoint_));
    OP_REQUIRES(c, npoint_ > 0, errors::InvalidArgument("FarthestPointSample expects positive npoint"));
  }
  void Compute(OpKernelContext *ctx) override {
    const Tensor &inp = ctx->input(0);
    OP_REQUIRES(ctx, inp.dims() == 3 && dim(inp, 2) == 3,
                errors::InvalidArgument("FarthestPointSample expects (batch_size,num_points,3) inp shape"));
    const int b = dim(inp, 0), n = dim(inp, 1), m = npoint_;
    Tensor *out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, m}, &out));
    Tensor ws_t;
    void *ws = nullptr;
    OP_REQUIRES_OK(ctx, scratch(ctx, pc_fps_workspace_bytes(b, n, m), &ws_t, &ws));
    PCSHIM_CHECK_RC(ctx, pc_fps(b, n, m, F(inp), ws, I(out), PCSHIM_STREAM(ctx)), "pc_fps");
  }

 private:
  int npoint_;
};
REGISTER_KERNEL_BUILDER(Name("FarthestPointSample").Device(DEVICE_GPU), FarthestPointSampleGpuOp);

class GatherPointGpuOp : public OpKernel {
 public:
  explicit GatherPointGpuOp(OpKernelConstruction *c) : OpKernel(c) {}
  void Compute(OpKernelContext *ctx) override {
    const Tensor &inp = ctx->input(0), &idx = ctx->input(1);
    OP_REQUIRES(ctx, inp.dims() == 3 && dim(inp, 2) == 3,
                errors::InvalidArgument("GatherPoint expects (batch_size,num_points,3) inp shape"));
    const int b = dim(inp, 0), n = dim(inp, 1);
    OP_REQUIRES(ctx, idx.dims() == 2 && dim(idx, 0) == b,
                errors::InvalidArgument("GatherPoint expects (batch_size,num_result) idx shape"));
    const int m = dim(idx, 1);
    Tensor *out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, m, 3}, &out));
    PCSHIM_CHECK_RC(ctx, pc_gather_point(b, n, m, F(inp), I(idx), F(out), PCSHIM_STREAM(ctx)), "pc_gather_point");
  }
};
REGISTER_KERNEL_BUILDER(Name("GatherPoint").Device(DEVICE_GPU), GatherPointGpuOp);

class GatherPointGradGpuOp : public OpKernel {
 public:
  explicit GatherPointGradGpuOp(OpKernelConstruction *c) : OpKernel(c) {}
  void Compute(OpKernelContext *ctx) override {
    const Tensor &inp = ctx->input(0), &idx = ctx->input(1), &out_g = ctx->input(2);
    OP_REQUIRES(ctx, inp.dims() == 3 && dim(inp, 2) == 3,
                errors::InvalidArgument("GatherPointGradGpuOp expects (batch_size,num_points,3) inp"));
    const int b = dim(inp, 0), n = dim(inp, 1);
    OP_REQUIRES(ctx, idx.dims() == 2 && dim(idx, 0) == b,
                errors::InvalidArgument("GatherPointGradGpuOp expects (batch_size,num_result) idx shape"));
    const int m = dim(idx, 1);
    OP_REQUIRES(ctx, out_g.dims() == 3 && dim(out_g, 0) == b && dim(out_g, 1) == m && dim(out_g, 2) == 3,
                errors::InvalidArgument("GatherPointGradGpuOp expects (batch_size,num_result,3) out_g shape"));
    Tensor *inp_g = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, n, 3}, &inp_g));
    Tensor ws_t;
    void *ws = nullptr;
    OP_REQUIRES_OK(ctx, scratch(ctx, pc_gather_point_grad_workspace_bytes(b, n, m), &ws_t, &ws));
    // no memset: pc_gather_point_grad overwrites every element (reference: cudaMemset + atomics, tf_sampling.cpp:174)
    PCSHIM_CHECK_RC(ctx, pc_gather_point_grad(b, n, m, F(out_g), I(idx), F(inp_g), ws, PCSHIM_STREAM(ctx)),
                    "pc_gather_point_grad");
  }
};
REGISTER_KERNEL_BUILDER(Name("GatherPointGrad").Device(DEVICE_GPU), GatherPointGradGpuOp);

}  // namespace pcshim
