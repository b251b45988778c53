// tf_attention_so.so: the per-neighbourhood attention of attention_points/attention_scannet/attention_layer.py as
// TensorFlow ops over libpcops.so.  These ops do NOT exist in the reference (its AttentionLayer.call composes stock TF
// ops: three Dense, reshape, matmul, softmax, matmul -- attention_layer.py:29-45); they are what a maintainer calls from
// that method instead:
//   PointAttentionContract(q, k, v){heads, key_dim} -> out           the reshape / matmul / softmax / matmul tail (:35-42)
//   PointAttentionContractGrad(q, k, v, grad_out)  -> dq, dk, dv     its gradient (registered in Python)
//   PointAttentionLayer(query, inp, wq, bq, wk, bk, wv, bv) -> out   the whole call (:29-45) on the tcgen05 tensor cores,
//                                                                    inference only, nsample = 32, C in {64,128,256,512}
#include "shim_common.h"

namespace pcshim {

REGISTER_OP("PointAttentionContract")
    .Attr("heads: int")
    .Attr("key_dim: int")
    .Input("q: float32")   // (b, np, heads*key_dim)
    .Input("k: float32")   // (b, np, nsample, heads*key_dim)
    .Input("v: float32")   // (b, np, nsample, heads*key_dim)
    .Output("out: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->input(0));
      return Status::OK();
    });

REGISTER_OP("PointAttentionContractGrad")
    .Attr("heads: int")
    .Attr("key_dim: int")
    .Input("q: float32")
    .Input("k: float32")
    .Input("v: float32")
    .Input("grad_out: float32")
    .Output("dq: float32")
    .Output("dk: float32")
    .Output("dv: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->input(0));
      c->set_output(1, c->input(1));
      c->set_output(2, c->input(2));
      return Status::OK();
    });

REGISTER_OP("PointAttentionLayer")
    .Input("query: float32")  // (b, np, 1, C): sample 0 of each group (attention_layer.py:259)
    .Input("inp: float32")    // (b, np, nsample, C)
    .Input("wq: float32")     // Dense kernels (C, C) laid out [in][out] as Keras stores them, biases (C)
    .Input("bq: float32")
    .Input("wk: float32")
    .Input("bk: float32")
    .Input("wv: float32")
    .Input("bv: float32")
    .Output("out: float32")   // (b, np, C)
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle in;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 4, &in));
      c->set_output(0, c->MakeShape({c->Dim(in, 0), c->Dim(in, 1), c->Dim(in, 3)}));
      return Status::OK();
    });

class AttentionAttrs : public OpKernel {
 public:
  explicit AttentionAttrs(OpKernelConstruction *c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("heads", &heads_));
    OP_REQUIRES_OK(c, c->GetAttr("key_dim", &key_dim_));
    OP_REQUIRES(c, heads_ > 0 && key_dim_ > 0, errors::InvalidArgument("PointAttentionContract expects positive heads, key_dim"));
  }

 protected:
  static bool same_shape(const Tensor &a, const Tensor &b) {
    if (a.dims() != b.dims()) return false;
    for (int i = 0; i < a.dims(); ++i)
      if (dim(a, i) != dim(b, i)) return false;
    return true;
  }
  // q (b,np,HD), k and v (b,np,S,HD) with HD = heads*key_dim
  bool shapes_ok(OpKernelContext *ctx) const {
    const Tensor &q = ctx->input(0), &k = ctx->input(1), &v = ctx->input(2);
    const int hd = heads_ * key_dim_;
    return q.dims() == 3 && k.dims() == 4 && dim(q, 2) == hd && dim(k, 3) == hd && dim(k, 0) == dim(q, 0) &&
           dim(k, 1) == dim(q, 1) && same_shape(v, k);
  }
  int heads_, key_dim_;
};

class PointAttentionContractGpuOp : public AttentionAttrs {
 public:
  explicit PointAttentionContractGpuOp(OpKernelConstruction *c) : AttentionAttrs(c) {}
  void Compute(OpKernelContext *ctx) override {
    OP_REQUIRES(ctx, shapes_ok(ctx),
                errors::InvalidArgument("PointAttentionContract expects q (b,np,heads*key_dim), k and v (b,np,nsample,heads*key_dim)"));
    const int G = dim(ctx->input(0), 0) * dim(ctx->input(0), 1), S = dim(ctx->input(1), 2);
    Tensor *out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, ctx->input(0).shape(), &out));
    PCSHIM_CHECK_RC(ctx, pc_attention_fwd(G, S, heads_, key_dim_, F(ctx->input(0)), F(ctx->input(1)), F(ctx->input(2)), F(out),
                                          PCSHIM_STREAM(ctx)),
                    "pc_attention_fwd");
  }
};
REGISTER_KERNEL_BUILDER(Name("PointAttentionContract").Device(DEVICE_GPU), PointAttentionContractGpuOp);

class PointAttentionContractGradGpuOp : public AttentionAttrs {
 public:
  explicit PointAttentionContractGradGpuOp(OpKernelConstruction *c) : AttentionAttrs(c) {}
  void Compute(OpKernelContext *ctx) override {
    OP_REQUIRES(ctx, shapes_ok(ctx),
                errors::InvalidArgument("PointAttentionContractGrad expects q (b,np,heads*key_dim), k and v (b,np,nsample,heads*key_dim)"));
    const int G = dim(ctx->input(0), 0) * dim(ctx->input(0), 1), S = dim(ctx->input(1), 2);
    const Tensor &go = ctx->input(3);
    OP_REQUIRES(ctx, same_shape(go, ctx->input(0)), errors::InvalidArgument("PointAttentionContractGrad expects grad_out shaped like q"));
    Tensor *dq = nullptr, *dk = nullptr, *dv = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, ctx->input(0).shape(), &dq));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, ctx->input(1).shape(), &dk));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, ctx->input(2).shape(), &dv));
    PCSHIM_CHECK_RC(ctx, pc_attention_bwd(G, S, heads_, key_dim_, F(ctx->input(0)), F(ctx->input(1)), F(ctx->input(2)), F(go), F(dq),
                                          F(dk), F(dv), PCSHIM_STREAM(ctx)),
                    "pc_attention_bwd");
  }
};
REGISTER_KERNEL_BUILDER(Name("PointAttentionContractGrad").Device(DEVICE_GPU), PointAttentionContractGradGpuOp);

class PointAttentionLayerGpuOp : public OpKernel {
 public:
  explicit PointAttentionLayerGpuOp(OpKernelConstruction *c) : OpKernel(c) {}
  void Compute(OpKernelContext *ctx) override {
    const Tensor &query = ctx->input(0), &inp = ctx->input(1);
    OP_REQUIRES(ctx, inp.dims() == 4, errors::InvalidArgument("PointAttentionLayer expects (b,np,nsample,C) inp shape"));
    const int b = dim(inp, 0), np = dim(inp, 1), S = dim(inp, 2), C = dim(inp, 3);
    OP_REQUIRES(ctx, query.dims() == 4 && dim(query, 0) == b && dim(query, 1) == np && dim(query, 2) == 1 && dim(query, 3) == C,
                errors::InvalidArgument("PointAttentionLayer expects (b,np,1,C) query shape"));
    for (int i = 2; i < 8; i += 2) {
      OP_REQUIRES(ctx, ctx->input(i).dims() == 2 && dim(ctx->input(i), 0) == C && dim(ctx->input(i), 1) == C,
                  errors::InvalidArgument("PointAttentionLayer expects (C,C) Dense kernels"));
      OP_REQUIRES(ctx, ctx->input(i + 1).dims() == 1 && dim(ctx->input(i + 1), 0) == C,
                  errors::InvalidArgument("PointAttentionLayer expects (C) Dense biases"));
    }
    const int G = b * np;
    const size_t ws_bytes = pc_attention_layer_workspace_bytes(G, S, C);
    OP_REQUIRES(ctx, G == 0 || ws_bytes > 0,
                errors::InvalidArgument("PointAttentionLayer supports nsample = 32 and C in {64,128,256,512}; use three Dense + "
                                      "PointAttentionContract otherwise"));
    Tensor *out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, np, C}, &out));
    Tensor ws_t;
    void *ws = nullptr;
    OP_REQUIRES_OK(ctx, scratch(ctx, ws_bytes, &ws_t, &ws));
    PCSHIM_CHECK_RC(ctx, pc_attention_layer_fwd(G, S, C, F(query), F(inp), F(ctx->input(2)), F(ctx->input(3)), F(ctx->input(4)),
                                                F(ctx->input(5)), F(ctx->input(6)), F(ctx->input(7)), F(out), ws, PCSHIM_STREAM(ctx)),
                    "pc_attention_layer_fwd");
  }
};
REGISTER_KERNEL_BUILDER(Name("PointAttentionLayer").Device(DEVICE_GPU), PointAttentionLayerGpuOp);

}  // namespace pcshim
