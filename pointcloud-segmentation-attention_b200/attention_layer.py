"""Per-neighbourhood attention, mirroring attention_points/attention_scannet/attention_layer.py.

``attention_contract(Q, K, V, num_heads, key_dim)`` is the hot part of ``AttentionLayer.call`` (:35-42): the raw
reshape to heads, QK^T / sqrt(d), softmax over the neighbourhood and the weighted value sum, as one CUDA kernel with a
matching backward kernel.  ``AttentionLayer`` keeps the reference class's constructor and ``[input, query]`` call
convention.  Its three Dense projections (:24-26,31-34; tf.layers.Dense there) run on this library's tcgen05 Dense
engine (csrc/gemm_tf32.cu, 3xTF32) in both directions -- forward, input gradient and weight gradient -- so a training
step never touches a vendor GEMM; without gradients the whole layer is ONE tensor-core kernel (pc_attention_layer_fwd).
``torch.nn.Linear`` is only the parameter container (weight (out, in), bias).
"""
import torch

from . import _lib


class _AttentionContract(torch.autograd.Function):
    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, Q, K, V, H, D):
        G, S, HD = K.shape
        out = torch.empty((G, HD), dtype=torch.float32, device=K.device)
        rc = _lib.lib().pc_attention_fwd(G, S, H, D, _lib.ptr(Q), _lib.ptr(K), _lib.ptr(V), _lib.ptr(out),
                                         _lib.stream())
        _lib.check(rc, "pc_attention_fwd")
        ctx.save_for_backward(Q, K, V)
        ctx.hd = (H, D)
        return out

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, dout):
        Q, K, V = ctx.saved_tensors
        H, D = ctx.hd
        G, S, HD = K.shape
        dout = _lib.cuda_f32(dout, "dout")
        dQ, dK, dV = torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
        rc = _lib.lib().pc_attention_bwd(G, S, H, D, _lib.ptr(Q), _lib.ptr(K), _lib.ptr(V), _lib.ptr(dout),
                                         _lib.ptr(dQ), _lib.ptr(dK), _lib.ptr(dV), _lib.stream())
        _lib.check(rc, "pc_attention_bwd")
        return dQ, dK, dV, None, None


FUSED_LAYER = True  # AttentionLayer.forward uses pc_attention_layer_fwd where it applies (no-grad, S = 32, C = 64)


@_lib.on_tensor_device
def attention_contract(Q, K, V, num_heads, key_dim):
    """Q (..., HD), K (..., S, HD), V (..., S, HD) -> (..., HD) with HD = num_heads*key_dim."""
    HD = int(num_heads) * int(key_dim)
    if K.shape[-1] != HD or V.shape != K.shape or Q.shape[-1] != HD or Q.shape[:-1] != K.shape[:-2]:
        raise ValueError("attention_contract expects Q (...,H*D), K and V (...,S,H*D)")
    lead = K.shape[:-2]
    S = K.shape[-2]
    Qf = _lib.cuda_f32(Q, "Q").reshape(-1, HD)
    Kf = _lib.cuda_f32(K, "K").reshape(-1, S, HD)
    Vf = _lib.cuda_f32(V, "V").reshape(-1, S, HD)
    out = _AttentionContract.apply(Qf, Kf, Vf, int(num_heads), int(key_dim))
    return out.reshape(*lead, HD)


@_lib.on_tensor_device
def attention_layer_fused(xq, x, wq, bq, wk, bk, wv, bv):
    """The whole AttentionLayer.call (attention_layer.py:29-45) in one tcgen05 kernel (csrc/attention_layer.cu):
    xq (G,C) query rows, x (G,S,C) grouped rows, Dense kernels w* as (C_in, C_out) like Keras, biases (C) or None
    -> (G, C).  Forward only; supported for S = 32, C in {64, 128, 256, 512} (key_dim = output_dim = 4), else
    NotImplementedError."""
    x = _lib.cuda_f32(x.detach(), "x")
    xq = _lib.cuda_f32(xq.detach(), "xq")
    G, S, C = x.shape
    L = _lib.lib()
    nbytes = L.pc_attention_layer_workspace_bytes(G, S, C)
    ws = _lib.workspace(max(nbytes, 16), x.device)
    out = torch.empty((G, C), dtype=torch.float32, device=x.device)
    args = [_lib.ptr(_lib.cuda_f32(t.detach(), "weights")) if t is not None else None for t in (wq, bq, wk, bk, wv, bv)]
    rc = L.pc_attention_layer_fwd(G, S, C, _lib.ptr(xq), _lib.ptr(x), *args, _lib.ptr(out), _lib.ptr(ws), _lib.stream())
    _lib.check(rc, "pc_attention_layer_fwd")
    return out


class PreparedAttentionLayer:
    """The fused layer with everything that depends on the weights alone kept across calls: the (in, out) copies of the
    three Dense kernels and the tensor-core operand image inside the workspace (pc_attention_layer_prepare).  Rebuilt
    only when a parameter changed (torch's version counters) or a larger G needs a larger workspace."""

    def __init__(self):
        self.key, self.ws, self.G, self.w = None, None, 0, None

    @_lib.on_tensor_device
    def __call__(self, xq, x, query_net, key_net, value_net):
        x = _lib.cuda_f32(x.detach(), "x")
        xq = _lib.cuda_f32(xq.detach(), "xq")
        G, S, C = x.shape
        L = _lib.lib()
        params = [query_net.weight, query_net.bias, key_net.weight, key_net.bias, value_net.weight, value_net.bias]
        key = tuple((None if p is None else (p.data_ptr(), p._version)) for p in params) + (S, C, str(x.device))
        if key != self.key or G > self.G:
            with torch.no_grad():
                self.w = [None if p is None else (p.detach().t().contiguous() if p.dim() == 2 else p.detach().contiguous())
                          for p in params]          # Dense kernels as (in, out), biases as they are
            self.ws = _lib.workspace(max(L.pc_attention_layer_workspace_bytes(G, S, C), 16), x.device)
            rc = L.pc_attention_layer_prepare(S, C, *[_lib.ptr(t) for t in self.w], _lib.ptr(self.ws), _lib.stream())
            _lib.check(rc, "pc_attention_layer_prepare")
            self.key, self.G = key, G
        out = torch.empty((G, C), dtype=torch.float32, device=x.device)
        rc = L.pc_attention_layer_fwd_prepared(G, S, C, _lib.ptr(xq), 0, _lib.ptr(x), *[_lib.ptr(t) for t in self.w],
                                               _lib.ptr(out), _lib.ptr(self.ws), _lib.stream())
        _lib.check(rc, "pc_attention_layer_fwd_prepared")
        return out


class AttentionLayer(torch.nn.Module):
    """AttentionLayer(output_dim, key_dim, num_heads) -- attention_layer.py:10-45.

    call([input (B,np,S,C), query (B,np,1,C)]) -> (B, np, num_heads*key_dim).  As in the reference the value heads
    are reshaped with key_dim (:35), so output_dim must equal key_dim (it does at every call site, :256,:314).
    """

    def __init__(self, output_dim, key_dim, num_heads=16, in_features=None):
        super().__init__()
        if output_dim != key_dim:
            raise ValueError("the reference reshapes V with key_dim (attention_layer.py:35): output_dim must equal key_dim")
        self.output_dim, self.key_dim, self.num_heads = output_dim, key_dim, num_heads
        self.in_features = in_features
        self.query_net = self.key_net = self.value_net = None
        self._prepared = PreparedAttentionLayer()
        if in_features is not None:
            self._build(in_features)

    def _build(self, cin):
        hd = self.key_dim * self.num_heads
        self.query_net = torch.nn.Linear(cin, hd)
        self.key_net = torch.nn.Linear(cin, hd)
        self.value_net = torch.nn.Linear(cin, self.output_dim * self.num_heads)

    @_lib.on_tensor_device
    def forward(self, inputs):
        inp, query = inputs
        if self.query_net is None:
            self._build(inp.shape[-1])
            self.to(inp.device)
        hd = self.key_dim * self.num_heads
        C = inp.shape[-1]
        if (FUSED_LAYER and not torch.is_grad_enabled() and C in (64, 128, 256, 512) and hd == C and inp.shape[-2] == 32
                and self.key_dim == 4 and query.shape[-1] == C):   # (a 3-wide query, pooling_attention_layer.py:38, is not)
            # inference at the four ScanNet attention widths: projections + contraction in one tensor-core kernel,
            # K and V never stored
            lead = inp.shape[:-2]
            out = self._prepared(query.reshape(-1, C), inp.reshape(-1, 32, C), self.query_net, self.key_net, self.value_net)
            return out.reshape(*lead, C)
        # training (or a shape the fused kernel does not cover): projections on the tcgen05 Dense engine, forward and
        # backward (sa_modules.dense_layer), contraction + its backward on pc_attention_fwd / pc_attention_bwd
        from .sa_modules import dense_layer
        Q = dense_layer(query, self.query_net.weight, self.query_net.bias, linear_layout=True)    # (B,np,1,HD)
        K = dense_layer(inp, self.key_net.weight, self.key_net.bias, linear_layout=True)          # (B,np,S,HD)
        V = dense_layer(inp, self.value_net.weight, self.value_net.bias, linear_layout=True)
        return attention_contract(Q[..., 0, :], K, V, self.num_heads, self.key_dim)
