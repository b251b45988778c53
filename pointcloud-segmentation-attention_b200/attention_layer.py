"""Per-neighbourhood attention, mirroring attention_points/attention_scannet/attention_layer.py.

``attention_contract(Q, K, V, num_heads, key_dim)`` is the hot part of ``AttentionLayer.call`` (:35-42): the raw
reshape to heads, QK^T / sqrt(d), softmax over the neighbourhood and the weighted value sum, as one CUDA kernel with a
matching backward kernel.  ``AttentionLayer`` keeps the reference class's constructor and ``[input, query]`` call
convention; its three Dense projections (:24-26,31-34) are stock dense layers (torch.nn.Linear here, tf.layers.Dense
there) and stay outside the accelerated path exactly like the reference's conv2d stack.
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
                and self.key_dim == 4):
            # inference at the four ScanNet attention widths: projections + contraction in one tensor-core kernel,
            # K and V never stored
            lead = inp.shape[:-2]
            out = attention_layer_fused(query.reshape(-1, C), inp.reshape(-1, 32, C),
                                        self.query_net.weight.t().contiguous(), self.query_net.bias,
                                        self.key_net.weight.t().contiguous(), self.key_net.bias,
                                        self.value_net.weight.t().contiguous(), self.value_net.bias)
            return out.reshape(*lead, C)
        Q = self.query_net(query)          # (B,np,1,HD)
        K = self.key_net(inp)              # (B,np,S,HD)
        V = self.value_net(inp)
        return attention_contract(Q[:, :, 0, :], K, V, self.num_heads, self.key_dim)
