"""GAT encoder on a fixed topology, batched over scenarios.

Reference: ``src/models/gat_encoder.py:9-53`` (3 x ``torch_geometric.nn.GATConv`` with
``edge_dim``, heads 4/4/1, residual + LayerNorm + ReLU, last layer LayerNorm + ELU, mean||max
pooling).  PyG is a third-party dependency the reference leaves unpinned and that is not
installed here; its GATConv semantics are restated (self loops with ``fill_value='mean'``,
``negative_slope=0.2``, softmax over the incoming links of each target node, bias after
aggregation, heads concatenated or averaged).  Parameter names follow PyG's state-dict layout
(``lin.weight``, ``lin_edge.weight``, ``att_src``, ``att_dst``, ``att_edge``, ``bias``) so the
reference's checkpoints map one-to-one.

Only the edge-softmax + neighbour aggregation runs in our CUDA kernel
(``b200_gat_edge_softmax_aggregate_{fwd,bwd}``); the dense projections stay torch GEMMs, the only
real contractions on this path (BASELINE.json north_star).  Because every scenario shares the
topology, a batch is a dense ``[B, N, F]`` tensor instead of PyG's concatenated graph.
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F

from .. import _native
from ..env.network import Network


class _EdgeSoftmaxAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xw, a_src, a_dst, a_edge, handle, slope):
        B, N, H, C = xw.shape
        xw, a_src, a_dst, a_edge = (t.contiguous().float() for t in (xw, a_src, a_dst, a_edge))
        out = torch.empty_like(xw)
        alpha = torch.empty_like(a_edge)
        stream = torch.cuda.current_stream(xw.device).cuda_stream
        _native.check(_native.lib().b200_gat_edge_softmax_aggregate_fwd(
            handle, B, H, C, xw.data_ptr(), a_src.data_ptr(), a_dst.data_ptr(), a_edge.data_ptr(), float(slope),
            out.data_ptr(), alpha.data_ptr(), stream), "b200_gat_edge_softmax_aggregate_fwd")
        ctx.save_for_backward(xw, a_src, a_dst, a_edge, alpha)
        ctx.handle, ctx.slope = handle, slope
        ctx.mark_non_differentiable(alpha)
        return out, alpha

    @staticmethod
    def backward(ctx, grad_out, _grad_alpha):
        xw, a_src, a_dst, a_edge, alpha = ctx.saved_tensors
        B, N, H, C = xw.shape
        grad_out = grad_out.contiguous().float()
        g_xw, g_as, g_ad, g_ae = (torch.empty_like(t) for t in (xw, a_src, a_dst, a_edge))
        stream = torch.cuda.current_stream(xw.device).cuda_stream
        _native.check(_native.lib().b200_gat_edge_softmax_aggregate_bwd(
            ctx.handle, B, H, C, xw.data_ptr(), a_src.data_ptr(), a_dst.data_ptr(), a_edge.data_ptr(), alpha.data_ptr(),
            grad_out.data_ptr(), float(ctx.slope), g_xw.data_ptr(), g_as.data_ptr(), g_ad.data_ptr(), g_ae.data_ptr(),
            stream), "b200_gat_edge_softmax_aggregate_bwd")
        return g_xw, g_as, g_ad, g_ae, None, None


def gat_edge_softmax_aggregate(network: Network, xw, a_src, a_dst, a_edge, negative_slope: float = 0.2):
    """xw [B,N,H,C], a_src/a_dst [B,N,H], a_edge [B,E+N,H] (rows E.. are the self loops) ->
    (out [B,N,H,C], alpha [B,E+N,H])."""
    if not xw.is_cuda:
        raise _native.NativeLibraryError("the GAT aggregation kernel runs only on a CUDA device")
    with torch.cuda.device(xw.device):
        handle = network.handle
    return _EdgeSoftmaxAggregate.apply(xw, a_src, a_dst, a_edge, handle, negative_slope)


class GATConvB200(nn.Module):
    """GATConv(in, out, heads, concat, edge_dim) on the fixed topology of ``network``."""

    def __init__(self, network: Network, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 edge_dim: int | None = None, negative_slope: float = 0.2, bias: bool = True):
        super().__init__()
        self.network, self.heads, self.out_channels, self.concat = network, heads, out_channels, concat
        self.negative_slope, self.edge_dim = negative_slope, edge_dim
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        if edge_dim is not None:
            self.lin_edge = nn.Linear(edge_dim, heads * out_channels, bias=False)
            self.att_edge = nn.Parameter(torch.empty(1, heads, out_channels))
        else:
            self.lin_edge = None
            self.register_parameter("att_edge", None)
        self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels)) if bias else None
        dst = torch.from_numpy(network.edge_index[1].copy())
        self.register_buffer("_dst", dst, persistent=False)
        deg = torch.bincount(dst, minlength=network.num_nodes).clamp(min=1).float()
        self.register_buffer("_in_deg", deg, persistent=False)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.lin.weight)
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)
        if self.lin_edge is not None:
            nn.init.xavier_uniform_(self.lin_edge.weight)
            nn.init.xavier_uniform_(self.att_edge)

    def edge_terms(self, x, edge_attr):
        """The torch-GEMM half: projections and attention dot products."""
        B, N, _ = x.shape
        H, C = self.heads, self.out_channels
        xw = self.lin(x).view(B, N, H, C)
        # <lin(x), att_src> and <lin(x), att_dst> per head as ONE small GEMM on x: the attention vectors
        # are folded into the projection ([2H, in] matrix) instead of re-reading the [B, N, H*C]
        # projected features twice (same parameters, same sums up to association)
        w3 = self.lin.weight.view(H, C, -1)
        w_att = torch.cat([(w3 * self.att_src.view(H, C, 1)).sum(dim=1), (w3 * self.att_dst.view(H, C, 1)).sum(dim=1)])
        a_both = F.linear(x, w_att)
        a_src, a_dst = a_both[..., :H], a_both[..., H:]
        E = self.network.num_edges
        if self.lin_edge is not None and edge_attr is not None:
            # self-loop attributes = mean of the attributes of the node's incoming links (fill_value='mean')
            loop_attr = torch.zeros(B, N, edge_attr.shape[-1], device=x.device, dtype=edge_attr.dtype)
            loop_attr.index_add_(1, self._dst, edge_attr)
            loop_attr = loop_attr / self._in_deg.view(1, N, 1)
            full = torch.cat([edge_attr, loop_attr], dim=1)
            # <lin_edge(e), att_edge> per head without the [B, E+N, H*C] projection PyG materialises:
            # the projected link features are used for nothing else, so the two linear maps fold
            # into one [H, edge_dim] matrix (same parameters, same sum up to association)
            w_eff = (self.lin_edge.weight.view(H, C, -1) * self.att_edge.view(H, C, 1)).sum(dim=1)
            a_edge = F.linear(full, w_eff)
        else:
            a_edge = torch.zeros(B, E + N, H, device=x.device, dtype=xw.dtype)
        return xw, a_src, a_dst, a_edge

    def forward(self, x, edge_attr=None, return_attention_weights: bool = False):
        B, N, _ = x.shape
        xw, a_src, a_dst, a_edge = self.edge_terms(x, edge_attr)
        out, alpha = gat_edge_softmax_aggregate(self.network, xw, a_src, a_dst, a_edge, self.negative_slope)
        out = out.reshape(B, N, self.heads * self.out_channels) if self.concat else out.mean(dim=2)
        if self.bias is not None:
            out = out + self.bias
        return (out, alpha) if return_attention_weights else out


class GATEncoderB200(nn.Module):
    """Same layer stack as the reference's GATEncoder (src/models/gat_encoder.py:9-53)."""

    def __init__(self, network: Network, in_dim: int, hidden_dim: int, out_dim: int, edge_dim: int, heads: int = 4,
                 num_layers: int = 3):
        super().__init__()
        self.num_layers = max(2, num_layers)
        self.layers = nn.ModuleList()
        self.layers.append(GATConvB200(network, in_dim, hidden_dim, heads=heads, concat=True, edge_dim=edge_dim))
        for _ in range(self.num_layers - 2):
            self.layers.append(GATConvB200(network, hidden_dim * heads, hidden_dim, heads=heads, concat=True,
                                           edge_dim=edge_dim))
        self.layers.append(GATConvB200(network, hidden_dim * heads, out_dim, heads=1, concat=False, edge_dim=edge_dim))
        self.input_proj = nn.Linear(in_dim, hidden_dim * heads)
        self.norms = nn.ModuleList(
            nn.LayerNorm(out_dim if i == self.num_layers - 1 else hidden_dim * heads) for i in range(self.num_layers))

    def forward(self, x, edge_attr, return_attention: bool = False):
        """x [B,N,in_dim], edge_attr [B,E,edge_dim] -> (node_emb [B,N,out], global_ctx [B,2*out], attn)."""
        attn = None
        last = len(self.layers) - 1
        for i, layer in enumerate(self.layers):
            if i == last:
                if return_attention:
                    x, attn = layer(x, edge_attr, return_attention_weights=True)
                else:
                    x = layer(x, edge_attr)
                x = F.elu(self.norms[i](x))
            else:
                x_in = self.input_proj(x) if i == 0 else x
                x = torch.relu(self.norms[i](layer(x, edge_attr)) + x_in)
        global_ctx = torch.cat([x.mean(dim=1), x.max(dim=1).values], dim=1)
        return x, global_ctx, attn
