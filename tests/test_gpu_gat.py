"""GAT edge-softmax + aggregation kernel vs a plain PyTorch fp32 restatement of the same op
(GATConv semantics: softmax over incoming links + self loop, leaky_relu 0.2)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def torch_reference(net, xw, a_src, a_dst, a_edge, slope=0.2):
    """Index-based fp32 reference: build the (E+N) link list, segment softmax by target."""
    B, N, H, C = xw.shape
    src = torch.from_numpy(np.concatenate([net.edge_index[0], np.arange(N)])).to(xw.device)
    dst = torch.from_numpy(np.concatenate([net.edge_index[1], np.arange(N)])).to(xw.device)
    e = torch.nn.functional.leaky_relu(a_src[:, src] + a_dst[:, dst] + a_edge, slope)  # [B, E+N, H]
    mx = torch.full((B, N, H), -float("inf"), device=xw.device).scatter_reduce(
        1, dst.view(1, -1, 1).expand(B, -1, H), e, reduce="amax")
    ex = torch.exp(e - mx[:, dst])
    den = torch.zeros((B, N, H), device=xw.device).index_add(1, dst, ex) + 1e-16
    alpha = ex / den[:, dst]
    out = torch.zeros_like(xw).index_add(1, dst, alpha.unsqueeze(-1) * xw[:, src])
    return out, alpha


@pytest.mark.parametrize("H,C", [(4, 256), (1, 256), (2, 12)])
def test_gat_forward_backward_vs_torch(b200env, native_lib, sf_graph, H, C):
    net = b200env.Network(sf_graph)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    B, N, E = 5, net.num_nodes, net.num_edges
    xw = torch.randn(B, N, H, C, device=dev, requires_grad=True)
    a_src = torch.randn(B, N, H, device=dev, requires_grad=True)
    a_dst = torch.randn(B, N, H, device=dev, requires_grad=True)
    a_edge = torch.randn(B, E + N, H, device=dev, requires_grad=True)
    out, alpha = b200env.gat_edge_softmax_aggregate(net, xw, a_src, a_dst, a_edge)
    ref_out, ref_alpha = torch_reference(net, xw, a_src, a_dst, a_edge)
    torch.testing.assert_close(alpha, ref_alpha, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out, ref_out, rtol=1e-4, atol=1e-5)
    # coefficients of each target node sum to one
    dst = torch.from_numpy(np.concatenate([net.edge_index[1], np.arange(N)])).to(dev)
    sums = torch.zeros(B, N, H, device=dev).index_add(1, dst, alpha)
    torch.testing.assert_close(sums, torch.ones_like(sums), rtol=1e-5, atol=1e-5)
    g = torch.randn_like(out)
    grads = torch.autograd.grad(out, (xw, a_src, a_dst, a_edge), g)
    ref_grads = torch.autograd.grad(ref_out, (xw, a_src, a_dst, a_edge), g)
    for got, ref, name in zip(grads, ref_grads, ("xw", "a_src", "a_dst", "a_edge")):
        torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4, msg=lambda m, n=name: f"grad {n}: {m}")


def test_gat_large_network_uses_gather_path(b200env, native_lib):
    g = b200env.synthetic.anaheim_size(0)
    net = b200env.Network(g)
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    B, N, E, H, C = 2, net.num_nodes, net.num_edges, 4, 64
    xw = torch.randn(B, N, H, C, device=dev, requires_grad=True)
    a_src = torch.randn(B, N, H, device=dev, requires_grad=True)
    a_dst = torch.randn(B, N, H, device=dev, requires_grad=True)
    a_edge = torch.randn(B, E + N, H, device=dev, requires_grad=True)
    out, alpha = b200env.gat_edge_softmax_aggregate(net, xw, a_src, a_dst, a_edge)
    ref_out, ref_alpha = torch_reference(net, xw, a_src, a_dst, a_edge)
    torch.testing.assert_close(out, ref_out, rtol=1e-4, atol=1e-5)
    g_ = torch.randn_like(out)
    grads = torch.autograd.grad(out, (xw, a_src, a_dst, a_edge), g_)
    ref_grads = torch.autograd.grad(ref_out, (xw, a_src, a_dst, a_edge), g_)
    for got, ref in zip(grads, ref_grads):
        torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)


def test_gat_encoder_shapes_and_training_step(b200env, native_lib, sf_graph):
    net = b200env.Network(sf_graph)
    dev = torch.device("cuda", 0)
    torch.manual_seed(2)
    enc = b200env.GATEncoderB200(net, in_dim=4, hidden_dim=64, out_dim=64, edge_dim=6).to(dev)
    assert set(dict(enc.layers[0].named_parameters())) == {"att_src", "att_dst", "att_edge", "bias", "lin.weight",
                                                            "lin_edge.weight"}  # PyG GATConv key layout
    x = torch.randn(7, 24, 4, device=dev)
    ea = torch.randn(7, 76, 6, device=dev)
    emb, ctx, attn = enc(x, ea, return_attention=True)
    assert emb.shape == (7, 24, 64) and ctx.shape == (7, 128) and attn.shape == (7, 100, 1)
    loss = emb.square().mean() + ctx.mean()
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in enc.parameters())
