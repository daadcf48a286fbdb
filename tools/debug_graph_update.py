import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200env
rl = importlib.import_module("sac-gat-her_transportationrl_b200.rl")
net = b200env.Network(b200env.load_sioux_falls())
dev = torch.device("cuda", 0)
def mk(graphed, warm=3):
    torch.manual_seed(5)
    a = rl.DiscreteSAC(net, 4, 6, 32, 16, num_layers=3, lr=1e-3, grad_clip=1.0, share_critic_encoder=False, device=dev, cuda_graph=graphed)
    a._GRAPH_WARMUP = warm
    return a
eager, graphed, capt = mk(False), mk(True), mk(True, 10**9)   # capt: capturable Adam, never captured
g = torch.Generator(device=dev).manual_seed(11)
def make(B):
    return {"node_x": torch.rand(B, 24, 4, device=dev, generator=g), "edge_attr": torch.rand(B, 76, 6, device=dev, generator=g),
            "action_mask": (torch.rand(B, 76, device=dev, generator=g) > 0.4).to(torch.uint8),
            "next_node_x": torch.rand(B, 24, 4, device=dev, generator=g), "next_edge_attr": torch.rand(B, 76, 6, device=dev, generator=g),
            "next_action_mask": (torch.rand(B, 76, device=dev, generator=g) > 0.4).to(torch.uint8),
            "action": torch.randint(0, 76, (B,), device=dev, generator=g), "reward": torch.rand(B, device=dev, generator=g),
            "done": (torch.rand(B, device=dev, generator=g) > 0.8).float()}, torch.rand(B, device=dev, generator=g) + 0.5
for it in range(8):
    batch, w = make(16)
    outs = [a.update(batch, weights=w, alpha_max=2.5) for a in (eager, graphed, capt)]
    def dmax(x, y):
        return max(float((p - q).abs().max()) for p, q in zip(x.actor.parameters(), y.actor.parameters()))
    def dmean(x, y):
        return max(float((p - q).abs().mean()) for p, q in zip(x.actor.parameters(), y.actor.parameters()))
    print(it, "critic_loss", [float(o["critic_loss"]) for o in outs], "actor max|d| eager-graph %.2e eager-capt %.2e graph-capt %.2e  mean %.2e %.2e"
          % (dmax(eager, graphed), dmax(eager, capt), dmax(graphed, capt), dmean(eager, graphed), dmean(graphed, capt)))
print("---- td_errors check")
eager, graphed = mk(False), mk(True)
for it in range(8):
    batch, w = make(16)
    oe = eager.update(batch, weights=w, alpha_max=2.5)
    keep = oe["td_errors"].clone()
    torch.cuda.synchronize()
    og = graphed.update(batch, weights=w, alpha_max=2.5)
    torch.cuda.synchronize()
    print(it, "eager td changed by graphed.update:", bool((keep != oe["td_errors"]).any()), "ptr", oe["td_errors"].data_ptr(),
          "max|og-keep| %.2e" % float((og["td_errors"] - keep).abs().max()), "og ptr", og["td_errors"].data_ptr())
