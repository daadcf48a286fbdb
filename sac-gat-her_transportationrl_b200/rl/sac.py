"""Discrete soft actor-critic over link-repair actions, dense-batched on a fixed topology.

Reference: src/rl/sac.py - ``Actor`` (:23-46), ``Critic`` (:49-78), ``DiscreteSAC`` (:81-287).
The reference batches graphs the PyG way (concatenated nodes/links + a ``batch`` vector,
``torch_geometric.utils.softmax`` / ``torch_scatter.scatter_sum`` per graph).  Every scenario has
the same topology here, so a batch is ``node_x [B,N,F]``, ``edge_attr [B,E,F]``,
``action_mask [B,E]``: the per-graph softmax is a row softmax, ``scatter_sum`` is ``sum(dim=1)``
and the per-item ``nonzero`` action indexing of src/train.py:841-845 disappears (the action is
the link id).  Module and parameter names are the reference's, so ``state_dict`` layouts match
(``actor.encoder.layers.0.lin.weight`` ...) and ``save`` / ``load`` use the reference's file
layout (:265-287).  The GAT message passing runs in the CUDA kernel of ``models/gat.py``; the
dense projections are torch GEMMs.

Multi-GPU: ``update`` averages gradients over ranks (NCCL all-reduce, ``dist.py``) before every
optimiser step when ``torch.distributed`` is initialised - a data-parallel learner.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..dist import allreduce_mean_gradients
from ..env.network import Network
from ..models.gat import GATEncoderB200


@dataclass
class SACOutput:
    action: torch.Tensor  # [B] link ids
    log_prob: torch.Tensor  # [B]
    probs: torch.Tensor  # [B, E]


class _NarrowLayerNorm(nn.LayerNorm):
    """``nn.LayerNorm`` over a handful of features (the 4 node / 6 link inputs).  torch's CUDA kernel for a feature
    count that is not a multiple of 4 launches one thread block per row - 0.38 ms for the 77 824 link rows of 1024
    scenarios, as long as a 1024 x 1024 projection - so the same formula (biased variance, eps inside the root) is
    written with elementwise ops here.  Same parameters and state-dict keys."""

    def forward(self, x):
        if not x.is_cuda or self.normalized_shape[0] > 8:
            return super().forward(x)
        mean = x.mean(dim=-1, keepdim=True)
        xc = x - mean
        var = (xc * xc).mean(dim=-1, keepdim=True)
        return xc * torch.rsqrt(var + self.eps) * self.weight + self.bias


class _EdgeHead(nn.Module):
    """Shared body of Actor and Critic: norms -> encoder -> per-link MLP on
    [emb(src) | emb(dst) | edge_attr | global mean||max context]."""

    def __init__(self, network: Network, node_in: int, edge_in: int, hidden: int, embed: int, num_layers: int = 3,
                 encoder: Optional[GATEncoderB200] = None):
        super().__init__()
        self.node_norm = _NarrowLayerNorm(node_in)
        self.edge_norm = _NarrowLayerNorm(edge_in)
        self.encoder = encoder if encoder is not None else GATEncoderB200(network, node_in, hidden, embed,
                                                                          edge_dim=edge_in, num_layers=num_layers)
        self.edge_mlp = nn.Sequential(nn.Linear(embed * 4 + edge_in, hidden), nn.ReLU(), nn.Linear(hidden, 1))
        self.register_buffer("_src", torch.from_numpy(network.edge_index[0].copy()), persistent=False)
        self.register_buffer("_dst", torch.from_numpy(network.edge_index[1].copy()), persistent=False)

    def edge_scores(self, node_x, edge_attr, return_attention: bool = False):
        node_x = self.node_norm(node_x)
        edge_attr = self.edge_norm(edge_attr)
        node_emb, ctx, attn = self.encoder(node_x, edge_attr, return_attention=return_attention)
        # First linear of edge_mlp on [emb(src) | emb(dst) | edge_attr | ctx] WITHOUT materialising the
        # [B, E, 4*embed + edge_in] concatenation the reference builds (src/rl/sac.py:43, :75): the
        # weight is split by column block, the node blocks are applied per NODE (N rows instead of E)
        # and gathered afterwards, the context block once per graph.  Same parameters, same sum.
        lin0 = self.edge_mlp[0]
        D = node_emb.shape[-1]
        w_src, w_dst, w_edge, w_ctx = torch.split(lin0.weight, [D, D, edge_attr.shape[-1], ctx.shape[-1]], dim=1)
        h = (F.linear(node_emb, w_src)[:, self._src] + F.linear(node_emb, w_dst)[:, self._dst]
             + F.linear(edge_attr, w_edge) + F.linear(ctx, w_ctx, lin0.bias).unsqueeze(1))
        return self.edge_mlp[2](self.edge_mlp[1](h)).squeeze(-1), attn


class Actor(_EdgeHead):
    def forward(self, node_x, edge_attr, action_mask, return_attention: bool = False):
        logits, attn = self.edge_scores(node_x, edge_attr, return_attention)
        logits = logits.masked_fill(action_mask <= 0, -1e9)
        probs = torch.softmax(logits, dim=1)  # PyG softmax per graph = row softmax on a fixed topology
        return logits, probs, attn


class Critic(_EdgeHead):
    def forward(self, node_x, edge_attr):
        q, _ = self.edge_scores(node_x, edge_attr)
        return q


class DiscreteSAC:
    _GRAPH_WARMUP = 3   # eager updates before the stages are captured

    def __init__(self, network: Network, node_in: int, edge_in: int, hidden: int, embed: int, num_layers: int = 3,
                 lr: float = 3e-4, actor_lr: Optional[float] = None, critic_lr: Optional[float] = None,
                 alpha_lr: Optional[float] = None, grad_clip: Optional[float] = None, gamma: float = 0.99,
                 target_tau: float = 0.005, target_entropy: Optional[float] = None, target_entropy_ratio: float = 0.1,
                 alpha_init: float = 0.1, share_critic_encoder: bool = True, device="cuda", cuda_graph: bool = False):
        dev = torch.device(device)
        self.device = dev
        # learner update as replayed CUDA graphs (see _update_graphed); Adam keeps its step count on the device
        self.cuda_graph = bool(cuda_graph) and dev.type == "cuda"
        self._graph, self._graph_calls = None, 0
        cap = dict(capturable=True) if self.cuda_graph else {}
        mk = dict(network=network, node_in=node_in, edge_in=edge_in, hidden=hidden, embed=embed, num_layers=num_layers)
        self.actor = Actor(**mk).to(dev)
        self.share_critic_encoder = share_critic_encoder
        if share_critic_encoder:
            enc = lambda: GATEncoderB200(network, node_in, hidden, embed, edge_dim=edge_in, num_layers=num_layers)  # noqa: E731
            self.critic_encoder, self.target_encoder = enc().to(dev), enc().to(dev)
            self.critic1 = Critic(**mk, encoder=self.critic_encoder).to(dev)
            self.critic2 = Critic(**mk, encoder=self.critic_encoder).to(dev)
            self.target1 = Critic(**mk, encoder=self.target_encoder).to(dev)
            self.target2 = Critic(**mk, encoder=self.target_encoder).to(dev)
            self.target_encoder.load_state_dict(self.critic_encoder.state_dict())
        else:
            self.critic1, self.critic2 = Critic(**mk).to(dev), Critic(**mk).to(dev)
            self.target1, self.target2 = Critic(**mk).to(dev), Critic(**mk).to(dev)
            self.target1.load_state_dict(self.critic1.state_dict())
            self.target2.load_state_dict(self.critic2.state_dict())
        for m in (self.target1, self.target2):
            for p in m.parameters():
                p.requires_grad_(False)
        actor_lr = lr if actor_lr is None else actor_lr
        critic_lr = lr if critic_lr is None else critic_lr
        alpha_lr = lr if alpha_lr is None else alpha_lr
        self._actor_params = list(self.actor.parameters())
        self.actor_opt = torch.optim.Adam(self._actor_params, lr=actor_lr, **cap)
        if share_critic_encoder:
            # exactly the reference's list (:119-123): the critics' input LayerNorms are not optimised
            self.critic_params = (list(self.critic_encoder.parameters()) + list(self.critic1.edge_mlp.parameters())
                                  + list(self.critic2.edge_mlp.parameters()))
        else:
            self.critic_params = list(self.critic1.parameters()) + list(self.critic2.parameters())
        self.critic_opt = torch.optim.Adam(self.critic_params, lr=critic_lr, **cap)
        self.log_alpha = torch.tensor(float(np.log(max(alpha_init, 1e-8))), requires_grad=True, device=dev)
        self.alpha_opt = torch.optim.Adam([self.log_alpha], lr=alpha_lr, **cap)
        self.gamma, self.target_tau = gamma, target_tau
        self.target_entropy, self.target_entropy_ratio, self.grad_clip = target_entropy, target_entropy_ratio, grad_clip

    @property
    def alpha(self):
        return self.log_alpha.exp()

    @torch.no_grad()
    def select_action(self, node_x, edge_attr, action_mask, deterministic: bool = False,
                      generator: Optional[torch.Generator] = None) -> SACOutput:
        """Batched ``select_action`` (:147-157): one action per scenario."""
        _, probs, _ = self.actor(node_x, edge_attr, action_mask)
        action = torch.argmax(probs, dim=1) if deterministic else torch.multinomial(probs, 1, generator=generator).squeeze(1)
        log_prob = torch.log(probs.gather(1, action.unsqueeze(1)).squeeze(1) + 1e-8)
        return SACOutput(action=action, log_prob=log_prob, probs=probs)

    # ---- one learner step, in four stages separated by the (optional) gradient all-reduces -----------
    # Each stage is a plain function of the tensors in ``s``; ``update`` runs them eagerly, or - with
    # ``cuda_graph=True`` and a fixed batch shape - replays one captured CUDA graph per stage (about a thousand
    # small launches per update otherwise: the update is launch-bound, 13 ms of GPU work in 18 ms of wall clock).
    def _stage_critic(self, s) -> None:
        with torch.no_grad():
            _, next_probs, _ = self.actor(s["next_node_x"], s["next_edge_attr"], s["next_action_mask"])
            q_next = torch.min(self.target1(s["next_node_x"], s["next_edge_attr"]),
                               self.target2(s["next_node_x"], s["next_edge_attr"]))
            v_next = (next_probs * (q_next - self.alpha * torch.log(next_probs + 1e-8))).sum(dim=1)
            target = s["reward"] + (1.0 - s["done"]) * self.gamma * v_next
        q1_all, q2_all = self.critic1(s["node_x"], s["edge_attr"]), self.critic2(s["node_x"], s["edge_attr"])
        q1, q2 = q1_all.gather(1, s["action"]).squeeze(1), q2_all.gather(1, s["action"]).squeeze(1)
        s["td_errors"] = (target - q1).detach().abs()
        critic_loss = (s["w"] * (F.mse_loss(q1, target, reduction="none") + F.mse_loss(q2, target, reduction="none"))).mean()
        self.critic_opt.zero_grad(set_to_none=not s["static"])
        critic_loss.backward()
        s["critic_loss"] = critic_loss.detach()
        s["q_all"] = torch.min(q1_all, q2_all).detach()
        s["q_taken"] = torch.min(q1, q2).detach().mean()

    def _stage_actor(self, s) -> None:
        if self.grad_clip:
            torch.nn.utils.clip_grad_norm_(self.critic_params, max_norm=self.grad_clip)
        self.critic_opt.step()
        _, probs, _ = self.actor(s["node_x"], s["edge_attr"], s["action_mask"])
        log_p = torch.log(probs + 1e-8)
        actor_loss = (probs * (self.alpha.detach() * log_p - s["q_all"])).sum(dim=1).mean()
        self.actor_opt.zero_grad(set_to_none=not s["static"])
        actor_loss.backward()
        s["actor_loss"] = actor_loss.detach()
        s["probs"], s["log_p"] = probs.detach(), log_p.detach()

    def _stage_alpha(self, s) -> None:
        if self.grad_clip:
            torch.nn.utils.clip_grad_norm_(self._actor_params, max_norm=self.grad_clip)
        self.actor_opt.step()
        if self.target_entropy is None:
            target_entropy = (self.target_entropy_ratio * torch.log((s["action_mask"] > 0).float().sum(dim=1) + 1e-8)).mean()
        else:
            target_entropy = self.target_entropy
        alpha_loss = -(self.log_alpha * (s["probs"] * (s["log_p"] + target_entropy)).sum(dim=1)).mean()
        self.alpha_opt.zero_grad(set_to_none=not s["static"])
        alpha_loss.backward()
        s["alpha_loss"] = alpha_loss.detach()

    def _stage_finish(self, s) -> None:
        if self.grad_clip:
            torch.nn.utils.clip_grad_norm_([self.log_alpha], max_norm=self.grad_clip)
        self.alpha_opt.step()
        with torch.no_grad():
            if s["alpha_max"] is not None:
                self.log_alpha.clamp_(max=float(np.log(s["alpha_max"])))
            self.log_alpha.clamp_(min=float(np.log(0.01)))
            if self.share_critic_encoder:
                self._soft_update(self.critic_encoder, self.target_encoder)
                self._soft_update(self.critic1.edge_mlp, self.target1.edge_mlp)
                self._soft_update(self.critic2.edge_mlp, self.target2.edge_mlp)
            else:
                self._soft_update(self.critic1, self.target1)
                self._soft_update(self.critic2, self.target2)
            s["alpha"] = self.alpha.detach()
            s["policy_entropy"] = -(s["probs"] * s["log_p"]).sum(dim=1).mean()
            s["q_mean"] = s["q_all"].mean()

    _STAGES = ("_stage_critic", "_stage_actor", "_stage_alpha", "_stage_finish")
    _OUT_KEYS = ("critic_loss", "actor_loss", "alpha", "alpha_loss", "policy_entropy", "q_taken", "q_mean", "td_errors")
    _IN_KEYS = ("node_x", "edge_attr", "action_mask", "next_node_x", "next_edge_attr", "next_action_mask", "reward", "done")

    def _reduce_after(self, stage: int) -> None:
        """Gradient averaging over ranks that follows stage ``stage`` (eager NCCL, between the stage graphs)."""
        if stage == 0:
            allreduce_mean_gradients(self.critic_params)
        elif stage == 1:
            allreduce_mean_gradients(self._actor_params)
        elif stage == 2:
            allreduce_mean_gradients([self.log_alpha])

    def update(self, batch: Dict[str, torch.Tensor], weights: Optional[torch.Tensor] = None,
               alpha_max: Optional[float] = None) -> Dict[str, torch.Tensor]:
        """One learner step on a dense batch (keys of ``DeviceReplay.store``); losses as in :159-263."""
        f32 = lambda k: batch[k].to(self.device, torch.float32)  # noqa: E731
        s = {k: f32(k) for k in self._IN_KEYS}
        s["action"] = batch["action"].to(self.device, torch.int64).unsqueeze(1)
        s["w"] = (torch.ones_like(s["reward"]) if weights is None
                  else weights.to(self.device, torch.float32).expand_as(s["reward"]))
        s["alpha_max"], s["static"] = alpha_max, False
        if self.cuda_graph and self.device.type == "cuda":
            out = self._update_graphed(s)
            if out is not None:
                return out
        # an eager update after the capture (a batch of another shape) must not move the gradient tensors the
        # graphs write to: zero them in place instead of dropping them
        s["static"] = self._graph is not None
        for i, name in enumerate(self._STAGES):
            getattr(self, name)(s)
            self._reduce_after(i)
        return {k: s[k] for k in self._OUT_KEYS}

    def _update_graphed(self, s) -> Optional[Dict[str, torch.Tensor]]:
        """Replay of the captured stage graphs.  The first ``_GRAPH_WARMUP`` updates run eagerly (they are real
        updates and create the optimiser state and the gradient tensors); then the four stages are captured once
        for this batch shape - gradients are zeroed in place from then on, so every tensor a graph touches keeps its
        address - and every later update copies the batch into the static inputs and replays.  A batch of another
        shape (or another ``alpha_max``) runs eagerly."""
        g = self._graph
        key = (tuple(s["node_x"].shape), tuple(s["edge_attr"].shape), s["alpha_max"])
        if g is None:
            self._graph_calls += 1
            if self._graph_calls <= self._GRAPH_WARMUP:
                return None
            static = {k: s[k].clone() for k in self._IN_KEYS + ("action", "w")}
            static["alpha_max"], static["static"] = s["alpha_max"], True
            graphs, pool = [], None
            torch.cuda.synchronize(self.device)
            try:
                for name in self._STAGES:
                    cg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(cg, pool=pool):
                        getattr(self, name)(static)
                    pool = cg.pool()
                    graphs.append(cg)
            except Exception as exc:   # capture is an optimisation of the torch side: say so and keep training eagerly
                import warnings
                warnings.warn(f"DiscreteSAC: CUDA-graph capture of the update failed ({exc}); updates stay eager")
                self.cuda_graph = False
                torch.cuda.synchronize(self.device)
                return None
            g = self._graph = {"key": key, "static": static, "graphs": graphs}
        if key != g["key"]:
            return None
        st = g["static"]
        for k in self._IN_KEYS + ("action", "w"):
            st[k].copy_(s[k])
        for i, cg in enumerate(g["graphs"]):
            cg.replay()
            self._reduce_after(i)
        out = {k: st[k] for k in self._OUT_KEYS}
        out["td_errors"] = st["td_errors"].clone()   # the caller keeps it past the next replay (priority update)
        return out

    @torch.no_grad()
    def _soft_update(self, src: nn.Module, tgt: nn.Module) -> None:
        # tp <- (1 - tau) tp + tau p (reference :289-291), one fused multi-tensor launch
        torch._foreach_lerp_(list(tgt.parameters()), list(src.parameters()), self.target_tau)

    # ---- checkpoints in the reference's layout (:265-287) ---------------------------------------
    def save(self, path: str) -> None:
        torch.save({"actor": self.actor.state_dict(), "critic1": self.critic1.state_dict(),
                    "critic2": self.critic2.state_dict(), "target1": self.target1.state_dict(),
                    "target2": self.target2.state_dict(), "log_alpha": self.log_alpha.detach().cpu()}, path)

    def load(self, path: str, map_location="cpu", strict: bool = True) -> Dict[str, Dict[str, list]]:
        """Load a checkpoint in the reference's file layout (src/rl/sac.py:278-287).  ``strict=False``
        accepts checkpoints of older revisions of the reference (the shipped
        ``history-data/outputs1/model_best.pt`` predates the input / per-layer LayerNorms): returns, per
        module, the keys the checkpoint did not provide (``missing``) and the ones it had in excess
        (``unexpected``) so the caller can assert what was consumed."""
        state = torch.load(path, map_location=map_location, weights_only=True)
        report = {}
        for name in ("actor", "critic1", "critic2", "target1", "target2"):
            res = getattr(self, name).load_state_dict(state[name], strict=strict)
            report[name] = {"missing": list(res.missing_keys), "unexpected": list(res.unexpected_keys)}
        with torch.no_grad():
            self.log_alpha.copy_(state["log_alpha"].to(self.device))
        return report
