"""Prioritised replay resident in HBM.

Reference: ``ReplayBuffer`` in src/train.py:27-91 - a Python sum tree over a list of pickled
``EnvState`` objects, one ``add`` per env step, and per-item ``state_to_data`` /
``Batch.from_data_list`` at sampling time (:804-861).  Every scenario shares the topology, so a
transition is a fixed-size record and the buffer is a set of preallocated ``[capacity, ...]``
tensors: adding a batch is one indexed copy per field, a sampled batch is a gather, batch assembly
is a reshape.  Semantics kept from the reference:

* priority ``p = (|td| + eps) ** alpha``; new items get the running maximum priority, and - like the
  reference's ``add`` (:50-58), which passes the maximum through ``abs(.) + eps`` and takes the maximum
  again - every default-priority insert raises the running maximum by ``eps``;
* sampling proportional to ``p`` with replacement (the sum-tree descent of :63-79 draws
  ``batch_size`` independent uniforms), importance weights ``(size * p / total) ** -beta``
  normalised by their maximum (:80-82);
* ``update_priorities`` with the learner's TD errors (:86-91): the reference loops over the sampled
  indices in order, so for an index drawn twice the LAST occurrence wins - reproduced
  deterministically (stable sort, keep the last of each run), never by a racing scatter;
* hindsight relabelling of a sampled item (:808-830): goal := links already repaired in the next
  state, reward recomputed for that goal, done := goal complete, goal column of the edge
  features overwritten.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


class DeviceReplay:
    FIELDS = ("node_x", "edge_attr", "action_mask", "next_node_x", "next_edge_attr", "next_action_mask")

    def __init__(self, capacity: int, num_nodes: int, num_edges: int, node_dim: int = 4, edge_dim: int = 6,
                 alpha: float = 0.6, beta: float = 0.4, eps: float = 1e-6, device="cuda",
                 feature_dtype: torch.dtype = torch.float32):
        self.capacity, self.N, self.E = int(capacity), num_nodes, num_edges
        self.alpha, self.beta, self.eps = float(alpha), float(beta), float(eps)
        dev = torch.device(device)
        self.device = dev
        fd = feature_dtype
        self.store: Dict[str, torch.Tensor] = {
            "node_x": torch.zeros((capacity, num_nodes, node_dim), dtype=fd, device=dev),
            "edge_attr": torch.zeros((capacity, num_edges, edge_dim), dtype=fd, device=dev),
            "action_mask": torch.zeros((capacity, num_edges), dtype=torch.uint8, device=dev),
            "next_node_x": torch.zeros((capacity, num_nodes, node_dim), dtype=fd, device=dev),
            "next_edge_attr": torch.zeros((capacity, num_edges, edge_dim), dtype=fd, device=dev),
            "next_action_mask": torch.zeros((capacity, num_edges), dtype=torch.uint8, device=dev),
            "action": torch.zeros((capacity,), dtype=torch.int64, device=dev),
            "reward": torch.zeros((capacity,), dtype=torch.float32, device=dev),
            "done": torch.zeros((capacity,), dtype=torch.float32, device=dev),
            "prev_tstt": torch.zeros((capacity,), dtype=torch.float64, device=dev),
            "next_tstt": torch.zeros((capacity,), dtype=torch.float64, device=dev),
            "initial_tstt": torch.zeros((capacity,), dtype=torch.float64, device=dev),
        }
        self.priority = torch.zeros((capacity,), dtype=torch.float32, device=dev)  # already ** alpha
        self._max_priority = torch.ones((), dtype=torch.float64, device=dev)    # on the device: no sync per update
        self.ptr, self.size = 0, 0

    @property
    def max_priority(self) -> float:
        """Running maximum of ``|td| + eps`` (reference attribute of the same name); reading it syncs."""
        return float(self._max_priority)

    def bytes_per_transition(self) -> int:
        return sum(t.element_size() * (t.numel() // self.capacity) for t in self.store.values()) + 4

    # ---- add ------------------------------------------------------------------------------------
    def add_batch(self, batch: Dict[str, torch.Tensor], priority: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Append ``n`` transitions (dict of ``[n, ...]`` tensors with the keys of ``self.store``);
        returns the slots written.  Wraps around like the reference's ring (:55-58)."""
        n = int(batch["action"].shape[0])
        if n == 0:
            return torch.empty(0, dtype=torch.int64, device=self.device)
        if n > self.capacity:
            batch = {k: v[-self.capacity:] for k, v in batch.items()}
            n = self.capacity
        slots = (torch.arange(n, device=self.device) + self.ptr) % self.capacity
        for k, dst in self.store.items():
            dst[slots] = batch[k].to(device=self.device, dtype=dst.dtype)
        if priority is None:
            # n sequential reference adds: item i enters at max + (i + 1) eps, which becomes the new maximum
            steps = torch.arange(1, n + 1, dtype=torch.float64, device=self.device)
            p = ((self._max_priority + steps * self.eps) ** self.alpha).float()
            self._max_priority = self._max_priority + n * self.eps
        else:
            pr = priority.abs().double().to(self.device)[-n:] + self.eps
            self._max_priority = torch.maximum(self._max_priority, pr.max())
            p = (pr ** self.alpha).float()
        self.priority[slots] = p
        self.ptr = (self.ptr + n) % self.capacity
        self.size = min(self.size + n, self.capacity)
        return slots

    # ---- sample ---------------------------------------------------------------------------------
    def sample(self, batch_size: int, generator: Optional[torch.Generator] = None):
        if self.size == 0:
            raise ValueError("Cannot sample from an empty replay buffer.")
        pr = self.priority[: self.size]
        total = pr.sum()
        if float(total) <= 0:
            raise ValueError("Cannot sample from an empty replay buffer.")
        idx = torch.multinomial(pr, batch_size, replacement=True, generator=generator)
        probs = pr[idx] / total
        weights = (self.size * probs) ** (-self.beta)
        weights = weights / weights.max().clamp(min=1e-30)
        batch = {k: v[idx] for k, v in self.store.items()}
        return batch, idx, weights.float()

    def update_priorities(self, idx: torch.Tensor, td_errors: torch.Tensor) -> None:
        if idx.numel() == 0:
            return
        pr = td_errors.detach().abs().double().to(self.device) + self.eps
        self._max_priority = torch.maximum(self._max_priority, pr.max())
        # an index sampled more than once: the reference's sequential loop leaves the LAST one in place
        sidx, perm = torch.sort(idx.to(self.device), stable=True)
        last = torch.ones_like(sidx, dtype=torch.bool)
        last[:-1] = sidx[1:] != sidx[:-1]
        self.priority[sidx[last]] = (pr[perm[last]] ** self.alpha).float()

    # ---- hindsight relabelling ------------------------------------------------------------------
    @staticmethod
    def relabel_hindsight(batch: Dict[str, torch.Tensor], her_ratio: float, reward_fn, reward_scale: float = 1.0,
                          generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
        """With probability ``her_ratio`` per item: goal := 1 - next_action_mask (what has been
        achieved), reward := reward_fn(prev_tstt, next_tstt, initial_tstt, next_action_mask, goal) *
        reward_scale, done := no goal link still damaged, and - exactly like the reference's
        ``apply_goal`` (src/train.py:125-135) - the LAST edge-feature column of both states is
        overwritten with the goal."""
        if her_ratio <= 0:
            return batch
        n = batch["action"].shape[0]
        dev = batch["action"].device
        pick = torch.rand(n, device=dev, generator=generator) < her_ratio
        if not bool(pick.any()):
            return batch
        goal = (1 - batch["next_action_mask"]).to(torch.uint8)
        r, d = reward_fn(batch["prev_tstt"], batch["next_tstt"], batch["initial_tstt"], batch["next_action_mask"], goal)
        out = dict(batch)
        out["reward"] = torch.where(pick, (r * reward_scale).float(), batch["reward"])
        out["done"] = torch.where(pick, d.float(), batch["done"])
        for key in ("edge_attr", "next_edge_attr"):
            ea = batch[key].clone()
            ea[:, :, -1] = torch.where(pick.unsqueeze(1), goal.to(ea.dtype), ea[:, :, -1])
            out[key] = ea
        return out
