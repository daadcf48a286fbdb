"""Batched repair environment: thousands of independent damage/repair scenarios per launch.

New API (the reference owns exactly one env per process).  Semantics per scenario are those of
``RepairEnv`` (reference: src/env/repair_env.py): ``reset`` (:167-205), ``step`` (:207-237),
``compute_flow_assignment`` (:299-345), ``get_state`` (:751-819).  All arrays are torch CUDA
tensors, scenario-major ``[B, E]``; PyTorch is only the owner of device memory and streams - the
arithmetic runs in the sm_100a kernels behind ``include/b200_assign.h``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _native
from .network import Network

METHODS = {"msa": 0, "fw": 1, "cfw": 2, "fw_ls": 3}
REWARD_MODES = {"delta": 0, "log_delta": 1, "neg_tstt": 2, "minimize_tstt": 3, "rel_improve": 4}
PRECISIONS = {"fp32": 0, "float32": 0, "f32": 0, "fp64": 1, "float64": 1, "f64": 1}


def _require_cuda(device: torch.device) -> None:
    if device.type != "cuda" or not torch.cuda.is_available():
        raise _native.NativeLibraryError(
            "the assignment path runs only on a CUDA device (sm_100a); there is no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


@dataclass
class BatchedEnvState:
    node_features: torch.Tensor  # [B, N, 4] fp32
    edge_features: torch.Tensor  # [B, E, 6] fp32
    edge_index: torch.Tensor  # [2, E] int64 (shared by all scenarios)
    action_mask: torch.Tensor  # [B, E] fp32
    log_tstt: torch.Tensor  # [B] fp64
    goal_mask: torch.Tensor  # [B, E] fp32


class AssignmentEngine:
    """Thin, allocation-free-per-call front end of the C ABI for one network on one device."""

    def __init__(self, network: Network, device="cuda", precision: str = "fp32", bpr_alpha: float = 0.15,
                 bpr_beta: float = 4.0, unassigned_penalty: float = 2e7, line_search_iters: int = 24):
        dev = torch.device(device)
        _require_cuda(dev)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.prec = PRECISIONS[precision]
        self.real = torch.float32 if self.prec == 0 else torch.float64
        self.net = network
        if network.device_index != (self.device.index or 0):
            if network._handle is not None:
                # the handle may be shared (another engine, a GAT module): never free it here
                raise ValueError(f"Network already lives on cuda:{network.device_index}; build another Network "
                                 f"for {self.device}")
            network.device_index = self.device.index or 0
        self.alpha, self.beta = float(bpr_alpha), float(bpr_beta)
        self.penalty, self.ls_iters = float(unassigned_penalty), int(line_search_iters)
        self.lib = _native.lib()
        with torch.cuda.device(self.device):
            self.h = network.handle
        self.N, self.E, self.O = network.num_nodes, network.num_edges, network.num_origins
        self.launches = 0

    @property
    def kernel_name(self) -> str:
        """Which fused kernel ``assign`` dispatches to for this network."""
        return self.lib.b200_assign_kernel_name(self.h).decode()

    @property
    def last_config(self) -> Dict[str, int]:
        """Launch configuration of the last ``assign`` (kernels, cluster size, warm start, warps per CTA,
        contraction-hierarchy shortest paths)."""
        import ctypes as C
        cfg = (C.c_int32 * 8)()
        _native.check(self.lib.b200_assign_last_config(self.h, cfg), "b200_assign_last_config")
        return dict(launches=cfg[0], cluster=cfg[1], warm=cfg[2], warps=cfg[3], cch=cfg[4])

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _chk(self, t: torch.Tensor, dtype, shape: Tuple[int, ...], name: str) -> None:
        if t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous() or t.device != self.device:
            raise ValueError(f"{name}: expected contiguous {dtype} {tuple(shape)} on {self.device}, got "
                             f"{t.dtype} {tuple(t.shape)} on {t.device}")

    # ---- stage calls (used by the parity tests and by callers that want the pieces) ----------
    def bpr(self, flow: torch.Tensor, dmg: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = flow.shape[0]
        self._chk(flow, self.real, (B, self.E), "flow")
        self._chk(dmg, torch.uint8, (B, self.E), "dmg")
        t = torch.empty_like(flow) if out is None else out
        _native.check(self.lib.b200_bpr(self.h, B, self.prec, _ptr(flow), _ptr(dmg), _ptr(t), self.alpha, self.beta,
                                        self._stream()), "b200_bpr")
        self.launches += 1
        return t

    def sssp(self, t: torch.Tensor):
        B = t.shape[0]
        self._chk(t, self.real, (B, self.E), "t")
        dist = torch.empty((B, self.O, self.N), dtype=torch.float64, device=self.device)
        pred = torch.empty((B, self.O, self.N), dtype=torch.int32, device=self.device)
        pedge = torch.empty((B, self.O, self.N), dtype=torch.int32, device=self.device)
        _native.check(self.lib.b200_sssp(self.h, B, self.prec, _ptr(t), _ptr(dist), _ptr(pred), _ptr(pedge),
                                         self._stream()), "b200_sssp")
        self.launches += 1
        return dist, pred, pedge

    def aon(self, pedge: torch.Tensor):
        B = pedge.shape[0]
        self._chk(pedge, torch.int32, (B, self.O, self.N), "pedge")
        aux = torch.empty((B, self.E), dtype=self.real, device=self.device)
        un = torch.empty((B,), dtype=torch.float64, device=self.device)
        _native.check(self.lib.b200_aon(self.h, B, self.prec, _ptr(pedge), _ptr(aux), _ptr(un), self._stream()),
                      "b200_aon")
        self.launches += 1
        return aux, un

    def update_reduce(self, method: str, it: int, flow, aux, t, dmg, unassigned=None, dprev=None, status=None):
        B = flow.shape[0]
        for name, x in (("flow", flow), ("aux", aux), ("t", t)):
            self._chk(x, self.real, (B, self.E), name)
        self._chk(dmg, torch.uint8, (B, self.E), "dmg")
        tstt = torch.empty((B,), dtype=torch.float64, device=self.device)
        gap = torch.empty((B,), dtype=torch.float64, device=self.device)
        _native.check(self.lib.b200_update_reduce(
            self.h, B, self.prec, METHODS[method], int(it), _ptr(flow), _ptr(aux), _ptr(t), _ptr(dmg), _ptr(dprev),
            _ptr(unassigned), _ptr(tstt), _ptr(gap), _ptr(status), self.alpha, self.beta, self.penalty, self.ls_iters,
            self._stream()), "b200_update_reduce")
        self.launches += 1
        return tstt, gap

    # ---- fused path ---------------------------------------------------------------------------
    def assign(self, dmg: torch.Tensor, flow: torch.Tensor, iters: int, method: str, out: Optional[Dict] = None,
               dprev: Optional[torch.Tensor] = None, active: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """compute_flow_assignment for every row; ``flow`` is the warm start and is overwritten.
        Rows with ``active[b] == 0`` (uint8 [B], optional) are skipped and keep all their outputs."""
        if method not in METHODS:
            raise NotImplementedError(f"assignment_method {method!r} is not supported (msa, fw, cfw, fw_ls)")
        B = dmg.shape[0]
        self._chk(dmg, torch.uint8, (B, self.E), "dmg")
        self._chk(flow, self.real, (B, self.E), "flow")
        if out is None:
            out = dict(tstt=torch.empty((B,), dtype=torch.float64, device=self.device),
                       gap=torch.empty((B,), dtype=torch.float64, device=self.device),
                       unassigned=torch.empty((B,), dtype=torch.float64, device=self.device),
                       status=torch.empty((B,), dtype=torch.int32, device=self.device))
        if method == "cfw" and dprev is None:
            dprev = torch.empty_like(flow)
        if active is not None:
            self._chk(active, torch.uint8, (B,), "active")
        _native.check(self.lib.b200_assign(
            self.h, B, self.prec, int(iters), METHODS[method], _ptr(dmg), _ptr(flow), _ptr(dprev), _ptr(out["tstt"]),
            _ptr(out["gap"]), _ptr(out["unassigned"]), _ptr(out["status"]), _ptr(active), self.alpha, self.beta,
            self.penalty, self.ls_iters, self._stream()), "b200_assign")
        self.launches += self.lib.b200_assign_launches(self.h)
        out["flow"] = flow
        return out

    def state_features(self, flow, dmg, goal, tstt, capacity_damage: float, return_raw_betweenness: bool = False):
        B = flow.shape[0]
        self._chk(flow, self.real, (B, self.E), "flow")
        self._chk(dmg, torch.uint8, (B, self.E), "dmg")
        self._chk(goal, torch.uint8, (B, self.E), "goal")
        self._chk(tstt, torch.float64, (B,), "tstt")
        node = torch.empty((B, self.N, 4), dtype=torch.float32, device=self.device)
        edge = torch.empty((B, self.E, 6), dtype=torch.float32, device=self.device)
        raw = torch.empty((B, self.N), dtype=torch.float32, device=self.device) if return_raw_betweenness else None
        _native.check(self.lib.b200_state_features(self.h, B, self.prec, _ptr(flow), _ptr(dmg), _ptr(goal), _ptr(tstt),
                                                   float(capacity_damage), _ptr(node), _ptr(edge), _ptr(raw),
                                                   self._stream()), "b200_state_features")
        self.launches += 1
        return (node, edge, raw) if return_raw_betweenness else (node, edge)


class BatchedRepairEnv:
    """B independent repair environments advanced in lock step on one GPU."""

    def __init__(self, graph_data, num_envs: int, damaged_ratio: float = 0.3, bpr_alpha: float = 0.15,
                 bpr_beta: float = 4.0, assignment_iters: int = 20, assignment_method: str = "msa",
                 device="cuda", precision: str = "fp32", reward_mode: str = "log_delta", reward_alpha: float = 1.0,
                 reward_beta: float = 10.0, reward_gamma: float = 0.1, reward_clip: float = 0.0,
                 capacity_damage: float = 1e-3, unassigned_penalty: float = 2e7, fixed_damage: bool = False,
                 fixed_damage_seed: Optional[int] = None, seed: int = 0, seeds: Optional[Sequence[int]] = None,
                 line_search_iters: int = 24, specialize_kernel: bool = False, device_sampling: bool = False,
                 sampler_seed: Optional[int] = None):
        if assignment_iters <= 0:
            raise ValueError("assignment_iters must be > 0 to update TSTT.")
        self.assignment_method = assignment_method.lower()
        if self.assignment_method == "gp":
            raise NotImplementedError("assignment_method 'gp' (path-based gradient projection, reference "
                                      "src/env/repair_env.py:352-419) is out of scope of the batched path")
        if self.assignment_method not in METHODS:
            raise ValueError(f"unknown assignment_method {assignment_method!r}")
        if reward_mode not in REWARD_MODES:
            reward_mode = "delta"  # the reference's else-branch (:285-286)
        # specialize_kernel: compile (once, cached) a straight-line kernel for this topology (small networks)
        self.network = graph_data if isinstance(graph_data, Network) else Network(graph_data, specialize=specialize_kernel)
        self.engine = AssignmentEngine(self.network, device, precision, bpr_alpha, bpr_beta, unassigned_penalty,
                                       line_search_iters)
        self.device = self.engine.device
        self.B, self.E, self.N = int(num_envs), self.network.num_edges, self.network.num_nodes
        self.assignment_iters = int(assignment_iters)
        self.reward_mode, self.reward_alpha, self.reward_beta = reward_mode, reward_alpha, reward_beta
        self.reward_gamma, self.reward_clip = reward_gamma, reward_clip
        self.capacity_damage = float(capacity_damage)
        self.damaged_ratio = damaged_ratio
        self.fixed_damage = bool(fixed_damage)
        seeds = list(seeds) if seeds is not None else [seed + b for b in range(self.B)]
        if len(seeds) != self.B:
            raise ValueError("need one seed per environment")
        self.rngs = [np.random.default_rng(s) for s in seeds]
        # fixed_damage, as in the reference (:79-81, :171-190; src/train.py:160-161 gives every rollout worker
        # the same fixed_damage_seed): with a seed, ONE scenario - the first draw of default_rng(seed) - shared
        # by every env of every rank; without one, every env keeps the first draw of its OWN generator
        # (`seeds`, which the trainer offsets by rank, so ranks do not duplicate each other).
        self.fixed_damage_seed = fixed_damage_seed
        self._fixed_damage = None
        self._fixed_damage_dev = None
        self.device_sampling = bool(device_sampling)
        self._sampler_seed = int(seeds[0]) * 0x9E3779B1 + 0x5851F42D if sampler_seed is None else int(sampler_seed)
        dev, real = self.device, self.engine.real
        self.is_damaged = torch.zeros((self.B, self.E), dtype=torch.uint8, device=dev)
        self.goal_mask = torch.zeros((self.B, self.E), dtype=torch.uint8, device=dev)
        self.flow = torch.zeros((self.B, self.E), dtype=real, device=dev)
        self._dprev = torch.empty((self.B, self.E), dtype=real, device=dev) if self.assignment_method == "cfw" else None
        self._out = dict(tstt=torch.zeros((self.B,), dtype=torch.float64, device=dev),
                         gap=torch.zeros((self.B,), dtype=torch.float64, device=dev),
                         unassigned=torch.zeros((self.B,), dtype=torch.float64, device=dev),
                         status=torch.zeros((self.B,), dtype=torch.int32, device=dev))
        self.tstt = self._out["tstt"]
        self.initial_tstt = torch.zeros((self.B,), dtype=torch.float64, device=dev)
        self._prev_tstt = torch.zeros((self.B,), dtype=torch.float64, device=dev)
        self._valid = torch.zeros((self.B,), dtype=torch.uint8, device=dev)
        self._reward = torch.zeros((self.B,), dtype=torch.float64, device=dev)
        self._done = torch.zeros((self.B,), dtype=torch.uint8, device=dev)
        self._act_status = torch.zeros((self.B,), dtype=torch.int32, device=dev)
        self._pinned_mask = torch.zeros((self.B, self.E), dtype=torch.uint8).pin_memory()
        self._sampler_counters = torch.zeros((self.B,), dtype=torch.int64, device=dev)   # device random-stream positions
        self._sampler_tries = torch.zeros((self.B,), dtype=torch.int32, device=dev)
        self.edge_index = torch.from_numpy(self.network.edge_index).to(dev)
        self.assignments_done = 0

    # ---- properties mirroring RepairEnv attribute names ---------------------------------------
    @property
    def num_envs(self) -> int:
        return self.B

    @property
    def gap(self) -> torch.Tensor:
        return self._out["gap"]

    @property
    def unassigned_demand(self) -> torch.Tensor:
        return self._out["unassigned"]

    @property
    def status(self) -> torch.Tensor:
        return self._out["status"]

    @property
    def capacities(self) -> torch.Tensor:
        cap0 = torch.from_numpy(self.network.initial_capacities).to(self.device)
        return torch.where(self.is_damaged.bool(), torch.full_like(cap0, self.capacity_damage), cap0)

    # ---- reset / step -------------------------------------------------------------------------
    def sample_damage_masks(self, damaged_ratio: float) -> np.ndarray:
        """Host draw with the reference's generator and retry rule (one numpy Generator per env): what makes
        env b reproduce ``RepairEnv(seed=seeds[b])`` draw for draw."""
        if self.fixed_damage and self._fixed_damage is not None:
            return self._fixed_damage
        masks = np.zeros((self.B, self.E), dtype=np.uint8)
        if self.fixed_damage and self.fixed_damage_seed is not None:
            masks[:, self.network.sample_damage(np.random.default_rng(self.fixed_damage_seed), damaged_ratio)] = 1
        else:
            for b in range(self.B):
                masks[b, self.network.sample_damage(self.rngs[b], damaged_ratio)] = 1
        if self.fixed_damage:
            self._fixed_damage = masks
        return masks

    def sample_damage_device(self, damaged_ratio: float, select: Optional[torch.Tensor] = None) -> None:
        """Redraw ``is_damaged`` of the rows with ``select[b]`` (all rows when None) ON THE DEVICE
        (``b200_sample_damage``: same rule as the reference - <= 50 tries for a strongly connected
        undamaged subgraph, then an unconstrained draw - from a counter-based random stream, not numpy's).
        No host synchronisation; used at episode boundaries of a lock-step batch."""
        count = max(1, int(self.E * damaged_ratio))
        eng = self.engine
        if select is not None:
            select = select.to(device=self.device, dtype=torch.uint8).contiguous()
        _native.check(eng.lib.b200_sample_damage(
            eng.h, self.B, count, 50, self._sampler_seed & 0xFFFFFFFFFFFFFFFF, _ptr(self._sampler_counters), _ptr(select),
            _ptr(self.is_damaged), _ptr(self._sampler_tries), eng._stream()), "b200_sample_damage")
        eng.launches += 1

    def reset(self, damaged_ratio: Optional[float] = None, damage_masks=None) -> BatchedEnvState:
        ratio = self.damaged_ratio if damaged_ratio is None else damaged_ratio
        if damage_masks is None and self.device_sampling and not self.fixed_damage:
            self.sample_damage_device(ratio)
        else:
            if damage_masks is None:
                damage_masks = self.sample_damage_masks(ratio)
            if isinstance(damage_masks, np.ndarray):
                self._pinned_mask.copy_(torch.from_numpy(np.ascontiguousarray(damage_masks, dtype=np.uint8)))
                self.is_damaged.copy_(self._pinned_mask, non_blocking=True)
            else:
                self.is_damaged.copy_(damage_masks.to(torch.uint8))
        self.goal_mask.copy_(self.is_damaged)
        self.flow.zero_()  # cold start (reference :198, :302-303)
        self.compute_flow_assignment()
        self.initial_tstt.copy_(self.tstt)
        return self.get_state()

    def reset_where(self, mask: torch.Tensor, damaged_ratio: Optional[float] = None) -> None:
        """Per-scenario reset (episode boundaries in a lock-step batch): scenarios with ``mask[b]``
        get a fresh damage draw (the cached draw under ``fixed_damage``, otherwise the device sampler),
        zero flows, goal := damage, a cold-start assignment and a new ``initial_tstt``; the others are
        untouched.  Everything stays on the device: no host synchronisation, no per-env Python loop (the
        reference resets one env at a time, :167-205)."""
        mask = mask.to(device=self.device, dtype=torch.bool)
        ratio = self.damaged_ratio if damaged_ratio is None else damaged_ratio
        sel = mask.to(torch.uint8)
        if self.fixed_damage:
            if self._fixed_damage_dev is None:
                self._fixed_damage_dev = torch.from_numpy(self.sample_damage_masks(ratio)).to(self.device)
            self.is_damaged.copy_(torch.where(mask.unsqueeze(1), self._fixed_damage_dev, self.is_damaged))
        else:
            self.sample_damage_device(ratio, sel)
        self.goal_mask.copy_(torch.where(mask.unsqueeze(1), self.is_damaged, self.goal_mask))
        self.flow.mul_((~mask).unsqueeze(1).to(self.flow.dtype))
        self.compute_flow_assignment(active=sel)
        self.initial_tstt.copy_(torch.where(mask, self.tstt, self.initial_tstt))

    def compute_flow_assignment(self, active: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.engine.assign(self.is_damaged, self.flow, self.assignment_iters, self.assignment_method, self._out,
                           self._dprev, active)
        self.assignments_done += self.B
        return self.tstt

    def step(self, actions, validate_actions: bool = True
             ) -> Tuple[BatchedEnvState, torch.Tensor, torch.Tensor, Dict[str, torch.Tensor]]:
        """One lock-step env step.  Out-of-range actions raise ``ValueError`` BEFORE anything is mutated,
        like the reference (:208-209): host-side actions (list / numpy / CPU tensor) are checked on the
        host at no cost; a CUDA tensor is checked with one device reduction (a stream sync) unless
        ``validate_actions=False`` - callers whose actions are in range by construction (a masked
        categorical over the E links, the trainer) pass False and keep the hot loop free of syncs; an
        out-of-range row then counts as an invalid action (reward -1, scenario untouched) and is
        reported in ``info["out_of_range"]``."""
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.shape != (self.B,):
            raise ValueError(f"actions must have shape ({self.B},)")
        if validate_actions or actions.device.type == "cpu":
            bad = (actions < 0) | (actions >= self.E)
            if bool(bad.any()):
                b = int(torch.nonzero(bad)[0])
                raise ValueError(f"action_edge_id {int(actions[b])} out of range (0..{self.E - 1})")
        actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        eng, lib = self.engine, self.engine.lib
        self._prev_tstt.copy_(self.tstt)
        self._act_status.zero_()
        _native.check(lib.b200_apply_actions(eng.h, self.B, _ptr(actions), _ptr(self.is_damaged), _ptr(self._valid),
                                             _ptr(self._act_status), eng._stream()), "b200_apply_actions")
        # Scenarios whose action was invalid keep flow/TSTT (reference :210-212): the kernel skips
        # rows whose `active` byte is 0.
        self.compute_flow_assignment(active=self._valid)
        _native.check(lib.b200_reward_done(
            eng.h, self.B, REWARD_MODES[self.reward_mode], _ptr(self._prev_tstt), _ptr(self.tstt),
            _ptr(self.initial_tstt), _ptr(self.is_damaged), _ptr(self.goal_mask), _ptr(self._valid),
            float(self.reward_alpha), float(self.reward_beta), float(self.reward_gamma), float(self.reward_clip or 0.0),
            _ptr(self._reward), _ptr(self._done), eng._stream()), "b200_reward_done")
        eng.launches += 2
        state = self.get_state()
        return state, self._reward.clone(), self._done.bool(), {"tstt": self.tstt.clone(), "valid": self._valid.bool(),
                                                                "out_of_range": self._act_status != 0}

    def get_state(self) -> BatchedEnvState:
        node, edge = self.engine.state_features(self.flow, self.is_damaged, self.goal_mask, self.tstt,
                                                self.capacity_damage)
        return BatchedEnvState(node_features=node, edge_features=edge, edge_index=self.edge_index,
                               action_mask=self.is_damaged.to(torch.float32),
                               log_tstt=torch.log10(torch.clamp(self.tstt, min=1.0)),
                               goal_mask=self.goal_mask.to(torch.float32))

    # ---- end-to-end host entry point (bench.py's e2e leg, greedy baseline) --------------------
    def assign_host(self, damage_masks: np.ndarray, out_tstt: Optional[np.ndarray] = None,
                    out_flow: Optional[np.ndarray] = None) -> np.ndarray:
        """Host buffers in, host buffers out: H2D of the uint8 masks, cold-start assignment,
        D2H of TSTT (and of the link flows when ``out_flow`` [B,E] is given - the reference leaves
        ``flow`` on the host too, :335).  ``damage_masks`` [B,E] uint8 (pinned memory recommended)."""
        src = torch.from_numpy(damage_masks) if isinstance(damage_masks, np.ndarray) else damage_masks
        self.is_damaged.copy_(src, non_blocking=True)
        self.flow.zero_()
        self.compute_flow_assignment()
        if not hasattr(self, "_pinned_tstt"):
            self._pinned_tstt = torch.empty((self.B,), dtype=torch.float64).pin_memory()
        self._pinned_tstt.copy_(self.tstt, non_blocking=True)
        if out_flow is not None:
            if not hasattr(self, "_pinned_flow"):
                self._pinned_flow = torch.empty((self.B, self.E), dtype=self.engine.real).pin_memory()
            self._pinned_flow.copy_(self.flow, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        if out_flow is not None:
            out_flow[:] = self._pinned_flow.numpy()
        res = self._pinned_tstt.numpy()
        if out_tstt is not None:
            out_tstt[:] = res
            return out_tstt
        return res.copy()
