"""Drop-in ``RepairEnv`` - the reference's single-scenario API on top of the B200 kernels.

Mirrors ``src/env/repair_env.py`` of the reference: same constructor keywords (:23-50), same
methods (``reset`` :167, ``step`` :207, ``get_state`` :751, ``compute_flow_assignment`` :299,
``compute_reward_with_goal`` :244, ``is_goal_complete`` :293, ``set_goal`` :296,
``compute_travel_time`` :667, ``compute_tstt`` :724), same public attributes (``tstt``, ``flow``,
``is_damaged``, ``capacities``, ``initial_capacities``, ``goal_mask``, ``edge_index``,
``edge_id_map``, ``betweenness_vec``, ``rng``, ``initial_tstt``, ``total_demand``, ``od_paths``,
``od_path_flows``, ``unassigned_demand`` ...), same error behaviour.  Host numpy arrays are the
source of truth between calls (callers such as the greedy baseline poke ``is_damaged`` /
``capacities`` / ``flow`` directly, reference src/baselines/__init__.py:43-67); every assignment
uploads them, runs the fused kernel with B = 1 and downloads flow / TSTT.

Differences, all forced by the north star ("no multi-backend dispatch, no CPU fallback"):
``use_cugraph`` / ``use_torch`` / ``sp_backend`` / ``force_gpu_sp`` are accepted and ignored,
``device`` only selects the CUDA ordinal, ``assignment_method='gp'`` raises
``NotImplementedError``.  Extra keywords: ``precision`` ("fp32" keeps the reference's dtypes,
"fp64" keeps link state in double) and ``line_search_iters`` (for ``assignment_method='fw_ls'``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np
import torch

from .batched_env import AssignmentEngine, METHODS
from .network import Network


@dataclass
class EnvState:
    node_features: np.ndarray
    edge_features: np.ndarray
    edge_index: np.ndarray
    action_mask: np.ndarray
    log_tstt: float
    goal_mask: np.ndarray


class RepairEnv:
    def __init__(
        self,
        graph_data,
        damaged_ratio: float = 0.3,
        bpr_alpha: float = 0.15,
        bpr_beta: float = 4.0,
        assignment_iters: int = 20,
        assignment_method: str = "msa",
        use_cugraph: bool = False,
        use_torch: bool = False,
        device: str = "cuda",
        sp_backend: str = "auto",
        force_gpu_sp: bool = False,
        reward_mode: str = "log_delta",
        reward_alpha: float = 1.0,
        reward_beta: float = 10.0,
        reward_gamma: float = 0.1,
        reward_clip: float = 0.0,
        capacity_damage: float = 1e-3,
        unassigned_penalty: float = 2e7,
        gp_step: float = 1.0,
        gp_keep_paths: int = 3,
        debug_reward: bool = False,
        debug_reward_every: int = 0,
        fixed_damage: bool = False,
        fixed_damage_seed: int | None = None,
        seed: int = 0,
        precision: str = "fp32",
        line_search_iters: int = 24,
        specialize_kernel: bool = False,
    ):
        self.graph_data = graph_data
        self.bpr_alpha, self.bpr_beta = bpr_alpha, bpr_beta
        self.assignment_iters = assignment_iters
        self.assignment_method = assignment_method.lower()
        if self.assignment_method == "gp":
            raise NotImplementedError("assignment_method 'gp' keeps per-OD path sets on the host (reference "
                                      "src/env/repair_env.py:352-419) and is out of scope of the B200 path")
        if self.assignment_method not in METHODS:
            raise ValueError(f"unknown assignment_method {assignment_method!r}")
        # accepted for signature compatibility, ignored (single CUDA backend)
        self.use_cugraph, self.use_torch, self.force_gpu_sp = use_cugraph, use_torch, bool(force_gpu_sp)
        self.sp_backend = (sp_backend or "auto").lower()
        dev = str(device)
        self.device = "cuda" if not dev.startswith("cuda") else dev
        self.reward_mode, self.reward_alpha, self.reward_beta = reward_mode, reward_alpha, reward_beta
        self.reward_gamma, self.reward_clip = reward_gamma, reward_clip
        self.capacity_damage, self.unassigned_penalty = capacity_damage, unassigned_penalty
        self.gp_step, self.gp_keep_paths = float(gp_step), int(gp_keep_paths)
        self.debug_reward, self.debug_reward_every, self._debug_step = debug_reward, debug_reward_every, 0
        self.rng = np.random.default_rng(seed)
        self.fixed_damage, self.fixed_damage_seed = bool(fixed_damage), fixed_damage_seed
        self._fixed_damaged_indices = None
        self._fixed_damage_rng = np.random.default_rng(fixed_damage_seed) if fixed_damage_seed is not None else None

        # specialize_kernel: compile (once, cached) a straight-line kernel for this topology (small networks)
        self.network = graph_data if isinstance(graph_data, Network) else Network(graph_data, specialize=specialize_kernel)
        net = self.network
        self.num_nodes, self.edges, self.num_edges = net.num_nodes, net.edges, net.num_edges
        self.edge_index = net.edge_index
        self.initial_capacities = net.initial_capacities
        self.capacities = self.initial_capacities.copy()
        self.t0 = net.t0
        self.max_capacity, self.max_t0 = net.max_capacity, net.max_t0
        self.edge_id_map = net.edge_id_map
        self.initial_tstt = None
        self.total_demand = net.total_demand
        self.unassigned_demand = 0.0
        self.relative_gap = float("inf")
        self.is_reset = True
        self.od_paths: Dict = {}
        self.od_path_flows: Dict = {}

        self._engine = AssignmentEngine(net, self.device, precision, bpr_alpha, bpr_beta, unassigned_penalty,
                                        line_search_iters)
        self._real = self._engine.real
        d, E = self._engine.device, self.num_edges
        self._d_dmg = torch.zeros((1, E), dtype=torch.uint8, device=d)
        self._d_goal = torch.zeros((1, E), dtype=torch.uint8, device=d)
        self._d_flow = torch.zeros((1, E), dtype=self._real, device=d)
        self._d_tstt = torch.zeros((1,), dtype=torch.float64, device=d)
        self._out = dict(tstt=self._d_tstt, gap=torch.zeros((1,), dtype=torch.float64, device=d),
                         unassigned=torch.zeros((1,), dtype=torch.float64, device=d),
                         status=torch.zeros((1,), dtype=torch.int32, device=d))
        self._d_dprev = torch.zeros((1, E), dtype=self._real, device=d)

        print(f"[RepairEnv] Initialized with backend=b200, device={self._engine.device}, "
              f"method={self.assignment_method}, iters={self.assignment_iters}")
        self._init_betweenness()
        self.reset(damaged_ratio=damaged_ratio)

    # ------------------------------------------------------------------------------------------
    def _init_betweenness(self) -> None:
        """Static betweenness of the undamaged network (reference :163-165)."""
        E = self.num_edges
        zeros = torch.zeros((1, E), dtype=torch.uint8, device=self._engine.device)
        one = torch.ones((1,), dtype=torch.float64, device=self._engine.device)
        _, _, raw = self._engine.state_features(self._d_flow, zeros, zeros, one, self.capacity_damage,
                                                return_raw_betweenness=True)
        self.betweenness_vec = raw[0].cpu().numpy()
        self.betweenness = {i: float(self.betweenness_vec[i]) for i in range(self.num_nodes)}

    def _upload(self) -> None:
        self._d_dmg.copy_(torch.from_numpy((np.asarray(self.is_damaged) > 0.5).astype(np.uint8)[None, :]))
        flow = np.asarray(self.flow, dtype=np.float32 if self._real == torch.float32 else np.float64)
        self._d_flow.copy_(torch.from_numpy(np.ascontiguousarray(flow)[None, :]))

    # ------------------------------------------------------------------------------------------
    def reset(self, damaged_ratio: float = 0.3) -> EnvState:
        self.is_damaged = np.zeros(self.num_edges, dtype=np.float32)
        damaged_indices = self._fixed_damaged_indices if self.fixed_damage else None
        if damaged_indices is None:
            rng = self._fixed_damage_rng if self.fixed_damage and self._fixed_damage_rng is not None else self.rng
            damaged_indices = self.network.sample_damage(rng, damaged_ratio)
            if self.fixed_damage:
                self._fixed_damaged_indices = damaged_indices
        self.is_damaged[damaged_indices] = 1.0
        self.capacities = self.initial_capacities.copy()
        self.capacities[damaged_indices] = self.capacity_damage
        self.goal_mask = self.is_damaged.copy()
        self.flow = np.zeros(self.num_edges, dtype=np.float32)
        self.od_paths, self.od_path_flows = {}, {}
        self.tstt = None
        self.is_reset = True
        self.compute_flow_assignment()
        self.initial_tstt = self.tstt
        self.is_reset = False
        return self.get_state()

    def step(self, action_edge_id: int) -> Tuple[EnvState, float, bool, Dict]:
        if action_edge_id < 0 or action_edge_id >= self.num_edges:
            raise ValueError(f"action_edge_id {action_edge_id} out of range (0..{self.num_edges - 1})")
        if self.is_damaged[action_edge_id] == 0:
            return self.get_state(), -1.0, False, {"tstt": self.tstt}
        prev_tstt = self.tstt
        self.is_damaged[action_edge_id] = 0.0
        self.capacities[action_edge_id] = self.initial_capacities[action_edge_id]
        self.compute_flow_assignment()
        reward = self.compute_reward_with_goal(
            prev_tstt, self.tstt, self.goal_mask, self.is_damaged, alpha=self.reward_alpha, beta=self.reward_beta,
            gamma=self.reward_gamma, mode=self.reward_mode, clip=self.reward_clip)
        if self.debug_reward:
            self._debug_step += 1
            if self.debug_reward_every <= 0 or self._debug_step % self.debug_reward_every == 0:
                print(f"[reward_debug] prev={prev_tstt:.6g} curr={self.tstt:.6g} diff={prev_tstt - self.tstt:.6g} "
                      f"reward={reward:.6g}")
        done = self.is_goal_complete(self.goal_mask, self.is_damaged)
        return self.get_state(), reward, bool(done), {"tstt": self.tstt}

    def compute_reward(self, prev_tstt, curr_tstt, alpha=1.0, beta=10.0, gamma=0.1) -> float:
        bonus = beta if self.is_damaged.sum() == 0 else 0.0
        return alpha * (prev_tstt - curr_tstt) + bonus - gamma

    def compute_reward_with_goal(self, prev_tstt, curr_tstt, goal_mask, damaged_mask, alpha=1.0, beta=10.0,
                                 gamma=0.1, mode="delta", clip=0.0) -> float:
        """Scalar reward arithmetic in Python floats, as in the reference (:244-291)."""
        bonus = beta if self.is_goal_complete(goal_mask, damaged_mask) else 0.0
        if mode in ("minimize_tstt", "rel_improve"):
            base = self.initial_tstt if self.initial_tstt is not None else prev_tstt
            ratio = curr_tstt / max(base, 1.0)
            if mode == "minimize_tstt":
                reward = -alpha * ratio
            else:
                reward = alpha * (((prev_tstt - curr_tstt) / max(base, 1.0)) * 100.0) - 1.0 * ratio
            reward = reward + bonus
        else:
            if mode == "neg_tstt":
                delta = -curr_tstt
            elif mode == "log_delta":
                delta = np.log10(max(prev_tstt, 1.0)) - np.log10(max(curr_tstt, 1.0))
            else:
                delta = prev_tstt - curr_tstt
            reward = alpha * delta + bonus - gamma
        if clip and clip > 0:
            reward = float(np.clip(reward, -clip, clip))
        return reward

    def is_goal_complete(self, goal_mask, damaged_mask) -> bool:
        return bool(np.sum(goal_mask * damaged_mask) == 0.0)

    def set_goal(self, goal_mask) -> None:
        self.goal_mask = goal_mask.astype(np.float32)

    # ------------------------------------------------------------------------------------------
    def compute_flow_assignment(self) -> None:
        if self.assignment_iters <= 0:
            raise ValueError("assignment_iters must be > 0 to update TSTT.")
        if self.flow is None or self.is_reset:
            self.flow = np.zeros(self.num_edges, dtype=np.float32)
        self._upload()
        self._engine.assign(self._d_dmg, self._d_flow, self.assignment_iters, self.assignment_method, self._out,
                            self._d_dprev)
        host = torch.cat([self._d_flow[0].double(), self._out["tstt"], self._out["gap"], self._out["unassigned"],
                          self._out["status"].double()]).cpu().numpy()
        E = self.num_edges
        self.flow = host[:E].astype(np.float32) if self._real == torch.float32 else host[:E].copy()
        self.tstt = float(host[E])
        self.relative_gap = float(host[E + 1])
        self.unassigned_demand = float(host[E + 2])
        self.assign_status = int(host[E + 3])
        if self.assign_status & 1:
            print("[RepairEnv] Error: Flow became NaN during assignment")

    def compute_travel_time(self, flow) -> np.ndarray:
        dtype = np.float32 if self._real == torch.float32 else np.float64
        f = torch.from_numpy(np.ascontiguousarray(np.asarray(flow, dtype=dtype))[None, :]).to(self._engine.device)
        dmg = torch.from_numpy((np.asarray(self.is_damaged) > 0.5).astype(np.uint8)[None, :]).to(self._engine.device)
        return self._engine.bpr(f, dmg)[0].cpu().numpy()

    def compute_tstt(self, flow, t, unassigned_demand: float = 0.0) -> float:
        base = float(np.sum(np.asarray(flow, np.float32) * np.asarray(t, np.float32)))
        total = max(self.total_demand, 1.0)
        out = base / total
        if unassigned_demand > 0:
            out += float(self.unassigned_penalty) * (float(unassigned_demand) / total)
        return out

    # ------------------------------------------------------------------------------------------
    def get_state(self) -> EnvState:
        self._d_dmg.copy_(torch.from_numpy((np.asarray(self.is_damaged) > 0.5).astype(np.uint8)[None, :]))
        self._d_goal.copy_(torch.from_numpy((np.asarray(self.goal_mask) > 0.5).astype(np.uint8)[None, :]))
        flow = np.asarray(self.flow, dtype=np.float32 if self._real == torch.float32 else np.float64)
        self._d_flow.copy_(torch.from_numpy(np.ascontiguousarray(flow)[None, :]))
        current = self.tstt if self.tstt is not None else self.initial_tstt
        self._d_tstt.fill_(float(current) if current is not None else 1.0)
        node, edge = self._engine.state_features(self._d_flow, self._d_dmg, self._d_goal, self._d_tstt,
                                                 self.capacity_damage)
        log_tstt = float(np.log10(max(current, 1.0))) if current is not None else 0.0
        return EnvState(
            node_features=node[0].cpu().numpy(),
            edge_features=edge[0].cpu().numpy(),
            edge_index=self.edge_index,
            action_mask=self.is_damaged.astype(np.float32),
            log_tstt=log_tstt,
            goal_mask=self.goal_mask.copy(),
        )
