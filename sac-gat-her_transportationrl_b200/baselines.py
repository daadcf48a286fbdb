"""Batched repair baselines and episode runner.

Reference: ``src/baselines/__init__.py`` - ``select_random`` (:11-13), ``select_max_vc`` (:16-19),
``select_max_flow`` (:22-25), ``select_max_betweenness`` (:28-32), ``select_greedy_one_step``
(:35-69), ``run_episode`` (:72-101), ``get_baseline_policies`` (:104-113).  There a policy maps
one ``EnvState`` to one link id; here it maps a ``BatchedEnvState`` to ``actions[B]``.

The one-step greedy baseline is the expensive one (22.9 s per Sioux Falls episode on the
reference's CPU path): "for every damaged link: repair it, run a full assignment
from the current flows, undo; take the link with the lowest TSTT".  Instead of save/restore on a
mutable env, all (scenario, candidate) pairs of the whole batch become ONE batch of independent
scenarios for the fused assignment kernel: ``sum_b #damaged_b`` warm-started assignments in a
single launch.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from .env.batched_env import BatchedEnvState, BatchedRepairEnv

Policy = Callable[[BatchedEnvState], torch.Tensor]


def _masked_argmax(scores: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    # the reference multiplies by the 0/1 mask and takes np.argmax (first maximum)
    return torch.argmax(scores * mask, dim=1).to(torch.int32)


def select_random(state: BatchedEnvState, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Uniform over the damaged links of each scenario (reference :11-13).  A scenario without a
    damaged link gets action 0 (its episode is over; ``step`` treats it as invalid)."""
    mask = state.action_mask
    w = mask + (mask.sum(dim=1, keepdim=True) == 0).to(mask.dtype)  # all-zero rows: uniform over everything
    return torch.multinomial(w, 1, generator=generator).squeeze(1).to(torch.int32)


def select_max_vc(state: BatchedEnvState) -> torch.Tensor:
    return _masked_argmax(state.edge_features[:, :, 2], state.action_mask)


def select_max_flow(state: BatchedEnvState) -> torch.Tensor:
    return _masked_argmax(state.edge_features[:, :, 2] * state.edge_features[:, :, 1], state.action_mask)


def select_max_betweenness(state: BatchedEnvState, node_betweenness: torch.Tensor) -> torch.Tensor:
    """node_betweenness: [N] (the reference's ``env.betweenness_vec`` of the undamaged network) or [B, N]."""
    src, dst = state.edge_index[0], state.edge_index[1]
    nb = node_betweenness if node_betweenness.dim() == 2 else node_betweenness.unsqueeze(0)
    edge_bw = (nb[:, src] + nb[:, dst]) / 2.0
    return _masked_argmax(edge_bw.expand_as(state.action_mask), state.action_mask)


def select_greedy_one_step(env: BatchedRepairEnv, state: Optional[BatchedEnvState] = None,
                           chunk: int = 1 << 16) -> torch.Tensor:
    """One-step look-ahead for every scenario of ``env`` (reference :35-69): each damaged link is
    repaired in a clone of its scenario, the clone is assigned from the scenario's current flows
    (the reference restores ``env.flow`` before every candidate), and the link with the lowest
    TSTT wins; ties go to the lowest link id (strict ``<`` in ascending candidate order).  The env
    itself is not modified."""
    dmg = env.is_damaged
    B, E = dmg.shape
    cand = torch.nonzero(dmg, as_tuple=False)  # [M, 2] rows sorted by (scenario, link)
    best = torch.zeros(B, dtype=torch.int32, device=dmg.device)
    if cand.shape[0] == 0:
        return best
    scen, link = cand[:, 0], cand[:, 1]
    M = cand.shape[0]
    tstt = torch.empty(M, dtype=torch.float64, device=dmg.device)
    for lo in range(0, M, chunk):
        hi = min(M, lo + chunk)
        s, l = scen[lo:hi], link[lo:hi]
        masks = dmg[s].clone()
        masks[torch.arange(hi - lo, device=dmg.device), l] = 0
        flows = env.flow[s].clone()
        out = env.engine.assign(masks, flows, env.assignment_iters, env.assignment_method)
        tstt[lo:hi] = out["tstt"]
    # first minimum per scenario: sort by (scenario, tstt, link) and take the head of every group
    order = torch.argsort(link, stable=True)
    order = order[torch.argsort(tstt[order], stable=True)]
    order = order[torch.argsort(scen[order], stable=True)]
    s_sorted = scen[order]
    head = torch.ones_like(s_sorted, dtype=torch.bool)
    head[1:] = s_sorted[1:] != s_sorted[:-1]
    best[s_sorted[head]] = link[order][head].to(torch.int32)
    return best


def make_actor_policy(agent, deterministic: bool = True, generator: Optional[torch.Generator] = None) -> Policy:
    """The trained policy as a batched ``Policy`` (reference src/eval.py:40-48 wraps
    ``agent.select_action`` the same way for one state): argmax of the masked link probabilities,
    or a sample from them."""
    def policy(state: BatchedEnvState) -> torch.Tensor:
        out = agent.select_action(state.node_features, state.edge_features, state.action_mask,
                                  deterministic=deterministic, generator=generator)
        return out.action.to(torch.int32)
    return policy


def evaluate(env: BatchedRepairEnv, policies: Dict[str, Policy], reward_scale: float = 1.0, max_steps: int = 0,
             damage_masks=None) -> Dict[str, Dict[str, float]]:
    """Every policy on the SAME damage scenarios (reference src/eval.py:51-86 loops over seeds and
    policies one episode at a time): mean over scenarios of the reference's episode metrics."""
    if damage_masks is None:
        damage_masks = env.sample_damage_masks(env.damaged_ratio)
    table = {}
    for name, pol in policies.items():
        res = run_episode(env, pol, reward_scale=reward_scale, max_steps=max_steps, damage_masks=damage_masks)
        table[name] = {k: float(res[k].double().mean()) for k in ("reward", "tstt_last", "tstt_mean", "tstt_auc", "steps")}
    return table


def get_baseline_policies(env: BatchedRepairEnv, node_betweenness: Optional[torch.Tensor] = None) -> Dict[str, Policy]:
    """Same names as the reference's ``get_baseline_policies`` (:104-113)."""
    pol: Dict[str, Policy] = {
        "random": select_random,
        "max_vc": select_max_vc,
        "max_flow": select_max_flow,
        "greedy": lambda s: select_greedy_one_step(env, s),
    }
    if node_betweenness is not None:
        pol["max_betweenness"] = lambda s: select_max_betweenness(s, node_betweenness)
    return pol


def run_episode(env: BatchedRepairEnv, policy: Policy, reward_scale: float = 1.0, max_steps: int = 0,
                damage_masks=None) -> Dict[str, torch.Tensor]:
    """All scenarios of ``env`` from reset to done (reference ``run_episode`` :72-101, per
    scenario).  A scenario that is done keeps its last TSTT and collects no further reward.
    Returns per-scenario tensors: ``reward`` [B], ``tstt_last`` [B], ``tstt_mean`` [B],
    ``tstt_auc`` [B] (trapezoid over the scenario's own curve, like ``np.trapz``), ``steps`` [B]
    and the padded curve ``tstt_curve`` [T, B]."""
    state = env.reset(damage_masks=damage_masks)
    B = env.num_envs
    dev = env.device
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    total = torch.zeros(B, dtype=torch.float64, device=dev)
    steps = torch.zeros(B, dtype=torch.int64, device=dev)
    curve = []
    t = 0
    while not bool(done.all()):
        actions = policy(state)
        # finished scenarios replay an undamaged link: invalid action, nothing changes (:210-212)
        state, reward, step_done, info = env.step(actions)
        live = ~done
        total += torch.where(live, reward * reward_scale, torch.zeros_like(reward))  # invalid action: -1 (:210-212)
        steps += live.to(torch.int64)
        curve.append(torch.where(live, info["tstt"], curve[-1] if curve else info["tstt"]))
        done = done | (live & step_done)
        t += 1
        if t >= (max_steps if max_steps > 0 else 10 * env.E):
            break
    c = torch.stack(curve) if curve else env.tstt.clone().unsqueeze(0)
    T = c.shape[0]
    idx = torch.arange(T, device=dev).unsqueeze(1)
    inside = idx < steps.clamp(min=1).unsqueeze(0)
    n = steps.clamp(min=1).to(torch.float64)
    mean = (c * inside).sum(0) / n
    last = c[(steps.clamp(min=1) - 1), torch.arange(B, device=dev)]
    pair = (idx[1:] < steps.clamp(min=1).unsqueeze(0))
    auc = (0.5 * (c[1:] + c[:-1]) * pair).sum(0) if T > 1 else torch.zeros(B, dtype=c.dtype, device=dev)
    return {"tstt_curve": c, "reward": total, "tstt_last": last, "tstt_mean": mean, "tstt_auc": auc, "auc": auc,
            "steps": steps}
