"""Drop-in tests: the REFERENCE'S OWN callers run on top of ``b200env.RepairEnv``.

The unmodified reference travels to the GPU box as ``oracle/_ref`` (oracle/fetch_ref.py).  These tests
import its ``src.baselines`` (``run_episode``, ``get_baseline_policies``, ``select_greedy_one_step`` -
src/baselines/__init__.py:35-113) and ``src.eval.evaluate`` (src/eval.py:28-185), execute them once on the
reference's own CPU ``RepairEnv`` and once on the B200 drop-in, and compare what the callers see: TSTT
curves, rewards, chosen actions, state features.  Tolerance: 1e-4 relative (north star, fp32 state).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@pytest.fixture(scope="module")
def ref():
    from oracle import fetch_ref
    from oracle import reference_env as R
    assert R.available(), ("oracle/_ref (verbatim copy of the reference) did not travel to this box - run "
                           "`python oracle/fetch_ref.py` in the build container before gpurun")
    assert fetch_ref.verify(), "oracle/_ref differs from the recorded sha256 manifest: the reference must be unmodified"
    return R


def _both_envs(ref, b200env, **kw):
    graph = ref.sioux_falls_graph()            # the reference's own GraphData: the drop-in must accept it
    kw.setdefault("assignment_iters", 20)
    theirs = ref.make_env(graph, **kw)
    ours = b200env.RepairEnv(graph, device="cuda:0", **kw)
    return theirs, ours


def test_reference_run_episode_with_greedy_and_betweenness_policies(ref, b200env, native_lib):
    """`run_episode` with the reference's `greedy_one_step` (pokes is_damaged / capacities / flow / tstt /
    unassigned_demand / od_paths and calls compute_flow_assignment, :43-67) and `max_betweenness`
    (reads betweenness_vec / edge_index, :105-106) policies, on both envs, one seed."""
    _, _, baselines = ref.modules()
    for policy_name, ratio in (("greedy", 0.1), ("max_betweenness", 0.3), ("max_vc", 0.3), ("max_flow", 0.3)):
        theirs, ours = _both_envs(ref, b200env, seed=43, damaged_ratio=ratio, assignment_method="msa")
        assert np.array_equal(theirs.is_damaged, ours.is_damaged)                 # same RNG stream + retry rule
        np.testing.assert_allclose(ours.betweenness_vec, theirs.betweenness_vec, rtol=1e-5, atol=1e-7)
        res = {}
        for name, env in (("ref", theirs), ("b200", ours)):
            # run_episode() starts with env.reset(): default damaged_ratio 0.3, the env's own rng continues;
            # greedy is cut at 6 steps (22 + 21 + .. + 17 candidate assignments on the reference's CPU path)
            policy = baselines.get_baseline_policies(env)[policy_name]
            res[name] = baselines.run_episode(env, policy, reward_scale=1.0, max_steps=6 if policy_name == "greedy" else 100)
        a, b = res["ref"], res["b200"]
        assert len(a["tstt_curve"]) == len(b["tstt_curve"]) > 0, policy_name
        np.testing.assert_allclose(b["tstt_curve"], a["tstt_curve"], rtol=RTOL, err_msg=policy_name)
        np.testing.assert_allclose(b["reward"], a["reward"], rtol=1e-3, atol=1e-4, err_msg=policy_name)
        np.testing.assert_allclose(b["tstt_auc"], a["tstt_auc"], rtol=RTOL)
        assert np.array_equal(theirs.is_damaged, ours.is_damaged), f"{policy_name}: the two episodes repaired different links"


def test_reference_state_and_step_agree_along_an_episode(ref, b200env, native_lib):
    """What src/train.py consumes: EnvState arrays, reward, done, info - step by step, cfw K=60 as in the
    shipped configs/sioux_falls.yaml (:5, :10)."""
    theirs, ours = _both_envs(ref, b200env, seed=7, damaged_ratio=0.3, assignment_method="cfw", assignment_iters=60,
                              reward_mode="rel_improve", reward_alpha=10.0, reward_gamma=0.01, reward_clip=50.0,
                              unassigned_penalty=1e4)
    rng = np.random.default_rng(0)
    done = False
    steps = 0
    while not done:
        cand = np.flatnonzero(theirs.is_damaged > 0)
        a = int(rng.choice(cand)) if steps != 4 else int(np.flatnonzero(theirs.is_damaged == 0)[0])  # one invalid action
        s1, r1, d1, i1 = theirs.step(a)
        s2, r2, d2, i2 = ours.step(a)
        assert d1 == d2 and abs(i1["tstt"] - i2["tstt"]) <= 5e-4 * abs(i1["tstt"])   # cfw: fp32 dot products feed back
        np.testing.assert_allclose(r2, r1, rtol=5e-3, atol=5e-3)
        np.testing.assert_allclose(s2.node_features, s1.node_features, rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(s2.edge_features, s1.edge_features, rtol=1e-3, atol=1e-4)
        assert np.array_equal(s2.action_mask, s1.action_mask) and np.array_equal(s2.goal_mask, s1.goal_mask)
        assert np.array_equal(s2.edge_index, s1.edge_index) and abs(s1.log_tstt - s2.log_tstt) < 1e-3
        done = d1
        steps += 1
    assert steps == 23   # 22 repairs + 1 invalid action


def test_reference_eval_evaluate_runs_on_the_drop_in(ref, b200env, native_lib, tmp_path, monkeypatch):
    """`src.eval.evaluate(cfg)` unmodified, once as shipped (reference env, CPU) and once with
    `src.eval.RepairEnv` replaced by the drop-in: the metrics files it writes must agree."""
    ev = ref.eval_module()
    results = {}
    for name in ("ref", "b200"):
        out_dir = tmp_path / name
        cfg = dict(data_dir=os.path.join(ref.root(), "data"), output_dir=str(out_dir), model_path=None, seed=42,
                   eval_seeds=[43], damaged_ratio=0.15, assignment_iters=20, assignment_method="fw",
                   sp_backend="scipy", use_torch_bpr=False, device="cpu", reward_mode="log_delta",
                   unassigned_penalty=1e4, max_steps=100, reward_scale=0.01, eval_baselines=True)
        if name == "b200":
            monkeypatch.setattr(ev, "RepairEnv", b200env.RepairEnv)
        np.random.seed(0)                      # the reference's `random` baseline uses the global numpy RNG
        ev.evaluate(cfg)
        with open(out_dir / "eval_metrics.json") as fh:
            results[name] = json.load(fh)["seed_43"]
    assert set(results["ref"]) == set(results["b200"]) == {"random", "max_vc", "max_flow", "max_betweenness", "greedy"}
    for pol, a in results["ref"].items():
        b = results["b200"][pol]
        assert len(a["tstt_curve"]) == len(b["tstt_curve"]), pol
        np.testing.assert_allclose(b["tstt_curve"], a["tstt_curve"], rtol=RTOL, err_msg=pol)
        np.testing.assert_allclose(b["tstt_auc"], a["tstt_auc"], rtol=RTOL, err_msg=pol)
        np.testing.assert_allclose(b["reward"], a["reward"], rtol=1e-3, atol=1e-6, err_msg=pol)
