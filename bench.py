#!/usr/bin/env python
"""bench.py - converged BPR assignments/sec of the batched env-step hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): Sioux Falls (24 nodes / 76 links / 528 OD), Frank-Wolfe with
the reference's fixed step 2/(it+2), K_iters = 20 iterations per assignment, 4096 random damage
scenarios (reference damage sampler: 22 of 76 links, connectivity retry) PER GPU, cold start.
One "step" = one batched compute_flow_assignment over the whole batch (= 4096 assignments).
Multi-GPU: scenarios are independent, every rank solves its own 4096 (weak scaling, no
data-path collective); value = all ranks' assignments / max-over-ranks device time.

`--impl reference` times the REFERENCE ITSELF: its unmodified `src.env.repair_env.RepairEnv`
(`sp_backend="scipy", use_torch=False` - numpy BPR + scipy.csgraph.dijkstra + Python predecessor
walk), imported from the verbatim copy in oracle/_ref/ (made by oracle/fetch_ref.py; git-ignored, it
travels to the GPU box with the snapshot), one env per process on all host cores - the structure of
the reference's rollout workers / run_multiseed - on a bounded sample of the same workload
(`cpu_baseline.kind = "reference"`).  Only if that copy is missing does it fall back to the numpy/scipy
restatement oracle/ref_port.py (`kind = "port"`).  `cpu_baseline_ref_gpu` is the reference's own GPU
path (`sp_backend="torch", use_torch=True, device="cuda"`: dense Floyd-Warshall in torch,
src/env/repair_env.py:520-573) on the same B200.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

METRIC = "converged BPR assignments/sec (batched env steps)"
UNIT = "assignments/s"
SCENARIOS_PER_GPU = 4096
ASSIGN_ITERS = 20
METHOD = "fw"
L2_FLUSH_BYTES = 256 << 20


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": "Sioux Falls TNTP (24 nodes, 76 links, 528 OD), Frank-Wolfe fixed step, "
                    f"K_iters={ASSIGN_ITERS}, cold start, {SCENARIOS_PER_GPU} random damage scenarios per GPU "
                    "(damaged_ratio 0.3, reference sampler)",
        "scenarios_per_gpu": SCENARIOS_PER_GPU,
        "global_scenarios": SCENARIOS_PER_GPU * n_gpus,
        "assignment_iters": ASSIGN_ITERS,
        "assignment_method": METHOD,
        "precision": "fp32 link state / fp64 distances (reference dtypes)",
        "parallelism": f"scenario-sharded x{n_gpus}, no data-path collective",
        "l2": f"flushed between timed steps by writing a {L2_FLUSH_BYTES >> 20} MiB buffer",
    }


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons of one GPU, sampled every few ms by a thread through NVML while
    the timed regions run (the timed region is tens of ms long - too short for `nvidia-smi -lms`,
    which is only the fallback when pynvml cannot be loaded)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, torch_index: int, period_s: float = 0.002):
        self.index, self.period = torch_index, period_s
        self.sm, self.mask, self.max_mhz = [], 0, None
        self.stop_flag = threading.Event()
        self.thread = None
        self.handle = None
        self.err = None

    def start(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception as exc:  # noqa: BLE001
            self.err = f"pynvml unavailable: {exc}"
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception as exc:  # noqa: BLE001
                self.err = str(exc)
                return
            time.sleep(self.period)

    def stop(self) -> dict:
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [self.err or "no samples"], "samples": 0}
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm), "how": "NVML, 2 ms period, during both timed regions"}


# ---------------------------------------------------------------------------------------------
# CPU legs (oracle = checker/baseline only; never on the product path)
# ---------------------------------------------------------------------------------------------
def _port_worker(args):
    seed0, n_assign, k_iters, method = args
    import b200env
    from oracle import ref_port as rp
    g = b200env.load_sioux_falls()
    net = rp.PortNetwork.from_graph_data(g)
    rng = np.random.default_rng(seed0)
    masks = []
    for _ in range(n_assign):
        m = np.zeros(net.num_edges, np.float32)
        m[rp.sample_damage(net, rng, 0.3)] = 1.0
        masks.append(m)
    t0 = time.perf_counter()
    acc = 0.0
    for m in masks:
        cap = net.cap0.copy()
        cap[m > 0] = 1e-3
        _, tstt, _ = rp.assign(net, cap, m, None, k_iters, method)
        acc += tstt
    return time.perf_counter() - t0, acc


def _draw_masks(seed0: int, n: int, row=None, col=None, num_nodes: int = 24):
    """n damage masks of the bench workload, drawn on the host with the reference's rule (reset,
    src/env/repair_env.py:167-190: 22 of 76 links, retry until the undamaged links form a strongly
    connected subgraph).  Self-contained (numpy + scipy) so that the reference arm imports nothing of
    the product package."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    E = int(row.shape[0])
    k = max(1, int(E * 0.3))
    rng = np.random.default_rng(seed0)
    masks = np.zeros((n, E), np.uint8)
    for i in range(n):
        cand = None
        for _ in range(50):
            cand = rng.choice(E, size=k, replace=False)
            act = np.ones(E, bool)
            act[cand] = False
            nodes = np.unique(np.concatenate([row[act], col[act]]))
            sub = csr_matrix((np.ones(int(act.sum())), (row[act], col[act])), shape=(num_nodes, num_nodes))
            _, labels = connected_components(sub, directed=True, connection="strong")
            if np.unique(labels[nodes]).size == 1:
                break
        masks[i, cand] = 1
    return masks


def _reference_worker(args):
    """One process = one reference RepairEnv (its CPU path), n cold-start compute_flow_assignment()."""
    seed0, n_assign, k_iters, method = args
    from oracle import reference_env as R
    env = R.make_env(assignment_method=method, assignment_iters=k_iters, damaged_ratio=0.3, seed=seed0)
    masks = _draw_masks(seed0, n_assign, env.edge_index[0], env.edge_index[1], env.num_nodes)
    wall = R.time_assignments(env, masks)
    return wall, float(env.tstt)


def _reference_available() -> bool:
    from oracle import reference_env as R
    return R.available()


def cpu_reference_rate(n_per_proc: int, procs: int) -> dict:
    """Aggregate assignments/s of the unmodified reference env, one env per process (spawned, like
    the reference's rollout workers src/train.py:138-214 / scripts/run_multiseed.py:37-38)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_reference_worker, [(1, 1, 2, METHOD)] * procs)  # import + warm-up
        t0 = time.perf_counter()
        pool.map(_reference_worker, [(5000 + i, n_per_proc, ASSIGN_ITERS, METHOD) for i in range(procs)])
        wall = time.perf_counter() - t0
    n = n_per_proc * procs
    return {"value": n / wall, "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": f"{n} cold-start Sioux Falls {METHOD} K={ASSIGN_ITERS} assignments, {procs} processes x "
                      f"{n_per_proc}, the reference's own RepairEnv.compute_flow_assignment (oracle/_ref copy of "
                      f"src/env/repair_env.py, sp_backend=scipy, use_torch=False), wall {wall:.1f}s"}


def reference_gpu_rate(n_assign: int = 12) -> dict:
    """The reference's own GPU path on this box's GPU: torch BPR + dense torch Floyd-Warshall
    (src/env/repair_env.py:520-573, :679-705; the setup of the reference's test_env_gpu.py:22-31)."""
    from oracle import reference_env as R
    env = R.make_env(assignment_method=METHOD, assignment_iters=ASSIGN_ITERS, damaged_ratio=0.3, seed=5000,
                     sp_backend="torch", use_torch=True, device="cuda", force_gpu_sp=True)
    masks = _draw_masks(5000, n_assign + 1, env.edge_index[0], env.edge_index[1], env.num_nodes)
    R.time_assignments(env, masks[:1])
    wall = R.time_assignments(env, masks[1:])
    return {"value": n_assign / wall, "unit": UNIT, "kind": "reference-gpu", "gpus": 1,
            "sample": f"{n_assign} cold-start Sioux Falls {METHOD} K={ASSIGN_ITERS} assignments, one reference "
                      f"RepairEnv(sp_backend='torch', use_torch=True, device='cuda', force_gpu_sp=True), wall {wall:.1f}s"}


def cpu_port_rate(n_per_proc: int, procs: int) -> dict:
    """Aggregate assignments/s of the numpy/scipy restatement, one env per process."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_port_worker, [(1, 1, 2, METHOD)] * procs)  # import + warm-up
        t0 = time.perf_counter()
        pool.map(_port_worker, [(5000 + i, n_per_proc, ASSIGN_ITERS, METHOD) for i in range(procs)])
        wall = time.perf_counter() - t0
    n = n_per_proc * procs
    return {"value": n / wall, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{n} cold-start Sioux Falls {METHOD} K={ASSIGN_ITERS} assignments, {procs} processes x "
                      f"{n_per_proc} (oracle/ref_port.py: numpy fp32 + scipy dijkstra + Python pred walk), "
                      f"wall {wall:.1f}s"}


def c_oracle_rate(seconds: float = 3.0) -> dict:
    import b200env
    from oracle.c_oracle import COracle
    g = b200env.load_sioux_falls()
    orc = COracle.from_graph_data(g)
    rng = np.random.default_rng(7)
    masks = np.zeros((64, orc.E), np.uint8)
    for b in range(64):
        masks[b, rng.choice(orc.E, 22, replace=False)] = 1
    orc.assign(masks[0], None, ASSIGN_ITERS, METHOD)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        orc.assign(masks[n % 64], None, ASSIGN_ITERS, METHOD)
        n += 1
    return {"value": n / (time.perf_counter() - t0), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} assignments, oracle/assign_oracle.c (gcc -O2), single thread"}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    per_step = 2  # assignments per process per "step": bounded sample
    total_steps = args.steps + args.warmup
    rate = cpu_reference_rate if _reference_available() else cpu_port_rate
    res = rate(per_step * max(total_steps, 1), procs)
    n_per_step = per_step * procs
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_per_step / res["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus), "cpu_baseline": res,
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def measured_peak() -> tuple:
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes():
    path = os.path.join(REPO, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get("dram_bytes_per_launch")
    return None


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import b200env

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the assignment path has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    b200env.build_native()

    g = b200env.load_sioux_falls()
    B = SCENARIOS_PER_GPU
    env = b200env.BatchedRepairEnv(g, B, assignment_method=METHOD, assignment_iters=ASSIGN_ITERS, device=dev,
                                   precision="fp32", seeds=[1000 + rank * B + b for b in range(B)])
    masks = env.sample_damage_masks(0.3)
    pinned = torch.from_numpy(masks).pin_memory()
    env.is_damaged.copy_(pinned)
    eng = env.engine
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    N, E, O = eng.N, eng.E, eng.O
    bytes_per_assignment = ASSIGN_ITERS * (33 * E + 8 * O * N)  # SURVEY.md §8(d), fp32 state

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step():
        env.flow.zero_()
        env.compute_flow_assignment()

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: device-resident inputs, CUDA events on the launching stream --------
    launches0 = eng.launches
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # L2 flush, outside the timed events
        starts[i].record()
        env.flow.zero_()
        mids[i].record()
        env.compute_flow_assignment()
        ends[i].record()
    barrier()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    kern_ms = [m.elapsed_time(e) for m, e in zip(mids, ends)]
    launches = eng.launches - launches0
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) / 1e3

    # ---- timed region 2: end to end through the host-buffer entry point -----------------------
    out_host = np.empty(B, dtype=np.float64)
    for _ in range(2):
        env.assign_host(pinned.numpy(), out_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        env.assign_host(pinned.numpy(), out_host)  # H2D masks + assignment + D2H TSTT, synchronous
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    gap_med = float(env.gap.median().item())
    tstt_chk = float(env.tstt.sum().item())

    other = None
    if not args.no_extra:
        other = measure_sf_extras(b200env, torch, dist, dev, rank, world, g, masks, measured_peak()[0])
        other.update(measure_other_configs(b200env, torch, dist, dev, rank, world, measured_peak()[0]))
    if rank == 0:
        peak, peak_src = measured_peak()
        kern_s = statistics.mean(kern_ms) / 1e3
        achieved = bytes_per_assignment * B / kern_s / 1e9
        value = B * world * args.steps / total_s
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": B * world * args.steps / float(e2e_s.item()), "unit": UNIT,
                    "h2d_bytes_per_step": int(B * E), "d2h_bytes_per_step": int(B * 8)},
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_bytes(), "peak_source": peak_src,
                         "kernel": ("assign_static_grp_kernel<TopoSiouxFalls>" if eng.last_config.get("cch") == 2
                                    else eng.kernel_name + "<float>"), "kernel_ms": 1e3 * kern_s,
                         "algorithmic_bytes_per_launch": bytes_per_assignment * B,
                         "note": "algorithmic bytes = K*(33E+8ON) per scenario (SURVEY.md 8d); the working set is "
                                 "register/shared-memory resident (traffic = ncu DRAM bytes per launch), the "
                                 "binding limit is the shared-memory wavefront pipe (DESIGN.md section 4)"},
            "quality": {"median_relative_gap_at_K": gap_med, "tstt_checksum": tstt_chk},
        }
        line["other_configs"] = other
        if world == 1 and not args.no_cpu_baseline:
            procs = os.cpu_count() or 1
            # ~0.1-0.2 s per assignment per core -> 100 per process is a 10-30 s bounded sample
            if _reference_available():
                line["cpu_baseline"] = cpu_reference_rate(100, procs)
                try:
                    line["cpu_baseline_ref_gpu"] = reference_gpu_rate()
                except Exception as exc:  # noqa: BLE001 - informational leg
                    line["cpu_baseline_ref_gpu"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
            else:
                line["cpu_baseline"] = cpu_port_rate(100, procs)
            line["cpu_baseline_c_oracle"] = c_oracle_rate()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _max_over_ranks(torch, dist, dev, world: int, value: float) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measure_sf_extras(b200env, torch, dist, dev, rank: int, world: int, g, masks, peak: float) -> dict:
    """Sioux Falls figures beyond the headline launch (all CUDA-event timed, max over ranks):
    the WHOLE env step, the north star's target configuration (4096 scenarios in TOTAL, sharded over
    the ranks = strong scaling), the line-search extension, and the end-to-end call with the link
    flows copied back to the host as well."""
    out = {}
    B, E = masks.shape

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, reps: int):
        ms = []
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sync()
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize(dev)
            ms.append(_max_over_ranks(torch, dist, dev, world, s.elapsed_time(e)))
        return ms

    # ---- one full BatchedRepairEnv.step: apply actions + warm-start assignment + reward/done + state features
    env = b200env.BatchedRepairEnv(g, B, assignment_method=METHOD, assignment_iters=ASSIGN_ITERS, device=dev,
                                   precision="fp32")
    env.reset(damage_masks=masks)
    order = np.argsort(-masks.astype(np.int8), axis=1, kind="stable")[:, :22]      # the damaged links, ascending
    acts = torch.from_numpy(np.ascontiguousarray(order.T.astype(np.int32))).to(dev)  # [22, B]: one valid repair per step
    for i in range(3):
        env.step(acts[i], validate_actions=False)
    l0 = env.engine.launches
    step_ms = []
    for i in range(3, 19):
        step_ms += timed(lambda i=i: env.step(acts[i], validate_actions=False), 1)
    ms = statistics.mean(step_ms)
    out["sf_step_4096"] = {"env_steps_per_s": B * world / ms * 1e3, "ms_per_step": ms, "steps_timed": len(step_ms),
                           "scenarios_per_gpu": B, "scaling": "weak", "launches_per_step": (env.engine.launches - l0) / 16,
                           "what": "BatchedRepairEnv.step: apply_actions + fused warm-start fw K=20 assignment + "
                                   "reward/done + state features (Brandes betweenness) + returned copies"}
    del env

    eng = b200env.AssignmentEngine(b200env.Network(g), dev, "fp32")
    bytes_per = ASSIGN_ITERS * (33 * eng.E + 8 * eng.O * eng.N)

    # ---- the target configuration: 4096 scenarios in total, sharded over the ranks
    lo, hi = B * rank // world, B * (rank + 1) // world
    dmg = torch.from_numpy(masks[lo:hi]).to(dev)
    flow = torch.zeros((hi - lo, E), dtype=torch.float32, device=dev)
    eng.assign(dmg, flow, ASSIGN_ITERS, METHOD)

    def strong():
        flow.zero_()
        eng.assign(dmg, flow, ASSIGN_ITERS, METHOD)
    ms = min(timed(strong, 5))
    out["sf_fw_4096_total_strong"] = {"assignments_per_s": B / ms * 1e3, "ms_per_launch": ms, "scenarios_total": B,
                                      "scenarios_per_gpu": hi - lo, "scaling": "strong",
                                      "roofline_frac_per_gpu": bytes_per * (hi - lo) / (ms / 1e3) / 1e9 / peak}

    # ---- fw with the bisection line search (north-star extension): rate and the gap it buys
    dmg_all = torch.from_numpy(masks).to(dev)
    flow_all = torch.zeros((B, E), dtype=torch.float32, device=dev)
    res = {}

    def fw_ls():
        flow_all.zero_()
        res["r"] = eng.assign(dmg_all, flow_all, ASSIGN_ITERS, "fw_ls")
    fw_ls()
    ms = min(timed(fw_ls, 3))
    out["sf_fw_ls_4096"] = {"assignments_per_s": B * world / ms * 1e3, "ms_per_launch": ms, "scaling": "weak",
                            "line_search_iters": eng.ls_iters,
                            "median_relative_gap_at_K": float(res["r"]["gap"].median().item())}
    del eng

    # ---- end to end with the link flows returned as well (masks in; TSTT + flows out)
    env = b200env.BatchedRepairEnv(g, B, assignment_method=METHOD, assignment_iters=ASSIGN_ITERS, device=dev,
                                   precision="fp32")
    pinned = torch.from_numpy(masks).pin_memory()
    out_t, out_f = np.empty(B, np.float64), np.empty((B, E), np.float32)
    for _ in range(2):
        env.assign_host(pinned.numpy(), out_t, out_f)
    sync()
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        env.assign_host(pinned.numpy(), out_t, out_f)
    wall = _max_over_ranks(torch, dist, dev, world, time.perf_counter() - t0)
    out["e2e_flows"] = {"value": B * world * reps / wall, "unit": UNIT, "h2d_bytes_per_step": int(B * E),
                        "d2h_bytes_per_step": int(B * 8 + B * E * 4), "steps": reps}
    del env
    return out


def measure_other_configs(b200env, torch, dist, dev, rank: int, world: int, peak: float) -> dict:
    """BASELINE.json configs[2] and [3] - the synthetic Anaheim-size and Chicago-Sketch-size networks -
    reported next to the headline line (informational: they are parity-test cases, not the metric).
    The named scenario counts are TOTALS, sharded over the ranks (strong scaling, no collective);
    time = max over ranks of the CUDA-event time of one fused launch (cold start, K = 20)."""
    out = {}
    cases = (("anaheim_size_fw_1024", lambda: b200env.synthetic.anaheim_size(0), 1024, "fw"),
             ("chicago_sketch_size_msa_256", lambda: b200env.synthetic.chicago_sketch_size(0), 256, "msa"))
    for name, make, total, method in cases:
        eng = b200env.AssignmentEngine(b200env.Network(make()), dev, "fp32")
        lo, hi = total * rank // world, total * (rank + 1) // world
        rng = np.random.default_rng(1234)
        masks = np.zeros((total, eng.E), np.uint8)
        k = max(1, int(eng.E * 0.3))
        for b in range(total):
            masks[b, rng.choice(eng.E, k, replace=False)] = 1
        dmg = torch.from_numpy(masks[lo:hi]).to(dev)
        flow = torch.zeros((hi - lo, eng.E), dtype=eng.real, device=dev)
        eng.assign(dmg, flow, ASSIGN_ITERS, method)
        reps, ms = 3, []
        for _ in range(reps):
            flow.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            s.record()
            res = eng.assign(dmg, flow, ASSIGN_ITERS, method)
            e.record()
            torch.cuda.synchronize(dev)
            t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms.append(float(t.item()))
        best = min(ms)
        bytes_per = ASSIGN_ITERS * (33 * eng.E + 8 * eng.O * eng.N)
        out[name] = {"assignments_per_s": total / best * 1e3, "ms_per_launch": best, "scenarios_total": total,
                     "scenarios_per_gpu": hi - lo, "method": method, "assignment_iters": ASSIGN_ITERS,
                     "nodes": eng.N, "links": eng.E, "origins": eng.O, "kernel": eng.kernel_name, "scaling": "strong",
                     "roofline_frac_per_gpu": bytes_per * (hi - lo) / (best / 1e3) / 1e9 / peak,
                     "median_relative_gap_at_K": float(res["gap"].median().item())}
        del eng
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the informational Anaheim / Chicago-Sketch-size figures")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
