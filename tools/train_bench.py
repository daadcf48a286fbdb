"""Throughput of the full SAC + GAT + PER loop on Sioux Falls (BASELINE.json config 5): env-steps/s,
assignments/s, learner updates/s, NCCL time share.  Not the bench.py metric; numbers quoted in
DESIGN.md.   python tools/train_bench.py [--envs 1024] [--steps 40]   (torchrun for several GPUs)"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import b200env


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=24)   # past the CUDA-graph capture of the learner update (4th update)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--method", default="cfw")
    ap.add_argument("--out", default="", help="also write the JSON line to this file (rank 0)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b200env.build_native()
    cfg = dict(num_envs=args.envs, assignment_iters=args.iters, assignment_method=args.method, batch_start=256,
               buffer_size=200_000)
    tr = b200env.rl.BatchedTrainer(b200env.load_sioux_falls(), cfg, device=f"cuda:{local}")
    tr.run(args.warmup)
    res = tr.run(args.steps)
    if int(os.environ.get("RANK", "0")) == 0:
        res["config"] = {"envs_per_gpu": args.envs, "assignment_iters": args.iters, "assignment_method": args.method,
                         "batch_size": tr.cfg["batch_size"], "update_every": tr.cfg["update_every"],
                         "hidden_dim": tr.cfg["hidden_dim"], "gat_layers": tr.cfg["gat_layers"]}
        print(json.dumps(res), flush=True)
        if args.out:
            with open(args.out, "w") as fh:
                json.dump(res, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
