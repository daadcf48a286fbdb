"""world_size-2 gloo tests (CPU) of the multi-rank host logic: scenario sharding, transition
all-gather order, gradient averaging."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    import importlib
    d = importlib.import_module("sac-gat-her_transportationrl_b200.dist")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = 10
        start, stop = d.scenario_shard(B, rank, world)
        # every rank fabricates "results" that depend only on the global scenario id
        ids = torch.arange(start, stop, dtype=torch.int64)
        local = {"scenario": ids, "tstt": ids.double() * 1.5, "obs": ids.float().view(-1, 1).repeat(1, 3)}
        gathered = d.all_gather_transitions(local)
        assert torch.equal(gathered["scenario"], torch.arange(B))
        assert torch.equal(gathered["tstt"], torch.arange(B).double() * 1.5)
        assert gathered["obs"].shape == (B, 3)
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Linear(8, 2))
        x = torch.full((3, 4), float(rank + 1))
        model(x).sum().backward()
        before = [p.grad.clone() for p in model.parameters()]
        d.allreduce_mean_gradients(model.parameters(), bucket_bytes=64)
        after = [p.grad.clone() for p in model.parameters()]
        torch.save({"before": before, "after": after}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_partition_properties():
    import importlib
    sys.path.insert(0, REPO)
    d = importlib.import_module("sac-gat-her_transportationrl_b200.dist")
    for B in (0, 1, 7, 4096, 1025):
        for world in (1, 2, 3, 8):
            blocks = [d.scenario_shard(B, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        d.scenario_shard(4, 2, 2)


def test_gloo_world2_gather_and_gradient_mean(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{r}.pt") for r in (0, 1))
    for b0, b1, a0, a1 in zip(r0["before"], r1["before"], r0["after"], r1["after"]):
        torch.testing.assert_close(a0, (b0 + b1) / 2)
        torch.testing.assert_close(a0, a1)
