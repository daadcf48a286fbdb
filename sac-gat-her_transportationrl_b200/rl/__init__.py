"""Learner side of the training loop around the batched env: discrete SAC on the GAT encoder,
device-resident prioritised replay, multi-GPU trainer (NCCL transition all-gather + gradient
all-reduce).  Reference: src/rl/sac.py, src/train.py."""
from .replay import DeviceReplay
from .sac import Actor, Critic, DiscreteSAC
from .trainer import BatchedTrainer, default_config

__all__ = ["Actor", "BatchedTrainer", "Critic", "DeviceReplay", "DiscreteSAC", "default_config"]
