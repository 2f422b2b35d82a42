"""Process-group helpers (mirrors /root/reference/simvg/utils/distributed.py:9-27 and detrex.utils)."""
import os
from datetime import timedelta

import torch
from torch import distributed as dist


def init_dist(backend=None):
    """One process per GPU, launched by torch.distributed.run (tools/dist_train.sh:8-9)."""
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        backend = backend or "nccl"
    else:
        backend = backend or "gloo"
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, timeout=timedelta(minutes=3))


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main():
    return get_rank() == 0


def reduce_mean(tensor):
    if not is_dist_avail_and_initialized():
        return tensor
    tensor = tensor.clone()
    dist.all_reduce(tensor.div_(dist.get_world_size()), op=dist.ReduceOp.SUM)
    return tensor
