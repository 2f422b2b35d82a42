"""World-size-2 data-parallel gradient exchange on CPU (gloo): FlatDDP must leave every rank with the mean of the per-rank
gradients in the flat buffer — the semantics of the reference's MMDistributedDataParallel (tools/train.py:102-103)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(6, 5)
        self.b = torch.nn.LayerNorm(5)
        self.unused = torch.nn.Parameter(torch.zeros(3))

    def forward(self, x):
        return self.b(self.a(x)).pow(2).sum()


class _Opt:  # just the part of FusedAdamAMSGrad that FlatDDP touches (the update kernel itself needs a GPU)
    def __init__(self, model):
        from simvg_b200.flat import FlatBuffer
        from simvg_b200.optim import _Segment
        self.segments = [_Segment("rest", FlatBuffer(list(model.named_parameters())), 1e-3)]
        self.segments[0].fb.attach_grads()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simvg_b200.optim import FlatDDP
    torch.manual_seed(0)
    model = _Toy()
    opt = _Opt(model)
    ddp = FlatDDP(model, opt)
    with torch.no_grad():                      # desynchronise, then broadcast from rank 0
        for p in model.parameters():
            p.add_(rank)
    ddp.broadcast_parameters()
    x = torch.randn(4, 6, generator=torch.Generator().manual_seed(100 + rank))
    model(x).backward()
    local = opt.segments[0].fb.grad.clone()
    ddp.finish()
    torch.save({"local": local, "avg": opt.segments[0].fb.grad.clone(), "w": model.a.weight.detach().clone()},
               os.path.join(out, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_flat_ddp_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["w"], r1["w"])                                   # broadcast
    want = (r0["local"] + r1["local"]) / 2
    assert torch.allclose(r0["avg"], want, atol=1e-7) and torch.equal(r0["avg"], r1["avg"])
    assert float(want.abs().sum()) > 0
