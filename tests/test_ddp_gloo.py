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
    res = {"local": local, "avg": opt.segments[0].fb.grad.clone(), "w": model.a.weight.detach().clone()}
    # a trainer that clears gradients with model.zero_grad(set_to_none=True): autograd then allocates fresh p.grad tensors that
    # no longer alias the flat buffer; the exchange must pick them up (and re-establish the aliasing), not reduce stale zeros
    model.zero_grad(set_to_none=True)
    opt.segments[0].fb.grad.zero_()
    model(x).backward()
    assert model.a.weight.grad.data_ptr() != opt.segments[0].fb.view(0, opt.segments[0].fb.grad).data_ptr()
    ddp.finish()
    res["avg_foreign"] = opt.segments[0].fb.grad.clone()
    res["realiased"] = all(p.grad is None or p.grad.data_ptr() == opt.segments[0].fb.view(i, opt.segments[0].fb.grad).data_ptr()
                           for i, p in enumerate(opt.segments[0].fb.params))
    torch.save(res, os.path.join(out, "r%d.pt" % rank))
    dist.destroy_process_group()


class _ToyEnc(torch.nn.Module):
    """The part of BEIT3's surface FlatDDP uses: a flat buffer with a `text_embed` table and the compact-gradient mailbox."""

    def __init__(self):
        super().__init__()
        self.table = torch.nn.Embedding(50, 8)
        self.lin = torch.nn.Linear(8, 8)
        self._flat = None
        self._ddp = None
        self.sparse_text_grad = {"defer": False, "ids": None, "rows": None}

    def flat(self):
        from simvg_b200.flat import FlatBuffer
        if self._flat is None:
            self._flat = FlatBuffer([("w", self.lin.weight), ("text_embed", self.table.weight), ("b", self.lin.bias)])
        return self._flat.ensure()


class _ToyDet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.vis_enc = _ToyEnc()
        self.head = torch.nn.Linear(8, 1)


class _Opt2:
    def __init__(self, model):
        from simvg_b200.flat import FlatBuffer
        from simvg_b200.optim import _Segment
        self.segments = [_Segment("vis_enc", model.vis_enc.flat(), 0),
                         _Segment("rest", FlatBuffer(list(model.head.named_parameters())), 2)]
        for s_ in self.segments:
            s_.fb.attach_grads()


def _worker_sparse(rank, world, port, out, deferred):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simvg_b200.optim import FlatDDP
    torch.manual_seed(0)
    model = _ToyDet()
    opt = _Opt2(model)
    ddp = FlatDDP(model, opt, deferred=bool(deferred))
    assert model.vis_enc.sparse_text_grad["defer"]
    g = torch.Generator().manual_seed(100 + rank)
    ids = torch.randint(0, 50, (6,), generator=g)
    ids[0] = ids[1]                                           # a repeated id inside one rank
    rows = torch.randn(6, 8, generator=g)
    fb = model.vis_enc.flat()
    dense_local = torch.zeros(50, 8).index_add_(0, ids, rows)  # what a dense backward would have produced on this rank
    fb.grad_of(0).copy_(torch.randn(8, 8, generator=g))
    fb.grad_of(2).copy_(torch.randn(8, generator=g))
    opt.segments[1].fb.grad.copy_(torch.randn(opt.segments[1].fb.numel, generator=g))
    local = [s_.fb.grad.clone() for s_ in opt.segments]
    local[0][fb.offsets[1]:fb.offsets[1] + 400] = dense_local.flatten()
    # the compact form the encoder backward leaves behind; the dense table gradient stays zero
    model.vis_enc.sparse_text_grad.update(ids=ids, rows=rows)
    if not deferred:
        ddp.on_encoder_backward_start()
        ddp.on_encoder_range_done(0, fb.numel)
    if deferred == 2:     # the chunked graph runtime's sequence: heads, then encoder ranges top-down, then wait()
        cut = fb.offsets[1] + 123                     # a boundary inside the sparsely exchanged table
        ddp.reduce_async(heads=True)
        ddp.reduce_async(cut, fb.numel)
        ddp.reduce_async(0, cut)
        ddp.wait()
    else:
        ddp.finish()
    torch.save({"local": local, "avg": [s_.fb.grad.clone() for s_ in opt.segments]}, os.path.join(out, "s%d_%d.pt" % (int(deferred), rank)))
    dist.destroy_process_group()


def test_sparse_text_embedding_exchange_world2_gloo(tmp_path):
    """The text-embedding gradient travels as (ids, rows) per rank (all-gather + local scatter-add) while every other range is
    all-reduced; all FlatDDP modes (overlap hooks / deferred exchange after the backward graph / asynchronous ranges between the chunked
    backward graphs) must leave each rank with
    exactly the mean of the dense per-rank gradients."""
    for deferred in (0, 1, 2):
        port = _free_port()
        mp.spawn(_worker_sparse, args=(2, port, str(tmp_path), deferred), nprocs=2, join=True)
        r0 = torch.load(tmp_path / ("s%d_0.pt" % int(deferred)))
        r1 = torch.load(tmp_path / ("s%d_1.pt" % int(deferred)))
        for k in range(2):
            want = (r0["local"][k] + r1["local"][k]) / 2
            assert torch.allclose(r0["avg"][k], want, atol=1e-6), (deferred, k)
            assert torch.equal(r0["avg"][k], r1["avg"][k])


def test_flat_ddp_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["w"], r1["w"])                                   # broadcast
    want = (r0["local"] + r1["local"]) / 2
    assert torch.allclose(r0["avg"], want, atol=1e-7) and torch.equal(r0["avg"], r1["avg"])
    assert float(want.abs().sum()) > 0
    assert torch.allclose(r0["avg_foreign"], want, atol=1e-7) and torch.equal(r0["avg_foreign"], r1["avg_foreign"])
    assert r0["realiased"] and r1["realiased"]


def _crit_worker(rank, world, port, out):
    """REC criterion under torch.distributed: num_boxes = clamp(all_reduce(B) / world, 1) stays a device tensor (no .item())."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simvg_b200.core.criterion.criterion import BatchedTargets, HungarianMatcher, SetCriterion
    g = torch.Generator().manual_seed(7 + rank)
    B = 3 + rank                                    # different local batch sizes: the all-reduce matters
    logits, boxes = torch.randn(B, 1, 2, generator=g), torch.rand(B, 1, 4, generator=g) * 0.4 + 0.2
    gt = torch.rand(B, 4, generator=g) * 0.3 + 0.3
    zeros = torch.zeros(B, 1, dtype=torch.int64)
    plain = [{"labels": zeros[i], "boxes": gt[i:i + 1]} for i in range(B)]
    batched = BatchedTargets(plain, boxes=gt, labels=zeros[:, 0])
    crit = SetCriterion(1, HungarianMatcher(1, 5.0, 2.0, "ce_cost"), {"loss_class": 1, "loss_bbox": 5.0, "loss_giou": 2.0},
                        loss_class_type="ce_loss", eos_coef=0.1)
    outp = {"pred_logits": logits, "pred_boxes": boxes}
    fast, generic = crit(outp, batched), crit(outp, plain)
    torch.save({"fast": {k: float(v) for k, v in fast.items()}, "generic": {k: float(v) for k, v in generic.items()},
                "l1_sum": float(torch.nn.functional.l1_loss(boxes[:, 0], gt, reduction="none").sum()), "B": B},
               os.path.join(out, "c%d.pt" % rank))
    dist.destroy_process_group()


def test_rec_criterion_num_boxes_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_crit_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        d = torch.load(tmp_path / ("c%d.pt" % r))
        for k, v in d["generic"].items():           # generic path = the reference's arithmetic (all_reduce + .item())
            assert abs(d["fast"][k] - v) <= 1e-6 * max(abs(v), 1.0), (r, k, d["fast"][k], v)
        assert abs(d["fast"]["loss_bbox"] - d["l1_sum"] / 3.5) < 1e-5   # (3 + 4) boxes / 2 ranks
