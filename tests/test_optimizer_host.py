"""CPU checks of FusedAdamAMSGrad's host side (no kernel launches): it is a torch.optim.Optimizer whose param_groups are the
reference's own three groups (/root/reference/tools/train.py:78-93), the reference's schedulers drive it, and its checkpoints
are interchangeable with torch.optim.Adam(amsgrad=True)'s — the optimiser the reference builds (core/optimizer.py:52-68)."""
import copy

import pytest
import torch


def _model(freeze_layer=-1):
    from simvg_b200.models import build_model
    from tools.synth import model_cfg
    cfg = model_cfg("base", 64, 32)
    cfg["vis_enc"]["freeze_layer"] = freeze_layer
    torch.manual_seed(0)
    return build_model(cfg)


def _reference_groups(model, lr, lr_vis, lr_lan):
    named = list(model.named_parameters())
    return [{"params": [p for n, p in named if "vis_enc" in n and p.requires_grad], "lr": lr_vis},
            {"params": [p for n, p in named if "lan_enc" in n and p.requires_grad], "lr": lr_lan},
            {"params": [p for n, p in named if "lan_enc" not in n and "vis_enc" not in n and p.requires_grad], "lr": lr}]


def test_is_torch_optimizer_with_reference_groups_and_schedulers():
    from simvg_b200.optim import FusedAdamAMSGrad
    model = _model()
    opt = FusedAdamAMSGrad(model, lr=5e-4, lr_vis_enc=5e-5, lr_lan_enc=5e-5, grad_norm_clip=0.15)
    assert isinstance(opt, torch.optim.Optimizer)
    ref = _reference_groups(model, 5e-4, 5e-5, 5e-5)
    assert len(opt.param_groups) == 3
    for g, r in zip(opt.param_groups, ref):
        assert g["lr"] == r["lr"] and len(g["params"]) == len(r["params"])
        assert all(a is b for a, b in zip(g["params"], r["params"]))     # same parameters, named_parameters() order
    assert len(opt.param_groups[1]["params"]) == 0                        # no lan_enc in any SimVG config
    # the reference's MultiStepLRWarmUp is a LambdaLR over the optimiser (core/scheduler.py); torch raises TypeError for
    # anything that is not an Optimizer
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda e: 0.1 if e >= 2 else 1.0)
    for _ in range(2):
        sched.step()
    assert opt.param_groups[0]["lr"] == pytest.approx(5e-6) and opt.param_groups[2]["lr"] == pytest.approx(5e-5)
    seg_lr = {s.name: opt._lr(s) for s in opt.segments}
    assert seg_lr["vis_enc"] == pytest.approx(5e-6) and seg_lr["rest"] == pytest.approx(5e-5)   # launches read the group's lr
    torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10)


def test_checkpoint_interchange_with_torch_adam():
    from simvg_b200.optim import FusedAdamAMSGrad
    model = _model()
    twin = copy.deepcopy(model)
    ref = torch.optim.Adam(_reference_groups(twin, 5e-4, 5e-5, 5e-5), betas=(0.9, 0.98), eps=1e-9, weight_decay=0, amsgrad=True)
    g = torch.Generator().manual_seed(1)
    for _ in range(2):
        for p in twin.parameters():
            p.grad = torch.randn(p.shape, generator=g) * 1e-2
        ref.step()
    sd = ref.state_dict()
    opt = FusedAdamAMSGrad(model, lr=1.0, lr_vis_enc=1.0, lr_lan_enc=1.0)
    opt.load_state_dict(copy.deepcopy(sd))
    assert opt.t == 2 and opt.param_groups[0]["lr"] == 5e-5 and opt.param_groups[2]["lr"] == 5e-4
    # every moment landed in the flat buffers under the right parameter (q/k/v are adjacent but in a different order there)
    names = dict((id(p), n) for n, p in model.named_parameters())
    twin_by_name = dict(twin.named_parameters())
    checked = 0
    for grp in opt.param_groups:
        for p in grp["params"]:
            seg, i = opt._where[id(p)]
            st = ref.state[twin_by_name[names[id(p)]]]
            assert torch.equal(seg.fb.view(i, seg.m), st["exp_avg"]), names[id(p)]
            assert torch.equal(seg.fb.view(i, seg.v), st["exp_avg_sq"])
            assert torch.equal(seg.fb.view(i, seg.vmax), st["max_exp_avg_sq"])
            checked += 1
    assert checked > 300
    # and back: our state_dict loads into a fresh torch Adam built the reference way
    out = opt.state_dict()
    assert [len(g_["params"]) for g_ in out["param_groups"]] == [len(g_["params"]) for g_ in sd["param_groups"]]
    ref2 = torch.optim.Adam(_reference_groups(copy.deepcopy(model), 1.0, 1.0, 1.0), betas=(0.9, 0.98), eps=1e-9, amsgrad=True)
    ref2.load_state_dict(out)
    k = sorted(sd["state"])[7]
    assert torch.equal(ref2.state_dict()["state"][k]["exp_avg"], sd["state"][k]["exp_avg"])
    assert float(ref2.state_dict()["state"][k]["step"]) == 2.0
    # a checkpoint with a different group layout is refused instead of being zipped silently
    bad = copy.deepcopy(sd)
    bad["param_groups"] = bad["param_groups"][:2]
    with pytest.raises(ValueError):
        opt.load_state_dict(bad)


def test_frozen_layers_are_left_out_of_groups_and_update_ranges():
    from simvg_b200.optim import FusedAdamAMSGrad
    model = _model(freeze_layer=2)
    opt = FusedAdamAMSGrad(model, lr=5e-4, lr_vis_enc=5e-5)
    frozen = [p for p in model.parameters() if not p.requires_grad]
    assert len(frozen) == 2 * 40                                          # two layers x (A, B) x 20 tensors
    in_groups = {id(p) for g in opt.param_groups for p in g["params"]}
    assert not any(id(p) in in_groups for p in frozen)
    seg = opt.segments[0]
    fb = seg.fb
    covered = torch.zeros(fb.numel, dtype=torch.bool)
    for lo, hi in seg.ranges():
        covered[lo:hi] = True
    for i, p in enumerate(fb.params):
        assert bool(covered[fb.offsets[i]:fb.offsets[i] + p.numel()].all()) == p.requires_grad, fb.names[i]
    assert len(seg.ranges()) == 2                                         # embeddings + final LN | layers 2..11


def test_ema_decay_schedule_and_views():
    from simvg_b200.optim import FusedAdamAMSGrad
    model = _model()
    opt = FusedAdamAMSGrad(model, lr=5e-4, ema_alpha=0.999)
    assert [opt.ema_decay(t) for t in (0, 1, 90)] == [0.1, 2.0 / 11.0, 91.0 / 100.0]     # min(alpha, (1+t)/(10+t)): models/utils.py:149
    assert opt.ema_decay(100000) == 0.999
    p = model.head.query_embed.weight
    assert torch.equal(opt.ema_view(p), p.detach()) and opt.ema_view(p).data_ptr() != p.data_ptr()


def test_exponential_moving_average_has_the_reference_interface():
    """simvg/models/utils.py:130-180: update_params() (decay = min(alpha, (1 + step) / (10 + step))), apply_shadow() / restore()
    around evaluation, `shadow` for the checkpoint — standalone (any optimiser) and backed by the fused optimiser's flat EMA
    stream (swap of two buffers instead of state-dict clones)."""
    from simvg_b200.models import ExponentialMovingAverage
    from simvg_b200.optim import FusedAdamAMSGrad
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 2))
    ema = ExponentialMovingAverage(model, 0.9)
    want = {k: v.detach().clone() for k, v in model.named_parameters()}
    for step in range(4):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn_like(p) * 0.1)
        decay = min(0.9, (step + 1) / (step + 10))
        for k, p in model.named_parameters():
            want[k] = decay * want[k] + (1 - decay) * p.detach()
        ema.update_params()
    assert ema.step == 4
    live = {k: v.detach().clone() for k, v in model.named_parameters()}
    for k, v in ema.shadow.items():
        assert torch.allclose(v, want[k], atol=1e-6)
    ema.apply_shadow()
    for k, p in model.named_parameters():
        assert torch.allclose(p, want[k], atol=1e-6)
    with pytest.raises(RuntimeError):
        ema.apply_shadow()
    ema.restore()
    for k, p in model.named_parameters():
        assert torch.equal(p, live[k])
    # backed by the fused optimiser: the average lives in its flat buffers; apply / restore swap in place
    opt = FusedAdamAMSGrad(model, lr=1e-3)
    ema2 = ExponentialMovingAverage(model, 0.999, optimizer=opt)
    assert opt.ema_alpha == 0.999
    p0 = next(model.parameters())
    with torch.no_grad():
        opt.ema_view(p0).add_(1.0)
    before = p0.detach().clone()
    ema2.apply_shadow()
    assert torch.allclose(p0, before + 1.0)
    assert torch.allclose(opt.ema_view(p0), before)
    ema2.restore()
    assert torch.equal(p0, before)
    ema2.load_shadow({k: torch.zeros_like(v) for k, v in model.named_parameters()})
    assert float(opt.ema_view(p0).abs().max()) == 0.0


def test_exponential_moving_average_matches_the_reference_golden(golden_dir):
    """tests/golden/ema.pt was produced by the reference's own ExponentialMovingAverage (simvg/models/utils.py:130-180) over seeded
    parameter / buffer changes: same shadow after five updates, same weights under apply_shadow(), restore() exact — for the
    standalone class and for the one backed by the fused optimiser's flat EMA stream (parameters; the shadow is advanced with
    the optimiser's decay schedule by hand here, since the update kernel itself needs a GPU)."""
    import os
    from simvg_b200.models import ExponentialMovingAverage

    class _EmaToy(torch.nn.Module):               # the module oracle/make_golden.py::golden_ema ran the reference class on
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(6, 5)
            self.b = torch.nn.LayerNorm(5)
            self.register_buffer("running", torch.zeros(5))

    from simvg_b200.optim import FusedAdamAMSGrad
    fx = torch.load(os.path.join(golden_dir, "ema.pt"), weights_only=False)

    def replay(model, after_step):
        for step in range(fx["steps"]):
            g = torch.Generator().manual_seed(100 + step)
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(torch.randn(p.shape, generator=g) * 0.1)
                model.running.add_(torch.randn(5, generator=g))
            after_step(step)

    model = _EmaToy()
    model.load_state_dict(fx["init"])
    ema = ExponentialMovingAverage(model, fx["alpha"])
    replay(model, lambda step: ema.update_params())
    assert ema.step == fx["ema_step"] == 5
    for k, v in ema.shadow.items():
        assert torch.allclose(v, fx["shadow"][k], atol=1e-6), k
    ema.apply_shadow()
    for k, v in model.state_dict().items():
        assert torch.allclose(v, fx["applied"][k], atol=1e-6), k
    ema.restore()
    for k, v in model.state_dict().items():
        assert torch.equal(v, fx["live"][k]), k
    # fused-optimiser backing: same decay schedule (ema_decay(t)) applied to the flat shadow views
    model2 = _EmaToy()
    model2.load_state_dict(fx["init"])
    opt = FusedAdamAMSGrad(model2, lr=1e-3, ema_alpha=fx["alpha"])

    def fused_like(step):
        d = opt.ema_decay(opt.ema_t)
        with torch.no_grad():
            for p in model2.parameters():
                opt.ema_view(p).mul_(d).add_(p.detach(), alpha=1 - d)
        opt.ema_t += 1

    replay(model2, fused_like)
    for k, p in model2.named_parameters():
        assert torch.allclose(opt.ema_view(p), fx["shadow"][k], atol=1e-6), k
