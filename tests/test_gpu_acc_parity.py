"""-m gpu (slow, ~3.5 min): Acc@0.5 parity on held-out synthetic boxes — the north_star's accuracy criterion.

Product (sm_100a kernels) and oracle (the reference's arithmetic, eager fp32 on the same device) are trained from identical
weights on identical batches of a task whose box is recoverable from the image (tests/acc_parity.py), then evaluated on the
same 256 held-out images.  Both must actually learn the task (>= 80 % Acc@0.5: chance is ~8 %) and agree on both branches
(measured gaps: 0.0-0.4 points on the decoder branch, 1.2-1.6 on the token branch; asserted <= 3).

Loss level: asserted only loosely.  The task's final loss is not a reproducible number — four product-only runs that differ in
nothing but kernel summation order and optimiser implementation (`tools/acc_bisect.py`: native / op-by-op head x fused / torch
Adam) ended at 1.37, 1.39, 2.03 and 2.19 (3.34-4.14 before the learning-rate decay), unrelated to either factor; the fp32
oracle ended at 1.54 and the oracle with the product's bf16 roundings emulated (`run(..., emulated=True)`) at 1.72
(`profiles/r02_acc_parity.json`).  So the first-step loss is held to 1e-3 (same weights, same batch: that IS reproducible) and the
final loss to the observed spread."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

STEPS, BATCH, IMG, LAYERS = 1500, 32, 128, 2


def test_acc05_parity_on_held_out_synthetic_boxes(lib):
    import acc_parity
    out = acc_parity.run(STEPS, BATCH, IMG, LAYERS, log_every=250, emulated=False)
    for branch in ("acc05_decoder", "acc05_token"):
        a, b = out[branch]["product"], out[branch]["oracle"]
        assert a >= 80.0 and b >= 80.0, (branch, out[branch])
        assert abs(a - b) <= 3.0, (branch, out[branch])
    assert out["first_step_rel_loss_gap"] <= 1e-3, out["first_step_rel_loss_gap"]
    fin = out["final_loss_mean50"]
    assert 0.5 * fin["oracle"] <= fin["product"] <= 2.0 * fin["oracle"], fin     # observed ratios 0.9-1.5 (see the docstring)
    assert fin["product"] <= 0.12 * out["first_loss"]["product"], (fin, out["first_loss"])     # converged: < 12 % of the initial loss
