"""-m gpu (slow, ~3 min): Acc@0.5 parity on held-out synthetic boxes — the north_star's accuracy criterion.

Product (sm_100a kernels) and oracle (the reference's arithmetic, eager fp32 on the same device) are trained from identical
weights on identical batches of a task whose box is recoverable from the image (tests/acc_parity.py), then evaluated on the
same 256 held-out images.  Both must actually learn the task (>= 80 % Acc@0.5: chance is ~8 %), agree within 2 accuracy
points on both branches, and end at the same loss level (<= 5 %).  The JSON is kept under gpurun_out/ (copied to profiles/)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

STEPS, BATCH, IMG, LAYERS = 2000, 32, 128, 2


def test_acc05_parity_on_held_out_synthetic_boxes(lib):
    import acc_parity
    out = acc_parity.run(STEPS, BATCH, IMG, LAYERS, log_every=250)
    for branch in ("acc05_decoder", "acc05_token"):
        a, b = out[branch]["product"], out[branch]["oracle"]
        assert a >= 80.0 and b >= 80.0, (branch, out[branch])
        assert abs(a - b) <= 2.0, (branch, out[branch])
    lp, lo = out["final_loss_mean50"]["product"], out["final_loss_mean50"]["oracle"]
    assert abs(lp - lo) <= 0.05 * lo, out["final_loss_mean50"]
    assert out["first_step_rel_loss_gap"] <= 1e-3, out["first_step_rel_loss_gap"]
