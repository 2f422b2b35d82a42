"""-m gpu (slow, ~6 min): Acc@0.5 parity on held-out synthetic boxes — the north_star's accuracy criterion.

Product (sm_100a kernels) and oracle (the reference's arithmetic, eager fp32 on the same device) are trained from identical
weights on identical batches of a task whose box is recoverable from the image (tests/acc_parity.py), then evaluated on the
same 256 held-out images.  Both must actually learn the task (>= 80 % Acc@0.5: chance is ~8 %) and agree within 2 accuracy
points on both branches.

Loss level: with the learning rate decayed, the fp32 oracle keeps refining box coordinates below the resolution the product's
bf16 GEMM / attention operands allow (measured, 2000 steps: 1.36 vs 1.91 of an initial 32.6, at 100.0 % vs 99.6 % Acc@0.5) —
a precision floor, not a trajectory difference.  The third arm attributes it: the SAME oracle with the product's bf16 operand
roundings emulated in fp32 arithmetic (oracle/bf16_emulation.py, the model the per-tensor gradient tests pin the kernels to)
must end where the product ends (<= 10 %), and before the decay — where the floor does not matter yet — product and fp32
oracle must agree within 10 %.  The fp32 gap itself is reported in the JSON (kept under gpurun_out/, copied to profiles/)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

STEPS, BATCH, IMG, LAYERS = 1500, 32, 128, 2


def test_acc05_parity_on_held_out_synthetic_boxes(lib):
    import acc_parity
    out = acc_parity.run(STEPS, BATCH, IMG, LAYERS, log_every=250)
    for branch in ("acc05_decoder", "acc05_token"):
        a, b, c = out[branch]["product"], out[branch]["oracle"], out[branch]["oracle_bf16"]
        assert a >= 80.0 and b >= 80.0 and c >= 80.0, (branch, out[branch])
        assert abs(a - b) <= 2.0, (branch, out[branch])
    assert out["first_step_rel_loss_gap"] <= 1e-3, out["first_step_rel_loss_gap"]
    mid = out["loss_before_lr_decay_mean50"]
    assert abs(mid["product"] - mid["oracle"]) <= 0.10 * mid["oracle"], mid
    fin = out["final_loss_mean50"]
    assert abs(fin["product"] - fin["oracle_bf16"]) <= 0.10 * fin["oracle_bf16"], fin
