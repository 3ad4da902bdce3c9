"""The CNN oracle (oracle/cnn_oracle.c) against fixtures produced by the reference itself
(tools/gen_golden.py): labels written by the unmodified use_model.py, logits of its ConvNet2."""
import os

import numpy as np

from conftest import GOLDEN


def test_labels_match_unmodified_use_model(oracle, weights):
    g = np.load(os.path.join(GOLDEN, "cnn_labels_416x240.npz"))
    lab = oracle.frame_labels(weights, g["Y"], g["U"], g["V"])
    assert lab.shape == g["labels"].shape == (28, 16)          # 7x4 CTUs, bottom row partial (zero pad)
    assert (lab == g["labels"]).all()


def test_logits_match_reference_convnet2(oracle, weights):
    g = np.load(os.path.join(GOLDEN, "cnn_logits.npz"))
    worst = 0.0
    for t64, ref in zip(g["rgb64"][::3], g["logits"][::3]):
        for q in range(4):
            oy, ox = (q // 2) * 32, (q % 2) * 32
            lg = oracle.convnet2_forward(weights, t64[:, oy:oy + 32, ox:ox + 32].copy(), t64)
            worst = max(worst, float(np.abs(lg - ref[q]).max()))
    assert worst < 1e-3, worst      # fp32 torch vs double-accumulated C: summation order only


def test_label_rules(oracle):
    # R1: zeros mixed with non-zeros become 1; R2: ones mixed become 2 (use_model.py:102-105)
    def lg(digs):
        out = np.full((4, 16), -1.0, np.float32)
        for q in range(4):
            for g in range(4):
                out[q, g * 4 + digs[q][g]] = 1.0
        return out
    lab, _ = oracle.ctu_labels(lg([[0, 0, 0, 0]] * 4))
    assert (lab == 0).all()
    lab, _ = oracle.ctu_labels(lg([[0, 1, 0, 0], [0, 0, 0, 0], [1, 1, 1, 1], [3, 0, 1, 2]]))
    # q0 -> 1111 ; q1 all-zero but label[0]!=0 -> 1111 ; q2 1111 ; q3: 0->1 then 1->2 => 3 2 2 2
    assert list(lab) == [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 3, 2, 1, 1, 2, 2]
    # argmax tie -> first maximum (torch.argmax)
    t = np.zeros((4, 16), np.float32)
    lab, mar = oracle.ctu_labels(t)
    assert (lab == 0).all() and (mar == 0).all()


def test_stage_zero_pads_outside_picture(oracle, pkg):
    Y, U, V = pkg.synth.synth_frame(416, 240, 0)
    t = oracle.stage_ctu_rgb(Y, U, V, 6, 3)       # bottom-right CTU: 32 valid cols, 48 valid rows
    assert (t[:, 48:, :] == 0).all() and (t[:, :, 32:] == 0).all() and t[:, :48, :32].any()


def test_torch_port_matches_reference_logits(host):
    """oracle/torch_ref.py (the CPU-baseline stand-in for use_model.py) against the reference's
    ConvNet2 logits."""
    import torch
    from oracle.torch_ref import TorchConvNet2
    m = TorchConvNet2(host.DEFAULT_WEIGHTS)
    g = np.load(os.path.join(GOLDEN, "cnn_logits.npz"))
    for t64, ref in zip(g["rgb64"][:6], g["logits"][:6]):
        x64 = (torch.from_numpy(t64).to(torch.float32) / 255.0)[None]
        for q in range(4):
            oy, ox = (q // 2) * 32, (q % 2) * 32
            lg = m.forward(x64[:, :, oy:oy + 32, ox:ox + 32].contiguous(), x64)[0].numpy()
            assert np.abs(lg - ref[q]).max() < 1e-4
