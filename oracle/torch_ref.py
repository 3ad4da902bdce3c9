"""oracle/torch_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU port of the reference sidecar's compute as the reference executes it: the same PyTorch
library calls as ConvNet2.forward (use_model.py:48-58) -- Conv2d / BatchNorm2d in TRAINING mode /
ReLU / MaxPool2d / Linear -- at batch 1, four forwards per CTU (conv64 recomputed each time,
use_model.py:89-100), weights from the HDLW blob.  The reference's own file cannot travel to the
GPU box; this is its closest runnable stand-in and is what bench.py times as the CPU baseline.
Pinned against tests/golden/cnn_logits.npz (tests/test_oracle_cnn.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as _o

_NAMES = [("c1w", (16, 3, 5, 5)), ("c1b", (16,)), ("g1", (16,)), ("b1", (16,)),
          ("c64w", (16, 3, 5, 5)), ("c64b", (16,)), ("g64", (16,)), ("b64", (16,)),
          ("c2w", (64, 32, 3, 3)), ("c2b", (64,)), ("g2", (64,)), ("b2", (64,)),
          ("c3w", (128, 64, 3, 3)), ("c3b", (128,)), ("g3", (128,)), ("b3", (128,)),
          ("f1w", (256, 2048)), ("f1b", (256,)), ("f2w", (64, 256)), ("f2b", (64,)), ("f3w", (16, 64)), ("f3b", (16,))]


class TorchConvNet2:
    def __init__(self, hdlw_path):
        w = _o.load_weights(hdlw_path)
        o = 0
        self.p = {}
        for name, shape in _NAMES:
            n = int(np.prod(shape))
            self.p[name] = torch.from_numpy(w[o:o + n].reshape(shape).copy())
            o += n

    def _block(self, x, w, b, g, bt, pad, pool):
        x = F.conv2d(x, w, b, padding=pad)
        x = F.batch_norm(x, None, None, g, bt, training=True, eps=1e-5)     # per-sample statistics at batch 1
        return F.max_pool2d(F.relu(x), pool)

    @torch.no_grad()
    def forward(self, x32, x64):
        p = self.p
        out = torch.cat([self._block(x32, p["c1w"], p["c1b"], p["g1"], p["b1"], 2, 2),
                         self._block(x64, p["c64w"], p["c64b"], p["g64"], p["b64"], 2, 4)], dim=1)
        out = self._block(out, p["c2w"], p["c2b"], p["g2"], p["b2"], 1, 2)
        out = self._block(out, p["c3w"], p["c3b"], p["g3"], p["b3"], 1, 2)
        out = out.view(out.size(0), -1)
        out = F.relu(F.linear(out, p["f1w"], p["f1b"]))
        out = F.relu(F.linear(out, p["f2w"], p["f2b"]))
        return F.linear(out, p["f3w"], p["f3b"])

    @torch.no_grad()
    def sidecar_ctus(self, Y, U, V, ctu_begin, ctu_end, pred_dir, frame=0):
        """The sidecar's whole per-CTU job (use_model.py:86-127) for CTUs [ctu_begin, ctu_end): crop, ToTensor, four
        batch-1 forwards, label rules AND the per-CTU handshake file (written under a temporary name, then renamed)."""
        import os
        d = os.path.join(pred_dir, str(frame))
        os.makedirs(d, exist_ok=True)
        lab = self.frame_labels(Y, U, V, ctu_begin, ctu_end, _write_dir=d)
        return lab

    @torch.no_grad()
    def frame_labels(self, Y, U, V, ctu_begin, ctu_end, _write_dir=None):
        """use_model.py:86-119 for CTUs [ctu_begin, ctu_end): stage RGB (oracle K0 definition), four
        batch-1 forwards per CTU, argmax + fix-ups (C oracle for the integer rules)."""
        H, W = Y.shape
        cw = (W + 63) // 64
        labels = np.zeros((ctu_end - ctu_begin, 16), np.uint8)
        for a in range(ctu_begin, ctu_end):
            rgb = _o.stage_ctu_rgb(Y, U, V, a % cw, a // cw)
            x64 = (torch.from_numpy(rgb).to(torch.float32) / 255.0)[None]
            lg = np.zeros((4, 16), np.float32)
            for q in range(4):
                oy, ox = (q // 2) * 32, (q % 2) * 32
                lg[q] = self.forward(x64[:, :, oy:oy + 32, ox:ox + 32].contiguous(), x64)[0].numpy()
            labels[a - ctu_begin], _ = _o.ctu_labels(lg)
            if _write_dir is not None:                                  # use_model.py:121-125
                import os
                tmp = os.path.join(_write_dir, "ctu.txt")
                with open(tmp, "w") as f:
                    f.write("".join("%d " % v for v in labels[a - ctu_begin]))
                os.rename(tmp, os.path.join(_write_dir, "ctu%d.txt" % a))
        return labels
