#!/usr/bin/env python
"""Convert the reference checkpoint (rec/hevc_encoder_model.pt, loaded at use_model.py:62) into
the flat HDLW blob the C-ABI loads (hevcdl_cfg.weights_path).

Layout: 8-byte magic b"HDLW0001", then little-endian float32 tensors in this fixed order
(BatchNorm running statistics are dropped: the reference runs BN in training mode, so they are
never read -- SURVEY.md fact 1):
  conv1.w[16,3,5,5] conv1.b[16] bn1.gamma[16] bn1.beta[16]
  conv64.w[16,3,5,5] conv64.b[16] bn64.gamma[16] bn64.beta[16]
  conv2.w[64,32,3,3] conv2.b[64] bn2.gamma[64] bn2.beta[64]
  conv3.w[128,64,3,3] conv3.b[128] bn3.gamma[128] bn3.beta[128]
  fc1.w[256,2048] fc1.b[256] fc2.w[64,256] fc2.b[64] fc3.w[16,64] fc3.b[16]
Usage: python tools/convert_weights.py /root/reference/rec/hevc_encoder_model.pt weights/hevc_encoder_model.hdlw
"""
import sys
import numpy as np
import torch

ORDER = [
    "conv1.0.weight", "conv1.0.bias", "conv1.1.weight", "conv1.1.bias",
    "conv64.0.weight", "conv64.0.bias", "conv64.1.weight", "conv64.1.bias",
    "conv2.0.weight", "conv2.0.bias", "conv2.1.weight", "conv2.1.bias",
    "conv3.0.weight", "conv3.0.bias", "conv3.1.weight", "conv3.1.bias",
    "fc1.0.weight", "fc1.0.bias", "fc2.0.weight", "fc2.0.bias", "fc3.weight", "fc3.bias",
]


def main(src, dst):
    sd = torch.load(src, map_location="cpu")
    with open(dst, "wb") as f:
        f.write(b"HDLW0001")
        n = 0
        for k in ORDER:
            a = sd[k].detach().to(torch.float32).contiguous().numpy().astype("<f4")
            f.write(a.tobytes())
            n += a.size
    print("wrote", dst, n, "floats")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
