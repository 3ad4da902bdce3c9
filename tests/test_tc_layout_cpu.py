"""CPU check of the tensor-core path's packed weights, operand layouts and descriptor addressing:
tools/tc_emulate.py replays every MMA of csrc/cnn_tc.cuh through the same shared-memory descriptors
over the blob made by tools/tc_pack.py; the result must agree with the fp32 oracle."""
import os
import sys

import numpy as np

from conftest import unsafe_label_mismatches

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_packed_blob_is_current(host):
    import tc_pack
    blob = tc_pack.pack(tc_pack.load_hdlw(host.DEFAULT_WEIGHTS))
    on_disk = open(host.DEFAULT_WEIGHTS[:-5] + ".hdlt", "rb").read()
    assert on_disk[:8] == b"HDLT0001" and on_disk[8:] == blob


def test_emulated_tensor_core_path_matches_oracle(pkg, host, oracle, weights):
    import tc_emulate
    import tc_pack
    blob = tc_emulate.Blob(tc_pack.pack(tc_pack.load_hdlw(host.DEFAULT_WEIGHTS)))
    Y, U, V = pkg.synth.synth_frame(192, 128, 5)           # 6 CTUs
    lg = tc_emulate.frame_logits(Y, U, V, blob, oracle)
    olab, olg, mar = oracle.frame_labels(weights, Y, U, V, want_logits=True)
    assert np.abs(lg - olg).max() < 0.25
    lab = np.stack([oracle.ctu_labels(l)[0] for l in lg])
    assert unsafe_label_mismatches(lab, olab, mar, 0.25)[0] == 0
