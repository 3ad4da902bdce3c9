"""examples/hotpath_min.c: a plain-C host program over include/hevcdl.h compiles with gcc against libhevcdl.so (no CUDA headers,
no torch), fails loudly without a device (no CPU fallback) and runs the hot path on one."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "hevc-deep-learning-pipeline_b200", "csrc")
WEIGHTS = os.path.join(ROOT, "weights", "hevc_encoder_model.hdlw")


def _build(tmp_path):
    exe = str(tmp_path / "hotpath_min")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "hotpath_min.c"),
                           "-L" + CSRC, "-lhevcdl", "-Wl,-rpath," + CSRC, "-o", exe])
    return exe


def test_c_example_builds_and_fails_loudly_without_a_device(tmp_path, built):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a device is present: covered by the gpu test")
    r = subprocess.run([exe, WEIGHTS, "416", "240", "2"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_example_runs_the_hot_path(tmp_path, built):
    exe = _build(tmp_path)
    r = subprocess.run([exe, WEIGHTS, "416", "240", "3"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-400:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("frame ")]
    assert len(lines) == 3
    for l in lines:
        m = re.match(r"frame \d+: depth labels (\d+) / (\d+) / (\d+) / (\d+), (\d+) PUs", l)
        assert m and sum(int(m.group(i)) for i in range(1, 5)) == 28 * 16 and int(m.group(5)) > 0, l
