"""hevc-deep-learning-pipeline_b200: B200-native CNN-gated intra CU-partition hot path.

Host-side mirror of the reference's sidecar interface (use_model.py / gen_frames.py) over the
C-ABI library csrc/libhevcdl.so (include/hevcdl.h).  Import with
importlib.import_module("hevc-deep-learning-pipeline_b200") (the directory name has a hyphen).
"""
from . import synth  # noqa: F401
