"""efficientq_b200 -- B200-native EfficientQ PTQ calibration hot path.

Host side: Python/PyTorch mirror of the reference's per-layer quantizer-module
interface (``PTQConv`` / ``EfficientQConv``), its orchestrator ``do_ptq`` and its
``entrance.py ptq`` CLI.  Compute: hand-written sm_100a CUDA kernels behind a
C-ABI shared library (``include/effq_b200.h``), loaded with ctypes.  There is no
CPU fallback: calling any op without the built library raises.
"""
__version__ = "0.1.0"
