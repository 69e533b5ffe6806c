"""Throughput of the sliding-window evaluation on one B200: BraTS-config net, synthetic
volumes one quarter-window longer than the 128^3 window (two windows per volume), FP forward (library fp32 convs) vs the
quantized deployment forward (NDHWC codes + tcgen05 conv) after a short W4A4 calibration.  Timed with CUDA events over
whole `validate_seg` passes (windows, stitching, metrics; volumes already on the device), after one warm-up pass.

    python tools/eval_bench.py [N volumes, default 4]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from efficientq_b200 import evaluate, ptqer, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    wl = dict(bench.WORKLOADS["brats_w4a4_32x128"])
    wl["n"] = 2
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model, margs = bench.build_model(wl)
    model.to(dev)
    patch, overlap = (128, 128, 128), (16, 16, 16)
    shape = (160, 128, 128)
    vols = []
    for i in range(n):
        img, lab = synth.volume(5000 + i, 4, shape, "brats")
        vols.append((f"synthetic_{5000 + i}", img.to(dev), lab.to(dev)))
    windows = len(evaluate.windows(shape, patch, overlap))
    x = synth.batch(2, 0, 4, wl["size"], "brats").to(dev)

    def timed(tag):
        evaluate.validate_seg(model, vols, dev, 3, 3, patch, overlap, "con", evaluate.split_label_brats)       # warm-up
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        sm = evaluate.validate_seg(model, vols, dev, 3, 3, patch, overlap, "con", evaluate.split_label_brats)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        vox = n * shape[0] * shape[1] * shape[2]
        print(f"{tag:28s}: {ms:8.1f} ms for {n} volumes of {shape} ({windows} windows each) = {n / ms * 1e3:6.2f} volumes/s, "
              f"{vox / ms / 1e3:8.1f} Mvoxel/s; dsc {sm[-1].metric['dsc']:.4f}")

    ptqer.set_fp(model)
    timed("FP forward (library fp32)")
    ptqer.calibrate(model, x, "brats", margs.init_stride, n_iter=10)          # short calibration: weights on their grids
    timed("quantized forward (tcgen05)")


if __name__ == "__main__":
    main()
