"""(Test infrastructure: the one script outside bench.py that runs the CPU port, hence under tests/.)  Validation of bench.py's CPU extrapolation model: BASELINE configs[0] (BraTS U-Net W4A4, 8 synthetic 4x64^3 volumes,
all 22 layers, all 200 ADMM iterations) run IN FULL with the CPU port of the reference on this host's cores, next to
what bench.cpu_sample predicts for the same job from its bounded sample (one volume, 16 of 200 iterations).

    python tests/cpu_model_check.py            # ~10 minutes on 8 cores
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from efficientq_b200 import synth  # noqa: E402
from efficientq_b200.qconv import PTQConv  # noqa: E402
from oracle import effq_oracle as O  # noqa: E402


def main():
    wl = bench.WORKLOADS["brats_w4a4_8x64"]
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    v, desc, spent, predicted = bench.cpu_sample(wl, threads=threads)
    print(f"sample: {spent:.1f} s of CPU work -> predicted {predicted:.1f} s for {wl['n']} volumes\n  {desc}", flush=True)

    model, _ = bench.build_model(wl)
    x = synth.batch(wl["n"], 0, 4, wl["size"], wl["task"])
    feats, hooks = {}, []
    for name, m in model.named_modules():
        if isinstance(m, PTQConv):
            m.set_fp()
            hooks.append(m.register_forward_hook(
                lambda mod, i, o, name=name: feats.__setitem__(name, (i[0].detach(), o.detach()))))
    with torch.no_grad():
        model(x)
    for h in hooks:
        h.remove()
    timers = {}
    t0 = time.perf_counter()
    for name, m in model.named_modules():
        if isinstance(m, PTQConv):
            xi, yo = feats.pop(name)
            O.admm_layer(xi, m.weight.data, m.bias.data, yo, m.stride, m.padding, m.qlvl_w, m.qlvl_act, m.q_act,
                         None, n_iter=200, timers=timers)
            print(f"  {name:45s} done at {time.perf_counter() - t0:7.1f} s", flush=True)
    full = time.perf_counter() - t0
    phases = {k: round(t, 1) for k, t in timers.items()}
    print(f"full run: {full:.1f} s on {threads} threads, phases(s) {json.dumps(phases)}")
    print(f"model / measured = {predicted / full:.3f}")


if __name__ == "__main__":
    main()
