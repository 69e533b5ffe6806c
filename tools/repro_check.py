"""Reproducibility probe of the network-level calibration.

Calibrates the reference-generated miniatures (tests/golden/toy_net*.npz) TWICE in this process and compares,
layer by layer, bit-pattern checksums of every intermediate the layer engine hands to its probe hook (input,
FP target, attention map, activation scale, codes, A0, B0, the five inverses, loss history, best weights, layer
output).  Prints the first quantity that differs per layer and writes all checksums to
gpurun_out/repro_check_<pid>.json so that two PROCESSES can be compared as well (tools/repro_check.py --diff a b).

    python tools/repro_check.py [brats] [lits]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def checksum(t):
    import torch
    t = t.detach().contiguous()
    raw = t.reshape(-1).view(torch.uint8)
    if raw.numel() % 4 == 0:
        bits = int(raw.view(torch.int32).to(torch.int64).sum().item())
    else:
        bits = int(raw.to(torch.int64).sum().item())
    val = float(t.double().sum().item()) if t.dtype in (torch.float32, torch.float64, torch.bfloat16) else 0.0
    return [bits, val]


def run_once(task):
    import numpy as np
    import torch
    from efficientq_b200 import fold_bn, ptqer, synth
    from efficientq_b200.layer_engine import LayerCalibrator
    from tests.test_gpu_layer import build_toy
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(ROOT, "tests", "golden", "toy_net.npz" if task == "brats" else "toy_net_lits.npz"))
    model, cfg = build_toy(task)
    model.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=False)
    model.eval()
    fold_bn.search_fold_and_remove_bn(model)
    model.to(dev)
    size = cfg["size"] if isinstance(cfg["size"], tuple) else (cfg["size"],) * 3
    data = synth.batch(cfg["n"], 0, cfg["num_mod"], size, cfg["task"]).to(dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    record = []

    def probe(name, tag, t):
        record.append([name, tag, checksum(t), t.detach().float().cpu().tolist() if tag == "hist" else None])

    LayerCalibrator.probe = staticmethod(probe)
    try:
        res = ptqer.calibrate(model, data, task, ",".join(str(v) for v in cfg["init_stride"]), keep_history=True)
    finally:
        LayerCalibrator.probe = None
    losses = [float(ln.rsplit(":", 1)[1]) for ln in res["layer_loss"]]
    return record, losses


def compare(a, b, label):
    """a, b: records of two runs.  Prints, per layer, the tags whose checksums differ (in execution order)."""
    same = True
    by_layer = {}
    for ra, rb in zip(a, b):
        assert ra[0] == rb[0] and ra[1] == rb[1], (ra[:2], rb[:2])
        if ra[2][0] != rb[2][0]:
            same = False
            extra = ""
            if ra[1] == "hist" and ra[3] and rb[3]:
                first = next((i for i, (u, v) in enumerate(zip(ra[3], rb[3])) if u != v), None)
                extra = f"(first differing iterate {first}: {ra[3][first]:.9e} vs {rb[3][first]:.9e})"
            else:
                d = abs(ra[2][1] - rb[2][1])
                extra = f"(sum {ra[2][1]:.12e} vs {rb[2][1]:.12e}, |d| {d:.3e})"
            by_layer.setdefault(ra[0], []).append(f"{ra[1]} {extra}")
    print(f"== {label}: {'IDENTICAL' if same else 'DIFFERENT'}")
    for name, tags in by_layer.items():
        print(f"  {name}")
        for t in tags:
            print(f"      {t}")
    return same


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "--diff":
        ja, jb = json.load(open(sys.argv[2])), json.load(open(sys.argv[3]))
        for task in ja:
            if task in jb:
                compare(ja[task]["run1"], jb[task]["run1"], f"{task}: process A run 1 vs process B run 1")
        return
    tasks = [a for a in sys.argv[1:] if a in ("brats", "lits")] or ["brats", "lits"]
    out = {}
    for task in tasks:
        r1, l1 = run_once(task)
        r2, l2 = run_once(task)
        compare(r1, r2, f"{task}: run 1 vs run 2 in one process")
        print("  layer losses run 1:", " ".join(f"{v:.6e}" for v in l1))
        print("  layer losses run 2:", " ".join(f"{v:.6e}" for v in l2))
        out[task] = {"run1": r1, "run2": r2, "losses1": l1, "losses2": l2}
        sys.stdout.flush()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"repro_check_{os.environ.get('REPRO_TAG', os.getpid())}.json")
    json.dump(out, open(path, "w"))
    print("checksums ->", path)


if __name__ == "__main__":
    main()
