"""BASELINE configs[0] (BraTS-config 3D U-Net, W4A4, 8 synthetic 4x64^3 volumes) on the GPU against the full run of
the UNMODIFIED reference on the same job (tools/ref_config0.py -> tests/golden/config0.npz; 679 s on 8 CPU threads,
profiles/r02_cpu_model_check.txt).  22 layers calibrate on each other's quantised outputs, so differences compound:
the first layers must agree tightly, the deep ones to the few percent the reference moves against itself at network
level (tests/golden/toy_net.npz::ensemble_losses: up to 3.6 % on a 10-layer miniature)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_config0_matches_reference_run(golden):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import bench
    from efficientq_b200 import ptqer, synth
    g = golden("config0.npz")
    wl = bench.WORKLOADS["brats_w4a4_8x64"]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model, margs = bench.build_model(wl)
    model.to(DEV)
    data = synth.batch(wl["n"], 0, 4, wl["size"], "brats").to(DEV)
    res = ptqer.calibrate(model, data, "brats", margs.init_stride)
    names = [ln.rsplit(":", 1)[0].strip() for ln in res["layer_loss"]]
    losses = np.array([float(ln.rsplit(":", 1)[1]) for ln in res["layer_loss"]])
    ref = g["layer_losses"]
    assert names == [str(s) for s in g["layer_names"]]
    rel = np.abs(losses - ref) / ref
    lines = [f"{nm:45s} ours {a:.6e} reference {b:.6e} rel {r:.2e}" for nm, a, b, r in zip(names, losses, ref, rel)]
    lines.append(f"GPU: FP pass {res['t_fp']:.3f} s + quantizing pass {res['t_ptq']:.3f} s; unmodified reference on "
                 f"{int(g['threads'])} CPU threads: {float(g['wall_s']):.1f} s")
    print("\n".join(lines))
    import os
    if os.path.isdir("gpurun_out"):
        open("gpurun_out/r02_config0_parity.txt", "w").write("\n".join(lines) + "\n")
    mods = dict(model.named_modules())
    for nm in names[:2]:                        # layer 2's input is conv0's output: 256-level weights, loss 1e-4
        if mods[nm].q_act:
            assert abs(float(mods[nm].alpha_act) - float(g[f"alpha_act::{nm}"])) <= 1e-3 * float(g[f"alpha_act::{nm}"]), nm
    assert rel[0] <= 1e-3
    assert (rel[:3] <= 5e-3).all()
    assert (rel <= 8e-2).all() and np.median(rel) <= 3e-2
