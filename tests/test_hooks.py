"""FP-target hook (reference src/models/hooks.py:5-6) on the CPU: the stored target must be the layer's
pre-activation output, not an alias that the next unit's in-place ReLU overwrites."""
import torch
import torch.nn.functional as F


def test_fp_target_hook_copies(golden):
    from efficientq_b200 import fold_bn, ptqer, synth
    from efficientq_b200.qconv import PTQConv
    from tests.test_gpu_layer import build_toy
    g = golden("toy_net.npz")
    model, cfg = build_toy("brats")
    model.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=False)
    model.eval()
    fold_bn.search_fold_and_remove_bn(model)
    data = synth.batch(cfg["n"], 0, cfg["num_mod"], (cfg["size"],) * 3, cfg["task"])
    ptqer.set_name(model)
    ptqer.set_fp(model)
    handles = ptqer.register_fp_hooks(model)
    with torch.no_grad():
        model(data)
    for h in handles:
        h.remove()
    mods = [(n, m) for n, m in model.named_modules() if isinstance(m, PTQConv)]
    conv0 = mods[0][1]
    want = F.conv3d(data, conv0.weight, conv0.bias, conv0.stride, conv0.padding)
    assert float(conv0.output_fp.min()) < 0.0                       # pre-activation: negative entries survive
    assert torch.equal(conv0.output_fp, want)
    # blk = mid is ReLU -> conv: every conv that feeds such a unit must keep negative target entries
    assert sum(float(m.output_fp.min()) < 0.0 for _, m in mods) == len(mods)
    # and equals the reference's own run with the copying form of its hook (fixture)
    assert abs(float(conv0.output_fp.min()) - float(g["conv0_target_min"])) <= 1e-5 * abs(float(g["conv0_target_min"]))
