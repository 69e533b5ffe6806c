"""Host-side behaviour of the glue modules (efficientq_b200/model_blk.py): on CPU tensors and under autograd they are the
stock PyTorch modules of the reference (src/models/factory_blk.py:18-93), with the fused neighbour applied explicitly."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def test_glue_modules_keep_the_reference_layout():
    """No parameters, same child names: the state-dict keys are the reference's (checked in full by test_interface)."""
    from efficientq_b200 import model_blk
    m = model_blk.GlueMaxPool3d(2, 2, fuse_relu=True)
    assert isinstance(m, nn.MaxPool3d) and not list(m.parameters())
    x = _rand((1, 2, 4, 4, 4), 7)
    assert torch.equal(m(x), F.relu(F.max_pool3d(x, 2, 2)))               # host tensors: stock ops
    u = model_blk.GlueUpsample(scale_factor=2, mode="trilinear")
    s = _rand((1, 2, 8, 8, 8), 8)
    assert torch.equal(u(x, s), F.interpolate(x, scale_factor=2, mode="trilinear") + s)


def test_down_unit_relu_is_fused_into_the_pool():
    from efficientq_b200 import model_blk
    down = model_blk._down(2, nn.Conv3d, model_blk.ReLU(True), nn.BatchNorm3d, "mid")(4, 8)
    assert down.pool.fuse_relu and isinstance(down.block.relu, model_blk.PassModule)
    ref = nn.Sequential(nn.MaxPool3d(2, 2), nn.ReLU(), down.block.conv, down.block.bn)
    x = _rand((1, 4, 4, 4, 4), 9)
    down.eval(), ref.eval()
    assert torch.equal(down(x), ref(x))
    pre = model_blk._down(2, nn.Conv3d, model_blk.ReLU(True), nn.BatchNorm3d, "pre")(4, 8)
    assert not pre.pool.fuse_relu and isinstance(pre.block.relu, nn.ReLU)


def test_oracle_glue_restatement_matches_the_library():
    """oracle/effq_oracle.py glue_* (explicit numpy fp32 arithmetic) against the reference's own stock modules."""
    from oracle import effq_oracle as O
    x = _rand((2, 3, 4, 6, 5), 21)
    x[0, 0, 0, 0, 0] = float("nan")
    y = _rand((2, 3, 4, 6, 5), 22)
    import numpy as np
    assert np.array_equal(O.glue_relu(x).numpy(), nn.ReLU()(x).numpy(), equal_nan=True)
    assert np.array_equal(O.glue_add(x, y).numpy(), (x + y).numpy(), equal_nan=True)
    for k in (2, (2, 2, 1), (1, 2, 2)):
        assert np.array_equal(O.glue_maxpool3d(x, k).numpy(), nn.MaxPool3d(k, k)(x).numpy(), equal_nan=True)
    x = _rand((2, 3, 4, 6, 5), 23)
    for f in (2, (2, 2, 1), 4, 3):
        sf = float(f) if isinstance(f, int) else tuple(float(v) for v in f)
        want = nn.Upsample(scale_factor=sf, mode="trilinear")(x)
        got = O.glue_upsample_trilinear(x, f)
        assert got.shape == want.shape
        # power-of-two factors (the only ones the configured nets use: 2 and (2, 2, 1)) have exact lambdas; with
        # factor 3 the library's fused / unfused evaluation of (1/3)(o + 0.5) - 0.5 moves a lambda by one ulp
        bar = 2.5e-7 if f != 3 else 1e-6
        assert float((got - want).abs().max()) <= bar * float(want.abs().max())
        skip = _rand(tuple(want.shape), 24)
        assert torch.equal(O.glue_upsample_trilinear(x, f, skip), got + skip)
