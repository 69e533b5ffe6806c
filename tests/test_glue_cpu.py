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
