"""Glue ops between the quantizer layers on the repo's own kernels (reference
src/models/factoryQ.py:66-81, factory_blk.py:18-93,147-166: nn.ReLU, nn.MaxPool3d, nn.Upsample(trilinear), `+`).
The oracle is the reference's own arithmetic library on the CPU (the reference runs these as stock torch modules):
ReLU / max / add are exact operations -> bit-exact; the interpolation is held to one rounding of its fp32 result."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from efficientq_b200 import ops as _ops
    _ops.capi.load()
    return _ops


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


@pytest.mark.parametrize("shape", [(2, 8, 4, 6, 8), (1, 3, 5, 7, 9), (2, 32, 16, 16, 16)])
def test_relu_and_add_bit_exact(ops, shape):
    a, b = _rand(shape, 1), _rand(shape, 2)
    a[0, 0, 0, 0, 0] = float("nan")
    a[0, 0, 0, 0, 1] = -0.0
    ad, bd = a.to(DEV), b.to(DEV)
    assert np.array_equal(ops.relu(ad).cpu().numpy(), F.relu(a).numpy(), equal_nan=True)
    assert np.array_equal(ops.add(ad, bd).cpu().numpy(), (a + b).numpy(), equal_nan=True)
    assert np.array_equal(ops.add(ad, bd, relu_after=True).cpu().numpy(), F.relu(a + b).numpy(), equal_nan=True)
    x = ad.clone()
    y = ops.relu(x, inplace=True)
    assert y.data_ptr() == x.data_ptr()
    assert np.array_equal(x.cpu().numpy(), F.relu(a).numpy(), equal_nan=True)


@pytest.mark.parametrize("shape,k", [((2, 8, 8, 8, 8), 2), ((1, 4, 6, 10, 12), (2, 2, 1)), ((1, 3, 5, 7, 9), 2),
                                     ((2, 16, 16, 16, 16), 2), ((1, 2, 4, 4, 6), (1, 2, 2))])
def test_maxpool_bit_exact(ops, shape, k):
    x = _rand(shape, 3)
    x[0, 0, 1, 1, 1] = float("nan")
    want = F.max_pool3d(x, k, k)
    got = ops.maxpool3d(x.to(DEV), k)
    assert got.shape == want.shape
    assert np.array_equal(got.cpu().numpy(), want.numpy(), equal_nan=True)
    got_r = ops.maxpool3d(x.to(DEV), k, relu_after=True)                  # the ReLU of the "mid" unit behind the pool
    assert np.array_equal(got_r.cpu().numpy(), F.relu(want).numpy(), equal_nan=True)


@pytest.mark.parametrize("shape,f", [((2, 8, 4, 4, 4), 2), ((1, 4, 3, 5, 6), (2, 2, 1)), ((1, 3, 2, 3, 5), 2),
                                     ((2, 16, 8, 8, 8), 2), ((1, 3, 4, 4, 4), (2, 2, 2)), ((1, 2, 3, 3, 3), 4)])
def test_upsample_trilinear_matches_library(ops, shape, f):
    x = _rand(shape, 4)
    want = F.interpolate(x, scale_factor=f if isinstance(f, int) else tuple(float(v) for v in f), mode="trilinear")
    got = ops.upsample_trilinear(x.to(DEV), f)
    assert got.shape == want.shape
    err = (got.cpu() - want).abs().max().item() / want.abs().max().item()
    assert err <= 2.5e-7, err                      # fp32 op-order differences between the CPU and GPU builds only
    lib = F.interpolate(x.to(DEV), scale_factor=f if isinstance(f, int) else tuple(float(v) for v in f), mode="trilinear")
    err_lib = (got - lib).abs().max().item() / want.abs().max().item()
    assert err_lib <= 2.5e-7, err_lib
    print(f"upsample {shape} x{f}: vs CPU {err:.2e}, vs the library's CUDA kernel {err_lib:.2e} "
          f"({'bit-identical' if torch.equal(got, lib) else 'last-bit differences'})")
    skip = _rand(tuple(want.shape), 5)
    got_s = ops.upsample_trilinear(x.to(DEV), f, skip.to(DEV))
    assert torch.equal(got_s, got + skip.to(DEV))                          # the fused add is one exact rounding


def test_glue_empty_and_ragged_inputs(ops):
    """Empty batches return empty tensors without a launch; sizes that are not multiples of the vector width take the
    scalar tails (numel % 4, odd widths) and still match the library bit for bit."""
    e = torch.empty(0, 4, 2, 2, 2, device=DEV)
    assert ops.relu(e).shape == e.shape and ops.add(e, e).shape == e.shape
    assert ops.maxpool3d(e, 2).shape == (0, 4, 1, 1, 1)
    assert ops.upsample_trilinear(e, 2).shape == (0, 4, 4, 4, 4)
    assert ops.maxpool3d(torch.zeros(1, 1, 1, 1, 1, device=DEV), 2).numel() == 0        # floor mode: nothing left
    for shape in [(1, 1, 1, 1, 1), (1, 3, 1, 1, 7), (2, 1, 3, 3, 3), (1, 5, 2, 3, 5)]:
        a, b = _rand(shape, 11), _rand(shape, 12)
        assert torch.equal(ops.relu(a.to(DEV)).cpu(), F.relu(a))
        assert torch.equal(ops.add(a.to(DEV), b.to(DEV)).cpu(), a + b)
        up = ops.upsample_trilinear(a.to(DEV), (2, 2, 1)).cpu()
        want = F.interpolate(a, scale_factor=(2.0, 2.0, 1.0), mode="trilinear")
        assert up.shape == want.shape and float((up - want).abs().max()) <= 2.5e-7 * float(want.abs().max())
    with pytest.raises(Exception):
        ops.add(torch.zeros(1, 1, 2, 2, 2, device=DEV), torch.zeros(1, 1, 2, 2, 4, device=DEV))
    with pytest.raises(Exception):
        ops.relu(torch.zeros(4))                                                          # host tensor: no CPU path


def test_network_forward_on_own_glue_kernels(ops):
    """FP forward of the BraTS miniature: glue on the repo's kernels vs EFFQ_GLUE=lib (stock modules)."""
    from efficientq_b200 import model_blk
    from tests.test_gpu_layer import build_toy
    torch.manual_seed(0)
    model, cfg = build_toy("brats")
    model.eval().to(DEV)
    from efficientq_b200 import ptqer
    ptqer.set_fp(model)
    assert any(isinstance(m, model_blk.GlueMaxPool3d) and m.fuse_relu for m in model.modules())
    assert any(isinstance(m, model_blk.GlueUpsample) for m in model.modules())
    x = _rand((2, cfg["num_mod"]) + (cfg["size"],) * 3, 6).to(DEV)
    launches0 = ops.capi.launch_count()
    with torch.no_grad():
        own = model(x.clone())
    assert ops.capi.launch_count() > launches0
    os.environ["EFFQ_GLUE"] = "lib"
    try:
        with torch.no_grad():
            lib = model(x.clone())
    finally:
        del os.environ["EFFQ_GLUE"]
    err = (own - lib).abs().max().item() / lib.abs().max().item()
    assert err <= 2e-6, err
    # autograd keeps the stock (differentiable) ops
    xg = x.clone().requires_grad_(True)
    out = model(xg)
    out.sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all()
