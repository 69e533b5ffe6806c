"""Snapshot / wire formats: the three files the reference's do_ptq writes
(src/ptqer.py:383-387, src/utils/tester.py:37-51) round-trip through store_int_weight / restore_fp_weight
(src/models/PTQConv.py:125-152), plus the sub-byte packing extension.  CPU only."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import effq_oracle as O


def _model(levels):
    from efficientq_b200.qconv import EfficientQConv
    torch.manual_seed(levels)
    net = nn.Sequential(EfficientQConv(4, 6, 3, 1, 1, qlvl=levels, qlvl_act=levels),
                        nn.ReLU(), EfficientQConv(6, 5, 1, 1, 0, qlvl=levels, qlvl_act=levels))
    for m in net:
        if isinstance(m, nn.Conv3d):
            a = torch.tensor(0.1 + 0.05 * torch.rand(()).item())
            m.alpha_w.data = a
            m.weight.data = O.quantize_w(torch.randn_like(m.weight) * 0.1, a, levels)
            m.alpha_act.data = torch.tensor(1.5)
    return net


@pytest.mark.parametrize("levels", [4, 16, 256])
def test_reference_snapshot_files_round_trip(tmp_path, levels):
    from efficientq_b200 import ptqer, snapshot
    net = _model(levels)
    fp = {k: v.clone() for k, v in net.state_dict().items()}
    snapshot.save(net, os.path.join(tmp_path, "state_in_fp.pkl"))
    ptqer.store_int_weight(net)
    assert net[0].weight.dtype == torch.uint8 and int(net[0].weight.max()) <= levels - 1
    snapshot.save(net, os.path.join(tmp_path, "state_in_int8.pkl"))
    snapshot.save(net, os.path.join(tmp_path, "state_in_int8_compress.npz"), compress=True)
    snapshot.save_packed(net, os.path.join(tmp_path, "state_in_packed.npz"))
    # fp file
    sd = torch.load(os.path.join(tmp_path, "state_in_fp.pkl"))["state_dict"]
    assert all(torch.equal(sd[k], fp[k]) for k in fp)
    # int8 file: codes equal the oracle's restatement of store_int_weight; restore gives the fp weights back
    sd8 = torch.load(os.path.join(tmp_path, "state_in_int8.pkl"))["state_dict"]
    for i in (0, 2):
        want = O.weight_to_int(fp[f"{i}.weight"], fp[f"{i}.alpha_w"], levels)
        assert torch.equal(sd8[f"{i}.weight"], want)
        back = O.int_to_weight(sd8[f"{i}.weight"], sd8[f"{i}.alpha_w"], levels)
        assert torch.allclose(back, fp[f"{i}.weight"], atol=1e-6)
    # compressed npz: the reference writes the dict as one pickled object array
    z = np.load(os.path.join(tmp_path, "state_in_int8_compress.npz"), allow_pickle=True)
    inner = z["arr_0"].item()["state_dict"]
    assert np.array_equal(inner["0.weight"], sd8["0.weight"].numpy())
    # packed extension: identical state dict, fewer bytes for <= 16 levels
    sdp = snapshot.load_packed(os.path.join(tmp_path, "state_in_packed.npz"))["state_dict"]
    assert set(sdp) == set(sd8)
    for k in sd8:
        assert torch.equal(sdp[k], sd8[k]), k
    raw = np.load(os.path.join(tmp_path, "state_in_packed.npz"))
    per = {4: 4, 16: 2, 256: 1}[levels]
    assert raw["packed::0.weight"].size == -(-fp["0.weight"].numel() // per)
    # a fresh model loads the int8 state and restores fp weights (what a reference user would do)
    net2 = _model(levels)
    ptqer.store_int_weight(net2)
    net2.load_state_dict(sdp)
    for m in net2:
        if isinstance(m, nn.Conv3d):
            m.restore_fp_weight()
    assert torch.allclose(net2[0].weight.data, fp["0.weight"], atol=1e-6)


def test_pack_codes_ragged_and_range():
    from efficientq_b200 import snapshot
    for levels, n in [(4, 1), (4, 7), (16, 5), (16, 64), (256, 9)]:
        c = np.random.default_rng(n).integers(0, levels, size=n).astype(np.uint8)
        assert np.array_equal(snapshot.unpack_codes(snapshot.pack_codes(c, levels), levels, n), c)
    with pytest.raises(ValueError):
        snapshot.pack_codes(np.array([16], dtype=np.uint8), 16)
    # codes beyond levels - 1 (reference quirk: alpha_w of the last iterate) widen the layer instead of failing
    c = np.array([0, 5, 16, 3], dtype=np.uint8)
    assert snapshot.bits_needed(c, 16) == 8 and snapshot.bits_needed(c[:2], 16) == 4 and snapshot.bits_needed(c, 4) == 8
    assert np.array_equal(snapshot.unpack_codes(snapshot.pack_codes(c, 16, 8), 16, 4, 8), c)
