"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/effq_b200.h declares (no compute calls -- there is no GPU here), and the host mirror
keeps the reference's module interface."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "effq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(effq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from efficientq_b200 import build, capi
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in effq_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names          # the ctypes table binds exactly the header
    lib2 = capi.load()
    assert lib2.effq_abi_version() == 1
    assert capi.launch_count() == 0               # nothing ran: no GPU in this container


def test_struct_layouts_match_header():
    from efficientq_b200 import capi
    assert ctypes.sizeof(capi.Geom) == 15 * 4
    assert capi.SCALE_STATE_BYTES == 4 * 8 + 4 * 4
    assert capi.ADMM_STATE_BYTES == 8 + 8 * 4


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly instead of being routed to PyTorch or the oracle."""
    from efficientq_b200 import capi, ops
    with pytest.raises(capi.EffqError):
        ops.fakequant(torch.randn(16), torch.ones(1), 16, 0.0, 1.0)
    from efficientq_b200.qconv import EfficientQConv
    m = EfficientQConv(8, 8, 3, 1, 1, qlvl=16, qlvl_act=16)
    m.output_fp = torch.zeros(1, 8, 4, 4, 4)
    m.set_quantizing()
    with pytest.raises(capi.EffqError):
        m(torch.randn(1, 8, 4, 4, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "efficientq_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in txt.replace("# oracle", ""), f"{fn} mentions the oracle"


def test_module_interface_mirrors_reference():
    """Constructor signature, parameters/state-dict keys, mode machine, int export
    (reference src/models/PTQConv.py:11-175)."""
    from efficientq_b200.qconv import EfficientQConv, PTQConv
    m = EfficientQConv(4, 6, 3, 2, 1, 1, 1, False, q_weight=True, qlvl=256, q_act=False, qlvl_act=256,
                       lwq_verbose=False, lwq_batchsz=1)
    assert isinstance(m, nn.Conv3d) and isinstance(m, PTQConv) and "QConv" in type(m).__name__
    assert set(k for k, _ in m.named_parameters()) == {"weight", "alpha_act", "alpha_w"}
    assert m.alpha_act.dim() == 0 and m.alpha_w.dim() == 0
    assert (m.lwq_iter, m.lwq_rho, m.lwq_rho_max, m.lwq_eta) == (200, 10, 1000, 1)
    for setter, flags in [("set_fp", (True, False, False, False)), ("set_quantizing", (False, True, False, False)),
                          ("set_quantized", (False, False, True, False)), ("set_init_act", (False, False, False, True))]:
        getattr(m, setter)()
        assert (m._fp, m._quantizing, m._quantized, m._init_act) == flags
    m.set_fp()
    x = torch.randn(1, 4, 8, 8, 8)
    assert torch.equal(m(x), nn.functional.conv3d(x, m.weight, None, 2, 1))     # FP mode = plain conv
    # int export round trip against the oracle's restatement of PTQConv.py:125-152
    from oracle import effq_oracle as O
    m.alpha_w.data = torch.tensor(0.25)
    q = O.quantize_w(torch.randn(6, 4, 3, 3, 3) * 0.2, m.alpha_w.data, 256)
    m.weight.data = q.clone()
    m.store_int_weight()
    assert m.weight.dtype == torch.uint8
    assert torch.equal(m.weight.data, O.weight_to_int(q, torch.tensor(0.25), 256))
    m.restore_fp_weight()
    assert torch.allclose(m.weight.data, q, atol=1e-6)
    m._mode()
    with pytest.raises(RuntimeError):
        m(x)


def test_cli_and_yaml_schema():
    from efficientq_b200 import definer, entrance
    a = entrance.build_parser().parse_args(["ptq", "--qlvl_w", "4", "--qlvl_a", "4", "--round", "1", "--config",
                                            os.path.join(ROOT, "config", "lits_ptq.yaml")])
    a = entrance.merge_config(a.config, a)
    assert a.task == "lits" and a.qconv == "effq" and a.q_first == "256,-1" and a.init_stride == "2,2,1"
    Q, info, kw = definer.get_conv_class(a)
    assert info == "effq_bothQw4a4" and set(kw) == {"lwq_batchsz", "lwq_dataid", "lwq_patchsz", "lwq_verbose"}
    cube, _ = definer.get_model_cube(a, Q, kw)
    from efficientq_b200.qconv import PTQConv
    mods = [m for m in cube["model"].modules() if isinstance(m, PTQConv)]
    assert len(mods) == 28                                     # LiTS config
    assert (mods[0].qlvl_w, mods[0].q_act) == (256, False) and (mods[1].qlvl_w, mods[1].qlvl_act) == (4, 4)


def test_launch_recorder_and_replay_semantics():
    """capi.record / ops.replay (layer_engine re-issues the steady-state ADMM iterations from a recorded launch
    sequence): launches are recorded with their converted arguments and the timer tag, size / support queries
    are not, replay calls the same functions with the same arguments and raises on a non-zero return code."""
    import ctypes as C
    from efficientq_b200 import capi, ops

    calls = []

    class FakeFn:
        def __init__(self, name, restype, argtypes, rc=0):
            self.name, self.restype, self.argtypes, self.rc = name, restype, argtypes, rc

        def __call__(self, *args):
            calls.append((self.name, args))
            return self.rc

    class FakeLib:
        effq_launch_a = FakeFn("a", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p])
        effq_launch_b = FakeFn("b", C.c_int, [C.c_float, C.c_void_p])
        effq_query_workspace = FakeFn("q", C.c_int64, [C.c_int32])
        effq_thing_supported = FakeFn("s", C.c_int, [C.c_void_p])
        effq_bad = FakeFn("bad", C.c_int, [C.c_void_p], rc=3)

    rec = capi._Recorder(FakeLib())
    rec.tag = ("kernel_a", {"bytes": 8})
    assert rec.effq_launch_a(None, 5, None) == 0
    rec.tag = None
    rec.effq_query_workspace(7)
    rec.effq_thing_supported(None)
    rec.effq_launch_b(1.5, None)
    assert [c[3] for c in rec.calls] == ["effq_launch_a", "effq_launch_b"]
    assert rec.calls[0][2] == ("kernel_a", {"bytes": 8}) and rec.calls[1][2] is None
    n0 = len(calls)
    ops.replay(rec.calls)
    assert [c[0] for c in calls[n0:]] == ["a", "b"] and calls[n0][1] == (None, 5, None) and calls[n0 + 1][1] == (1.5, None)
    # a failing launch surfaces as EffqError (the message comes from the real library's effq_last_error)
    rec.effq_bad(None)
    with pytest.raises(capi.EffqError):
        ops.replay(rec.calls[-1:])
    # the recorder is only installed inside `with capi.record()`
    assert capi._recorder is None
    with capi.record() as r2:
        assert capi.load() is r2
    assert capi._recorder is None and capi.load() is not r2


def test_att_exactness_memo_follows_the_tensor_object_not_its_address():
    """ops.att_is_exact remembers its answer per tensor OBJECT and version (the caching allocator reuses the
    address of a freed attention map for the next one of the same size)."""
    import gc
    import torch
    from efficientq_b200 import ops
    ops._att_exact_cache.clear()
    a = torch.tensor([[1.0, 2.0], [3.0, 1.0]])
    assert ops.att_is_exact(None, 15) is True
    assert ops.att_is_exact(a, 15) is True and len(ops._att_exact_cache) == 1
    assert ops.att_is_exact(a, 15) is True and len(ops._att_exact_cache) == 1          # memo hit
    assert ops.att_is_exact(a, 255) is False                                             # 3 * 255 > 256
    a[0, 0] = 1.5                                                                        # in-place write: new version
    assert ops.att_is_exact(a, 15) is False
    b = torch.tensor([[0.5, 2.0]])
    assert ops.att_is_exact(b, 15) is False
    del a, b
    gc.collect()
    c = torch.tensor([[2.0, 2.0]])
    assert ops.att_is_exact(c, 15) is True                                               # dead entries are purged
    assert all(v[0]() is not None for v in ops._att_exact_cache.values())
