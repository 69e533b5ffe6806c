"""GPU parity tests: every C-ABI kernel against the CPU oracle and the golden fixtures
generated from the reference.  Integer / level work is bit-exact; floating-point
contractions carry their tolerance in the test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import effq_oracle as O  # noqa: E402

LEVEL_CASES = [(4, 0, 1), (16, 0, 1), (256, 0, 1), (4, -1, 1), (16, -1, 1), (256, -1, 1)]


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from efficientq_b200 import ops as _ops
    _ops.capi.load()
    return _ops


DEV = "cuda:0"


# ---------------------------------------------------------------- fake-quant (a1, a2)
@pytest.mark.parametrize("L,lo,hi", LEVEL_CASES)
def test_fakequant_matches_reference_discretize(ops, golden, L, lo, hi):
    g = golden("discretize.npz")
    v = torch.from_numpy(g[f"L{L}_lo{lo}_f32_in"]).to(DEV)
    want = g[f"L{L}_lo{lo}_f32_out"]
    y, codes = ops.fakequant(v, torch.ones(1, device=DEV), L, lo, hi, want_values=True, want_codes=True)
    assert np.array_equal(y.cpu().numpy(), want)                          # bit-exact incl. ties
    want_codes = O.discretize_codes(torch.from_numpy(g[f"L{L}_lo{lo}_f32_in"]), L, lo, hi).numpy()
    assert np.array_equal(codes.cpu().numpy().astype(np.int32), want_codes)


@pytest.mark.parametrize("L", [4, 16, 256])
def test_fakequant_module_golden(ops, golden, L):
    g = golden("fakequant_module.npz")
    x = torch.from_numpy(g[f"L{L}_x"]).to(DEV)
    w = torch.from_numpy(g[f"L{L}_w"]).to(DEV)
    a_act = torch.tensor([float(g[f"L{L}_alpha_act"])], device=DEV)
    a_w = torch.tensor([float(g[f"L{L}_alpha_w"])], device=DEV)
    qa, _ = ops.fakequant(x, a_act, L, 0.0, 1.0)
    qw, cw = ops.fakequant(w, a_w, L, -1.0, 1.0, want_codes=True)
    assert np.array_equal(qa.cpu().numpy(), g[f"L{L}_qact"])
    assert np.array_equal(qw.cpu().numpy(), g[f"L{L}_qw"])
    assert np.array_equal(cw.cpu().numpy(), g[f"L{L}_wint"])              # == store_int_weight codes


@pytest.mark.parametrize("numel", [1, 3, 4, 1023, 1 << 20, (1 << 22) + 5])
def test_fakequant_ragged_sizes(ops, numel):
    torch.manual_seed(numel)
    x = (torch.randn(numel) * 2).relu()
    alpha = torch.tensor([1.37])
    want = O.quantize_act(x, alpha[0], 16)
    y, c = ops.fakequant(x.to(DEV), alpha.to(DEV), 16, 0.0, 1.0, want_codes=True)
    assert torch.equal(y.cpu(), want)
    assert torch.equal(c.cpu().int(), O.discretize_codes(x / alpha[0], 16, 0, 1))


def test_fakequant_state_and_ndhwc_codes(ops):
    torch.manual_seed(3)
    x = torch.relu(torch.randn(2, 32, 5, 6, 7)) * 1.3
    a, b = O.project_by_iter(x, 16, 0, 1)
    st = ops.ScaleState(torch.device(DEV))
    st.set_a(a)
    y = ops.fakequant_state(x.to(DEV), st, 16, 0.0, 1.0)
    assert torch.equal(y.cpu(), a * b)                                     # fp32(a) * fp32(level)
    codes = ops.quantize_act_ndhwc(x.to(DEV), 16, state=st)
    want = O.discretize_codes(x.double() / a, 16, 0, 1).permute(0, 2, 3, 4, 1)
    assert torch.equal(codes.cpu().float().int(), want)
    # fp32 flavour (PTQConv._quantize_act arithmetic)
    alpha = torch.tensor([a], dtype=torch.float32)
    codes32 = ops.quantize_act_ndhwc(x.to(DEV), 16, alpha=alpha.to(DEV))
    want32 = O.discretize_codes(x / alpha[0], 16, 0, 1).permute(0, 2, 3, 4, 1)
    assert torch.equal(codes32.cpu().float().int(), want32)
    # e4m3 copy of the same codes, written by the same pass (<= 16 levels are exact in e4m3)
    c16, c8 = ops.quantize_act_ndhwc(x.to(DEV), 16, state=st, e4m3=True)
    assert c8.dtype == torch.float8_e4m3fn and torch.equal(c16, codes)
    assert torch.equal(c8.cpu().float().int(), want)
    none16, c8b = ops.quantize_act_ndhwc(x.to(DEV), 16, alpha=alpha.to(DEV), bf16=False, e4m3=True)
    assert none16 is None and torch.equal(c8b.cpu().float().int(), want32)
    with pytest.raises(Exception):
        ops.quantize_act_ndhwc(x.to(DEV), 256, state=st, e4m3=True)       # 256 levels do not fit e4m3


@pytest.mark.parametrize("n,c,sp,L", [(2, 32, (8, 16, 16), 16), (1, 64, (4, 8, 8), 4), (1, 16, (4, 8, 8), 16),
                                      (1, 128, (2, 8, 8), 256), (1, 512, (1, 4, 8), 16), (3, 8, (2, 4, 6), 16),
                                      (1, 256, (5, 5, 4), 4), (1, 24, (2, 4, 8), 16)])
def test_ndhwc_codes_register_transpose_kernel(ops, n, c, sp, L):
    """The register-transpose NDHWC kernel (full tiles, dhw % 16 == 0) and the shared-memory one (other
    shapes) against the oracle's fp64 / fp32 discretize, incl. values planted on and next to rounding ties
    and clamp boundaries (the fast fp32 index must hand those to the exact sequence)."""
    torch.manual_seed(c + L)
    x = torch.relu(torch.randn(n, c, *sp)) * 1.3
    a, _ = O.project_by_iter(x, L, 0, 1)
    flat = x.view(-1)
    k = torch.arange(min(L - 1, 64), dtype=torch.float64)
    ties = ((k + 0.5) / (L - 1) * a)
    planted = torch.cat([ties, ties * (1 + 1e-7), ties * (1 - 1e-7), ties * (1 + 3e-5), torch.tensor([a, a * (1 + 1e-7), 0.0])]).float()
    flat[:planted.numel()] = planted[: flat.numel()]
    st = ops.ScaleState(torch.device(DEV))
    st.set_a(a)
    want = O.discretize_codes(x.double() / a, L, 0, 1).permute(0, 2, 3, 4, 1)
    codes = ops.quantize_act_ndhwc(x.to(DEV), L, state=st)
    assert torch.equal(codes.cpu().float().int(), want)
    alpha = torch.tensor([a], dtype=torch.float32)
    want32 = O.discretize_codes(x / alpha[0], L, 0, 1).permute(0, 2, 3, 4, 1)
    codes32 = ops.quantize_act_ndhwc(x.to(DEV), L, alpha=alpha.to(DEV))
    assert torch.equal(codes32.cpu().float().int(), want32)
    if L <= 16 and c % 16 == 0:
        c16, c8 = ops.quantize_act_ndhwc(x.to(DEV), L, state=st, e4m3=True)
        assert torch.equal(c16, codes) and torch.equal(c8.cpu().float().int(), want)
        _, c8b = ops.quantize_act_ndhwc(x.to(DEV), L, alpha=alpha.to(DEV), bf16=False, e4m3=True)
        assert torch.equal(c8b.cpu().float().int(), want32)


@pytest.mark.parametrize("L,lo", [(4, 0), (16, 0), (256, 0), (16, -1), (4, -1), (256, -1), (300, 0)])
def test_fakequant_state_table_kernel_on_ties(ops, L, lo):
    """effq_fakequant_state (fp32 index + table of the L possible fp32(a)*fp32(level) values; L > 256 takes the fp64
    kernel) against the reference's op order in fp64 (layer_helper.py:25-37,67), with values planted on and next to
    every rounding tie and clamp boundary, NaN and an odd tail."""
    torch.manual_seed(L + 7 * (lo + 1))
    x = torch.randn(3 * 4099 + 3) * 1.1
    if lo == 0:
        x = torch.relu(x)
    a, _ = O.project_by_iter(x, L, lo, 1)
    delta = (1.0 - lo) / (L - 1)
    k = torch.arange(L - 1, dtype=torch.float64)
    ties = ((k + 0.5) * delta + lo) * a
    planted = torch.cat([ties, ties * (1 + 1e-7), ties * (1 - 1e-7), ties + 3e-5 * a * delta, ties - 3e-5 * a * delta,
                         torch.tensor([a, a * (1 + 1e-7), lo * a, lo * a * (1 + 1e-7), 0.0])]).float()
    x[:planted.numel()] = planted[: x.numel()]
    x[-2] = float("nan")
    st = ops.ScaleState(torch.device(DEV))
    st.set_a(a)
    y = ops.fakequant_state(x.to(DEV), st, L, float(lo), 1.0).cpu()
    want = torch.tensor(a, dtype=torch.float32) * O.discretize(x.double() / a, L, lo, 1).float()
    assert np.array_equal(y.numpy(), want.numpy(), equal_nan=True)


# ---------------------------------------------------------------- scale search (a3)
@pytest.mark.parametrize("name,lo,hi", [("act", 0, 1), ("wt", -1, 1)])
@pytest.mark.parametrize("L", [4, 16, 256])
def test_scale_search_golden(ops, golden, name, lo, hi, L):
    g = golden("project_by_iter.npz")
    v = torch.from_numpy(g[name]).to(DEV)
    st = ops.ScaleState(torch.device(DEV))
    ops.scale_search(v, L, float(lo), float(hi), st)
    s = st.read()
    a_ref = float(g[f"{name}_L{L}_a"])
    assert s["failed"] == 0 and s["converged"] == 1
    assert abs(s["a"] - a_ref) <= 1e-9 * abs(a_ref)
    assert s["passes"] == int(g[f"{name}_L{L}_passes"])
    b = ops.fakequant_state(v, st, L, float(lo), float(hi)).cpu().numpy()
    assert np.array_equal(b, np.float32(a_ref) * g[f"{name}_L{L}_b"])


@pytest.mark.parametrize("L,lo", [(16, 0.0), (4, 0.0), (16, -1.0)])
def test_scale_search_streamed_interval_passes(ops, L, lo):
    """Tensors that are streamed from memory every pass (> 3.6 M elements) use interval-stable partial
    sums: same level index for every element as a plain pass, so the scale agrees with the oracle and
    with the plain-pass kernel to fp64 summation noise, in the same number of passes."""
    torch.manual_seed(31 + L)
    n = 6_000_000
    v = torch.randn(n) * 1.3
    if lo == 0.0:
        v = torch.relu(v)
    a_ref, _, passes = O.project_by_iter(v, L, lo, 1, return_iters=True)
    vd = v.to(DEV)
    st = ops.ScaleState(torch.device(DEV))
    ops.scale_search(vd, L, lo, 1.0, st)
    s = st.read()
    diag = ops.scale_search_diag(torch.device(DEV))
    print("passes", s["passes"], diag)
    assert s["converged"] == 1 and s["passes"] == passes
    assert abs(s["a"] - a_ref) <= 1e-9 * abs(a_ref)
    assert diag["classifying_passes"] >= 1 and diag["list_passes"] >= 1
    assert diag["classifying_passes"] + diag["list_passes"] <= passes
    # plain passes only (workspace without room for the list)
    small = ops.workspace(ops.capi.load().effq_scale_search_workspace(0), torch.device(DEV))
    st2 = ops.ScaleState(torch.device(DEV))
    ops.scale_search(vd, L, lo, 1.0, st2, ws=small)
    s2 = st2.read()
    assert s2["passes"] == passes and abs(s2["a"] - s["a"]) <= 1e-12 * abs(s["a"])


def test_scale_search_strided_sum_and_multi_gpu_blocks(ops):
    """v = w*[:, :K] + dual with w* carrying a bias column (ld = K+1); and the one-pass
    building blocks used when volumes are sharded reach the same fixed point."""
    torch.manual_seed(5)
    c2, k = 24, 432
    sol = torch.randn(c2, k + 1) * 0.1
    dual = torch.randn(c2, k) * 0.01
    a_ref, b_ref, passes = O.project_by_iter(sol[:, :k] + dual, 16, -1, 1, return_iters=True)
    sol_d, dual_d = sol.to(DEV), dual.to(DEV)
    st = ops.ScaleState(torch.device(DEV))
    ops.scale_search(sol_d[:, :k], 16, -1.0, 1.0, st, v2=dual_d)
    s = st.read()
    assert abs(s["a"] - a_ref) <= 1e-9 * abs(a_ref) and s["passes"] == passes
    # multi-GPU formulation on one device: partial sums -> (all-reduce) -> step
    st2 = ops.ScaleState(torch.device(DEV))
    sums = torch.zeros(2, dtype=torch.float64, device=DEV)
    ws = ops.workspace(ops.capi.load().effq_scale_search_workspace(0), torch.device(DEV))
    v = (sol[:, :k] + dual).contiguous().to(DEV)
    ops.scale_partial(v, 16, -1.0, 1.0, st2, 0, sums, ws)
    ops.scale_step(st2, sums, 0, 16)
    for _ in range(passes + 3):
        ops.scale_partial(v, 16, -1.0, 1.0, st2, 1, sums, ws)
        ops.scale_step(st2, sums, 1, 16)
    s2 = st2.read()
    assert s2["converged"] == 1 and s2["passes"] == passes
    assert abs(s2["a"] - a_ref) <= 1e-9 * abs(a_ref)


BUCKET_CASES = [
    # numel, levels, lo, distribution          -> variant of csrc/scale_search_bucket.cu
    (96, 256, -1.0, "normal"),                 # direct evaluation, one CTA (final_cls)
    (3456, 256, -1.0, "normal"),               # direct evaluation (conv0)
    (5000, 16, -1.0, "normal"),                # direct evaluation, below the bucket threshold
    (40000, 256, -1.0, "normal"),              # direct evaluation on a cluster (many levels)
    (27648, 16, -1.0, "normal"),               # bucketed, one CTA, slice in shared memory
    (32768, 4, -1.0, "normal"),                # bucketed, 4 levels (LiTS)
    (110592, 16, -1.0, "normal"),              # bucketed, cluster of 4
    (442368, 16, -1.0, "laplace"),             # bucketed, cluster of 16 (or 8 with the slices in global memory)
    (1769472, 16, -1.0, "normal"),             # bucketed, slices sorted into the workspace
    (1769472, 16, -1.0, "outliers"),           # values far outside the bucket range (end buckets)
    (300000, 16, 0.0, "relu"),                 # activation-like: half the elements exactly zero
    (65536, 16, -1.0, "lattice"),              # many elements exactly ON rounding thresholds of the final scale
]


@pytest.mark.parametrize("numel,L,lo,dist", BUCKET_CASES)
def test_scale_search_bucketed(ops, numel, L, lo, dist, monkeypatch):
    """The bucket-sorted search (experimental variant, EFFQ_SS_BUCKET=1: O(levels) work per pass, integer sums) against
    the oracle: same number of passes, scale to 1e-9 (fixed-point rounding of v: ~1e-12), and bit-identical between two
    launches."""
    monkeypatch.setenv("EFFQ_SS_BUCKET", "1")
    g = torch.Generator().manual_seed(numel + L)
    if dist == "normal":
        v = torch.randn(numel, generator=g) * 0.037
    elif dist == "laplace":
        v = torch.distributions.Laplace(0.0, 0.02).sample((numel,))
    elif dist == "outliers":
        v = torch.randn(numel, generator=g) * 0.01
        v[::1000] *= 300.0
    elif dist == "relu":
        v = torch.relu(torch.randn(numel, generator=g)) * 1.7
    else:
        a0, _ = O.project_by_iter(torch.randn(numel, generator=g), L, lo, 1)
        v = torch.randn(numel, generator=g)
        ties = ((torch.arange(L - 1, dtype=torch.float64) + 0.5) * 2 / (L - 1) - 1) * a0          # thresholds of a nearby scale
        v[: 50 * (L - 1)] = ties.repeat(50).float()
    a_ref, _, passes = O.project_by_iter(v, L, lo, 1, return_iters=True)
    vd = v.to(DEV)
    res = []
    for _ in range(2):
        st = ops.ScaleState(torch.device(DEV))
        ops.scale_search(vd, L, lo, 1.0, st)
        res.append(st.read())
    s = res[0]
    assert s["failed"] == 0 and s["converged"] == 1
    assert s["passes"] == passes, (s["passes"], passes)
    assert abs(s["a"] - a_ref) <= 1e-9 * abs(a_ref), (s["a"], a_ref)
    assert res[1]["a"] == s["a"] and res[1]["s_bv"] == s["s_bv"] and res[1]["s_bb"] == s["s_bb"]


def test_scale_search_bucketed_strided_and_degenerate(ops, monkeypatch):
    monkeypatch.setenv("EFFQ_SS_BUCKET", "1")
    torch.manual_seed(9)
    c2, k = 128, 3456                                               # 442 K elements, rows with a bias column, v = w* + dual
    sol = torch.randn(c2, k + 1) * 0.05
    dual = torch.randn(c2, k) * 0.004
    a_ref, _, passes = O.project_by_iter(sol[:, :k] + dual, 16, -1, 1, return_iters=True)
    sol_d, dual_d = sol.to(DEV), dual.to(DEV)
    st = ops.ScaleState(torch.device(DEV))
    ops.scale_search(sol_d[:, :k], 16, -1.0, 1.0, st, v2=dual_d)
    s = st.read()
    assert abs(s["a"] - a_ref) <= 1e-9 * abs(a_ref) and s["passes"] == passes
    z = torch.zeros(20000, device=DEV)
    ops.scale_search(z, 16, -1.0, 1.0, st)
    s = st.read()
    assert s["a"] == 0.0 and s["failed"] == 0
    z[7] = float("nan")
    ops.scale_search(z, 16, -1.0, 1.0, st)
    s = st.read()
    assert s["a"] != s["a"] and s["passes"] == 0                    # the reference's a0 is NaN and its loop never runs


def test_scale_search_large_activation(ops):
    torch.manual_seed(6)
    x = torch.relu(torch.randn(2, 32, 32, 32, 32))                 # 2.1 M elements, many CTAs
    a_ref, _, passes = O.project_by_iter(x, 16, 0, 1, return_iters=True)
    st = ops.ScaleState(torch.device(DEV))
    ops.scale_search(x.to(DEV), 16, 0.0, 1.0, st)
    s = st.read()
    assert abs(s["a"] - a_ref) <= 1e-9 * abs(a_ref)
    assert abs(s["passes"] - passes) <= 1


# ---------------------------------------------------------------- FP conv on the tcgen05 kernel (f.4)
FP_CASES = [
    # n, c1, c2, k, spatial
    (2, 32, 32, 3, (16, 16, 16)),
    (1, 64, 64, 3, (8, 24, 16)),
    (2, 128, 128, 3, (8, 8, 8)),
    (1, 256, 256, 3, (4, 8, 8)),
    (2, 32, 64, 1, (8, 16, 16)),
    (2, 256, 128, 1, (4, 8, 8)),
    (1, 16, 48, 3, (8, 8, 12)),               # ragged tiles, C2 not a multiple of 32 (epilogue loads the partial sums itself)
]


@pytest.mark.parametrize("n,c1,c2,k,sp", FP_CASES)
def test_conv3d_fp_matches_fp64(ops, n, c1, c2, k, sp):
    """The FP pass's convolution on the tcgen05 kernel (eight products of fixed-point digit planes, exact integer
    accumulation) against an fp64 conv (reference src/ptqer.py:333-335 computes its targets with F.conv3d in fp32).
    Tolerance 5e-7 of max|out|, independent of K: what is left is the 2^-24 fixed-point resolution relative to the
    channel maximum (the inputs here have channels of very different magnitude) and one fp32 rounding.  Measured
    2.8e-7 .. 4.1e-7; the library's fp32 conv on the same inputs: 1.8e-7 .. 2.2e-6 depending on its algorithm."""
    torch.manual_seed(n * 1000 + c1 + c2 + k)
    x = (torch.randn(n, c1, *sp) * torch.rand(1, c1, 1, 1, 1) * 3).to(DEV)          # channels of different magnitude
    x[:, : c1 // 2] = torch.relu(x[:, : c1 // 2])
    w = (torch.randn(c2, c1, k, k, k) / (c1 * k ** 3) ** 0.5).to(DEV)
    b = torch.randn(c2).to(DEV)
    assert ops.conv3d_fp_supported(x.shape, c2, (k,) * 3, (1, 1, 1), (k // 2,) * 3)
    out = ops.conv3d_fp(x, w, b, (k,) * 3)
    ref64 = torch.nn.functional.conv3d(x.double(), w.double(), b.double(), 1, k // 2)
    torch.backends.cudnn.allow_tf32 = False
    lib32 = torch.nn.functional.conv3d(x, w, b, 1, k // 2)
    scale = float(ref64.abs().max())
    e_own = float((out.double() - ref64).abs().max()) / scale
    e_lib = float((lib32.double() - ref64).abs().max()) / scale
    print(f"conv3d_fp c{c1}x{c2}k{k}: max err / max|out|  own {e_own:.2e}  library fp32 {e_lib:.2e}")
    assert e_own <= 5e-7
    # every plane product is exact and the sums are fp32: two runs are bit-identical
    assert torch.equal(out, ops.conv3d_fp(x, w, b, (k,) * 3))


# ---------------------------------------------------------------- conv + squared error (a10)
CONV_CASES = [
    # n, c1, c2, k, s, p, spatial
    (2, 4, 32, 3, 2, 1, (16, 16, 16)),     # conv0-like
    (2, 32, 3, 1, 1, 0, (8, 16, 8)),       # final_cls-like
    (1, 16, 16, 3, 1, 1, (5, 7, 9)),       # ragged
    (2, 32, 64, 1, 1, 0, (4, 6, 10)),
    (1, 8, 40, 3, 1, 1, (6, 6, 6)),
]


@pytest.mark.parametrize("n,c1,c2,k,s,p,sp", CONV_CASES)
def test_conv3d_f32_and_sse(ops, n, c1, c2, k, s, p, sp):
    torch.manual_seed(c1 * 100 + c2)
    x = torch.randn(n, c1, *sp)
    w = torch.randn(c2, c1, k, k, k) * 0.1
    b = torch.randn(c2)
    want = F.conv3d(x, w, b, s, p)
    target = want + 0.05 * torch.randn_like(want)
    att = torch.rand(n, *want.shape[2:]) + 0.5
    out, sse = ops.conv3d_f32(x.to(DEV), w.to(DEV), b.to(DEV), s, p, want_out=True, target=target.to(DEV),
                              att=att.to(DEV))
    torch.testing.assert_close(out.cpu(), want, rtol=1e-4, atol=1e-4)          # fp32 conv, different summation order
    ref = (att.unsqueeze(1).double() * (want.double() - target.double()) ** 2).sum().item()
    assert abs(sse.item() - ref) <= 1e-4 * ref
    _, sse2 = ops.conv3d_f32(x.to(DEV), w.to(DEV), b.to(DEV), s, p, want_out=False, target=target.to(DEV))
    ref2 = ((want.double() - target.double()) ** 2).sum().item()
    assert abs(sse2.item() - ref2) <= 1e-4 * ref2


TC_CASES = [
    # n, c1, c2, k, spatial, La, Lw
    (1, 32, 32, 1, (2, 16, 8), 16, 16),
    (1, 32, 32, 3, (3, 16, 8), 16, 16),
    (2, 32, 32, 3, (4, 16, 16), 16, 16),
    (1, 64, 64, 3, (4, 16, 8), 16, 16),
    (1, 32, 64, 1, (4, 8, 8), 16, 16),
    (1, 128, 128, 3, (2, 8, 8), 16, 16),     # two channel groups, streamed weights, ragged tile (H=8)
    (1, 256, 128, 1, (2, 8, 8), 4, 4),
    (1, 16, 48, 3, (3, 10, 12), 256, 256),   # ragged H/W, N=48 (x16 TMEM load), 256 levels
    (2, 64, 32, 3, (5, 20, 12), 4, 16),
    (1, 32, 32, 3, (3, 12, 10), 16, 16),     # W % 4 != 0: the epilogue loads the target itself (no target TMA)
    (1, 32, 64, 3, (4, 24, 20), 16, 16),     # two target boxes per tile, ragged H tile
    (1, 128, 512, 3, (2, 5, 4), 4, 4),       # C2 > 256: two chunks of 256 output channels (LiTS deepest level)
    (2, 64, 512, 1, (2, 8, 8), 16, 16),      # same with the target TMA ring (W % 4 == 0)
]


@pytest.mark.parametrize("n,c1,c2,k,sp,la,lw", TC_CASES)
def test_conv3d_tc_exact_on_codes(ops, n, c1, c2, k, sp, la, lw):
    """tcgen05 conv: integer codes in, exact integer accumulation -> must equal the fp32
    reference conv on the same codes up to the final scale/bias rounding."""
    assert ops.conv3d_tc_supported((n, c1, *sp), c2, k, 1, (k - 1) // 2)
    torch.manual_seed(c1 + c2 + k)
    xc = torch.randint(0, la, (n, c1, *sp)).float()
    wc = (2 * torch.randint(0, lw, (c2, c1, k, k, k)) - (lw - 1)).float()
    b = torch.randn(c2)
    scale = 0.0123
    want = F.conv3d(xc.double(), wc.double(), None, 1, (k - 1) // 2).float() * scale + b.view(1, -1, 1, 1, 1)
    target = want + 0.1 * torch.randn_like(want)
    att = torch.rand(n, *sp) + 0.5
    xq = xc.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16).to(DEV)
    wq = ops.pack_weight_codes(wc.to(DEV))
    cs = torch.tensor([scale], dtype=torch.float32, device=DEV)
    out, sse = ops.conv3d_tc(xq, wq, b.to(DEV), cs, c2, k, want_out=True, target=target.to(DEV), att=att.to(DEV))
    torch.cuda.synchronize()
    torch.testing.assert_close(out.cpu(), want, rtol=2e-6, atol=2e-6)
    ref = (att.unsqueeze(1).double() * (want.double() - target.double()) ** 2).sum().item()
    assert abs(sse.item() - ref) <= 1e-5 * ref
    # no output written, unweighted: the per-iteration scoring call
    _, sse2 = ops.conv3d_tc(xq, wq, b.to(DEV), cs, c2, k, want_out=False, target=target.to(DEV))
    ref2 = ((want.double() - target.double()) ** 2).sum().item()
    assert abs(sse2.item() - ref2) <= 1e-5 * ref2


FP8_CASES = [
    # n, c1, c2, k, spatial, La, Lw   (c1 = 32 -> 32 B rows / SW32, 64 -> SW64, 128 -> SW128, 256 -> two blocks)
    (1, 32, 32, 3, (3, 16, 8), 16, 16),
    (2, 32, 32, 1, (4, 16, 16), 16, 16),
    (1, 64, 64, 3, (4, 16, 8), 16, 16),
    (1, 64, 32, 3, (5, 20, 12), 4, 16),
    (1, 128, 128, 3, (2, 8, 8), 16, 16),
    (1, 128, 64, 1, (3, 9, 7), 16, 4),
    (1, 256, 128, 3, (2, 8, 8), 16, 16),
    (1, 256, 48, 1, (2, 10, 8), 4, 4),
    (1, 128, 512, 3, (2, 8, 8), 4, 4),       # chunked C2
]


@pytest.mark.parametrize("n,c1,c2,k,sp,la,lw", FP8_CASES)
def test_conv3d_tc_e4m3_bit_identical_to_bf16(ops, n, c1, c2, k, sp, la, lw):
    """<= 16 levels: the e4m3 operand path (kind::f8f6f4, K = 32) must give exactly the bits of the
    bf16 path (kind::f16) -- same exact integer sums, same epilogue -- and of the fp64 reference."""
    assert ops.conv3d_tc_supported((n, c1, *sp), c2, k, 1, (k - 1) // 2, ops.CODE_E4M3)
    torch.manual_seed(c1 * 7 + c2 + k)
    xc = torch.randint(0, la, (n, c1, *sp)).float()
    wc = (2 * torch.randint(0, lw, (c2, c1, k, k, k)) - (lw - 1)).float()
    b = torch.randn(c2).to(DEV)
    cs = torch.tensor([0.0123], dtype=torch.float32, device=DEV)
    target = torch.randn(n, c2, *sp).to(DEV)
    att = (torch.rand(n, *sp) + 0.5).to(DEV)
    xcl = xc.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    x16, x8 = xcl.to(torch.bfloat16), xcl.to(torch.float8_e4m3fn)
    w16 = ops.pack_weight_codes(wc.to(DEV))
    w8 = ops.pack_weight_codes(wc.to(DEV), ops.CODE_E4M3)
    o16, s16 = ops.conv3d_tc(x16, w16, b, cs, c2, k, want_out=True, target=target, att=att)
    o8, s8 = ops.conv3d_tc(x8, w8, b, cs, c2, k, want_out=True, target=target, att=att)
    torch.cuda.synchronize()
    assert torch.equal(o8, o16)
    assert s8.item() == s16.item()
    want = F.conv3d(xc.double(), wc.double(), None, 1, (k - 1) // 2).float() * 0.0123 + b.cpu().view(1, -1, 1, 1, 1)
    torch.testing.assert_close(o8.cpu(), want, rtol=2e-6, atol=2e-6)
    with pytest.raises(Exception):
        ops.conv3d_tc(x8, w16, b, cs, c2, k)                 # mixed operand types are refused


def test_conv3d_tc_e4m3_unsupported_shapes(ops):
    assert not ops.conv3d_tc_supported((1, 16, 4, 16, 8), 32, 3, 1, 1, ops.CODE_E4M3)     # 16 B rows
    assert not ops.conv3d_tc_supported((1, 192, 4, 16, 8), 32, 3, 1, 1, ops.CODE_E4M3)    # not a multiple of 128
    assert ops.conv3d_tc_supported((1, 192, 4, 16, 8), 32, 3, 1, 1, ops.CODE_BF16)


def test_conv3d_tc_linearity_at_scale(ops):
    """Full-size property check (too big for the CPU oracle): conv is linear in the weights,
    and agrees with the generic fp32 kernel on the same codes."""
    torch.manual_seed(9)
    n, c, sp = 2, 32, (32, 64, 64)
    xc = torch.randint(0, 16, (n, c, *sp), device=DEV).float()
    w1 = (2 * torch.randint(0, 16, (c, c, 3, 3, 3), device=DEV) - 15).float()
    xq = xc.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)
    cs = torch.ones(1, device=DEV)
    o1, _ = ops.conv3d_tc(xq, ops.pack_weight_codes(w1), None, cs, c, 3)
    o2, _ = ops.conv3d_tc(xq, ops.pack_weight_codes(-w1), None, cs, c, 3)
    assert torch.equal(o1, -o2)
    ref, _ = ops.conv3d_f32(xc, w1, None, 1, 1)
    assert torch.equal(o1, ref)                      # integers < 2^24: both paths exact
    o8, _ = ops.conv3d_tc(xq.to(torch.float8_e4m3fn), ops.pack_weight_codes(w1, ops.CODE_E4M3), None, cs, c, 3)
    assert torch.equal(o8, ref)


# ---------------------------------------------------------------- normal equations (a7, a8)
@pytest.mark.parametrize("name", ["k3s1p1", "k3s2p1", "k1s1p0"])
def test_gram_golden(ops, golden, name):
    g = golden("solver.npz")
    k, s, p = [int(t) for t in g[f"{name}_geom"]]
    x = torch.from_numpy(g[f"{name}_x"]).to(DEV)
    y = torch.from_numpy(g[f"{name}_y"]).to(DEV)
    att = torch.from_numpy(g[f"{name}_att"]).to(DEV)
    a0, b0 = ops.gram(x, y, att, (k, k, k), s, p, has_bias=True)
    np.testing.assert_allclose(a0.cpu().numpy(), g[f"{name}_A0"], rtol=2e-5, atol=1e-4)
    np.testing.assert_allclose(b0.cpu().numpy(), g[f"{name}_B0"], rtol=2e-5, atol=1e-4)
    # symmetry (size-independent property)
    assert torch.allclose(a0, a0.T, rtol=1e-6, atol=1e-6)


def test_gram_codes_with_scale_and_no_bias(ops):
    torch.manual_seed(12)
    xc = torch.randint(0, 16, (2, 8, 6, 8, 8)).float()
    y = torch.randn(2, 5, 6, 8, 8)
    sc = 0.37
    a0, b0 = ops.gram(xc.to(DEV), y.to(DEV), None, (3, 3, 3), 1, 1, has_bias=False,
                      x_scale=torch.tensor([sc], device=DEV))
    cols = O.im2col(xc * sc, 3, 3, 3, 1, 1).double()
    ym = y.permute(1, 0, 2, 3, 4).reshape(5, -1).double()
    np.testing.assert_allclose(a0.cpu().numpy(), (2 * cols @ cols.T).float().numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(b0.cpu().numpy(), (2 * ym @ cols.T).float().numpy(), rtol=1e-5, atol=1e-4)


GRAM_TC_CASES = [
    # n, c1, c2, spatial, levels, integer att
    (2, 32, 32, (6, 16, 8), 16, True),
    (1, 64, 64, (4, 8, 8), 16, False),
    (1, 16, 24, (5, 10, 12), 4, False),       # ragged 8x8 blocks, partial 128/256-row blocks
    (1, 128, 16, (2, 8, 8), 16, True),
    (3, 32, 8, (3, 9, 7), 256, False),
    (2, 256, 16, (3, 5, 4), 4, True),         # LiTS deep level: volume smaller than one 8x8 block, V << K' = 6913
]


@pytest.mark.parametrize("n,c1,c2,sp,la,int_att", GRAM_TC_CASES)
def test_gram_tc_matches_generic_and_oracle(ops, n, c1, c2, sp, la, int_att):
    """tcgen05 Gram (K x K block on codes, att via bf16 hi+lo split) + generic rows vs the fp64
    im2col oracle.  Tolerance: 2^-17 per weighted term for fractional att, exact for integer att."""
    torch.manual_seed(c1 + c2)
    codes = torch.randint(0, la, (n, c1, *sp)).float()
    sc = 0.173
    x = codes * sc
    y = torch.randn(n, c2, *sp)
    att = (torch.rand(n, *sp) * 3).floor() + 1 if int_att else torch.rand(n, *sp) * 2 + 0.25
    xq = codes.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16).to(DEV)
    a0, b0, _, flag = ops.gram_tc(xq, torch.tensor([sc], device=DEV), y.to(DEV), att.to(DEV), True)
    assert int(flag.item()) == 0
    cols = O.im2col(x, 3, 3, 3, 1, 1).double()
    cols = torch.cat([cols, torch.ones(1, cols.shape[1], dtype=torch.float64)], 0)
    wcols = cols * att.reshape(1, -1).double()
    ym = y.permute(1, 0, 2, 3, 4).reshape(c2, -1).double()
    a_ref, b_ref = (2 * cols @ wcols.T), (2 * ym @ wcols.T)
    tol = 1e-6 if int_att else 2e-5
    scale = a_ref.abs().max().item()
    assert (a0.cpu().double() - a_ref).abs().max().item() <= tol * scale
    assert (b0.cpu().double() - b_ref).abs().max().item() <= 2e-5 * b_ref.abs().max().item()
    assert torch.equal(a0, a0.T) or torch.allclose(a0, a0.T, rtol=1e-6, atol=1e-6 * scale)
    g0, gb0 = ops.gram(x.to(DEV), y.to(DEV), att.to(DEV), (3, 3, 3), 1, 1, has_bias=True)
    assert (a0 - g0).abs().max().item() <= tol * scale
    assert torch.allclose(b0, gb0, rtol=1e-5, atol=1e-5 * b_ref.abs().max().item())
    # single-term mode: chosen when every att * code is exact in bf16 (integer masks) -> identical bits
    att_d = att.to(DEV)
    exact = ops.att_is_exact(att_d, la - 1)
    assert exact == (int_att and 3 * (la - 1) <= 256)
    if exact:
        a1, b1, _, flag1 = ops.gram_tc(xq, torch.tensor([sc], device=DEV), y.to(DEV), att_d, True, att_exact=True)
        assert int(flag1.item()) == 0
        assert torch.equal(a1, a0) and torch.equal(b1, b0)


@pytest.mark.parametrize("n,c1,c2,k,s,p,sp", [(2, 4, 32, 3, 2, 1, (16, 16, 16)), (2, 32, 3, 1, 1, 0, (8, 16, 8))])
def test_quadform_scoring_equals_conv_sse(ops, n, c1, c2, k, s, p, sp):
    """Conv-free scoring of the un-quantised-input layers: fp64 statistics + quadratic form must
    reproduce sum((conv(x,G)+b-y)^2) (here against an fp64 conv on the CPU)."""
    torch.manual_seed(17)
    x = torch.randn(n, c1, *sp)
    w0 = torch.randn(c2, c1, k, k, k) * 0.1
    b0 = torch.randn(c2) * 0.1
    y = F.conv3d(x, w0, b0, s, p)
    acc = ops.gram_f64(x.to(DEV), y.to(DEV), (k, k, k), s, p, has_bias=True)
    yy = (y.double() ** 2).sum().item()
    ws = ops.workspace(16 + 8 * 1024, torch.device(DEV))
    sse = torch.zeros(1, dtype=torch.float64, device=DEV)
    for eps in (1e-2, 1e-3):                      # quantisation-sized perturbations of the weights
        g = w0 + eps * torch.randn_like(w0)
        b = b0 + eps * torch.randn_like(b0)
        ref = ((F.conv3d(x.double(), g.double(), b.double(), s, p) - y.double()) ** 2).sum().item()
        ops.quadform_sse(acc, yy, g.reshape(c2, -1).contiguous().to(DEV), b.to(DEV), sse, ws)
        assert abs(sse.item() - ref) <= 1e-4 * ref, (sse.item(), ref)


# ---------------------------------------------------------------- ADMM update (a9, a11)
def test_admm_elementwise_kernels(ops):
    torch.manual_seed(13)
    c2, c1, taps = 16, 16, 27
    k = c1 * taps
    kp = k + 1
    b0, w0p = torch.randn(c2, kp), torch.randn(c2, kp)
    g, dual = torch.randn(c2, k) * 0.1, torch.randn(c2, k) * 0.01
    rho, eta = 12.5, 1.25
    want_b = b0 + eta * w0p
    want_b[:, :k] += rho * (g - dual)
    out = torch.empty(c2, kp, device=DEV)
    ops.admm_rhs(b0.to(DEV), w0p.to(DEV), g.to(DEV), dual.to(DEV), rho, eta, out)
    assert torch.equal(out.cpu(), want_b)
    a0 = torch.randn(kp, kp)
    qe = torch.eye(kp)
    qe[-1, -1] = 0
    want_a = a0 + rho * qe + eta * torch.eye(kp)
    aout = torch.empty(kp, kp, device=DEV)
    ops.admm_lhs(a0.to(DEV), rho, eta, True, aout)
    assert torch.equal(aout.cpu(), want_a)

    # projection: G = a*b, dual = (w* - G + dual)/2, codes in the tensor-core layout
    sol = torch.randn(c2, kp) * 0.1
    a_w, b_w = O.project_by_iter(sol[:, :k] + dual, 16, -1, 1)
    g_ref = a_w * b_w
    dual_ref = (sol[:, :k] - g_ref + dual) / 2
    dev = torch.device(DEV)
    wst, xst, st = ops.ScaleState(dev), ops.ScaleState(dev), ops.AdmmState(dev)
    xst.set_a(2.5)
    sol_d, dual_d = sol.to(DEV), dual.to(DEV)
    ops.scale_search(sol_d[:, :k], 16, -1.0, 1.0, wst, v2=dual_d)
    g_d, bstar = torch.empty(c2, k, device=DEV), torch.empty(c2, device=DEV)
    wcodes = torch.empty(taps * (c1 // 8) * c2 * 8, dtype=torch.bfloat16, device=DEV)
    # fused assembly of the NEXT right-hand side (what admm_rhs would produce from the new G / dual)
    b0n, w0n = torch.randn(c2, c1 * taps + 1, device=DEV), torch.randn(c2, c1 * taps + 1, device=DEV)
    nplanes = torch.empty((3, c2, ops.split3_ld(c1 * taps + 1)), dtype=torch.bfloat16, device=DEV)
    ops.admm_project(sol_d, dual_d, wst, xst, 16, 16, c2, c1, taps, True, 2.0, g_d, bstar, wcodes, st,
                     next_rhs=(b0n, w0n, 20.0, 1.5, nplanes))
    want_planes = torch.empty_like(nplanes)
    ops.admm_rhs(b0n, w0n, g_d, dual_d, 20.0, 1.5, None, planes=want_planes)
    assert torch.equal(nplanes, want_planes)
    assert torch.equal(g_d.cpu(), g_ref)
    assert torch.equal(dual_d.cpu(), dual_ref)
    assert torch.equal(bstar.cpu(), sol[:, k])
    codes_ref = (2 * O.discretize_codes((sol[:, :k] + dual).double() / a_w, 16, -1, 1) - 15).float()
    assert torch.equal(wcodes.cpu().float(), ops.pack_weight_codes(codes_ref.view(c2, c1, 3, 3, 3).to(DEV)).cpu().float().flatten())
    s = st.read()
    assert abs(s["a_w"] - np.float32(a_w)) == 0
    want_cs = np.float32(np.float64(np.float32(2.5)) / 15 * np.float64(np.float32(a_w)) / 15)
    assert abs(s["conv_scale"] - want_cs) <= 1e-7 * want_cs

    # best-iterate tracking: strict '<', iterate 0 always seeds
    best_g, best_b = torch.zeros(c2, k, device=DEV), torch.zeros(c2, device=DEV)
    hist = torch.zeros(4, device=DEV)
    sse = torch.zeros(1, dtype=torch.float64, device=DEV)
    st.reset()
    for it, val in enumerate([5.0, 7.0, 5.0, 3.0]):
        sse.fill_(val * 100)
        g_d.fill_(float(it))
        ops.admm_track(st, sse, 100.0, g_d, bstar, best_g, best_b, hist)
        s = st.read()
        assert s["iter"] == it + 1
        assert s["best_iter"] == (0 if it < 3 else 3)
        assert best_g[0, 0].item() == (0.0 if it < 3 else 3.0)
    assert hist.cpu().tolist() == [5.0, 7.0, 5.0, 3.0]


def test_project_keep_and_decide_equal_track(ops):
    """The loop's fused form -- effq_admm_decide (or the tail of effq_quadform_delta) scores, the NEXT
    effq_admm_project saves the best iterate before overwriting it, effq_admm_keep flushes the last one -- must
    leave the same best iterate, codes, history and state as effq_admm_track after every iterate."""
    torch.manual_seed(31)
    c2, c1, taps = 16, 16, 27
    k = c1 * taps
    dev = torch.device(DEV)
    losses = [5.0, 7.0, 4.0, 4.0, 3.0, 6.0]

    def run(fused):
        wst, xst, st = ops.ScaleState(dev), ops.ScaleState(dev), ops.AdmmState(dev)
        xst.set_a(2.5)
        g, bstar = torch.zeros(c2, k, device=DEV), torch.zeros(c2, device=DEV)
        wcodes = torch.zeros(taps * c1 * c2, dtype=torch.bfloat16, device=DEV)
        best_g, best_b, best_w = torch.zeros_like(g), torch.zeros_like(bstar), torch.zeros_like(wcodes)
        dual = torch.zeros(c2, k, device=DEV)
        hist = torch.zeros(len(losses), device=DEV)
        sse = torch.zeros(1, dtype=torch.float64, device=DEV)
        gen = torch.Generator().manual_seed(5)
        snaps = []
        for it, val in enumerate(losses):
            sol = (torch.randn(c2, k + 1, generator=gen) * 0.1).to(DEV)
            ops.scale_search(sol[:, :k], 16, -1.0, 1.0, wst, v2=dual)
            ops.admm_project(sol, dual, wst, xst, 16, 16, c2, c1, taps, True, 1.0, g, bstar, wcodes, st,
                             keep=(best_g, best_b, best_w) if fused else None)
            sse.fill_(val * 100)
            if fused:
                ops.admm_decide(st, sse, 100.0, hist)
            else:
                ops.admm_track(st, sse, 100.0, g, bstar, best_g, best_b, hist, wcodes, best_w)
            snaps.append((g.clone(), bstar.clone(), wcodes.clone()))
        if fused:
            ops.admm_keep(st, g, bstar, best_g, best_b, wcodes, best_w)
        return best_g, best_b, best_w, hist, st.read(), snaps
    bg1, bb1, bw1, h1, s1, snaps = run(True)
    bg0, bb0, bw0, h0, s0, _ = run(False)
    assert s1["best_iter"] == s0["best_iter"] == 4 and s1["iter"] == s0["iter"] == len(losses)
    assert s1["best_conv_scale"] == s0["best_conv_scale"] and s1["best_loss"] == s0["best_loss"]
    assert torch.equal(h1, h0)
    assert torch.equal(bg1, bg0) and torch.equal(bb1, bb0) and torch.equal(bw1, bw0)
    assert torch.equal(bg1, snaps[4][0]) and torch.equal(bb1, snaps[4][1]) and torch.equal(bw1, snaps[4][2])


@pytest.mark.parametrize("n,c1,c2,sp", [(2, 32, 32, (6, 16, 8)), (1, 64, 64, (4, 8, 16)), (1, 16, 24, (5, 10, 12))])
def test_quadform_delta_equals_conv_sse(ops, n, c1, c2, sp):
    """Residual-form conv-free scoring of a QUANTISED layer: statistics of R = Y - conv(first iterate) from the
    tcgen05 Gram kernel (integer codes, att = None) + the tiled fp64 quadratic form must reproduce the squared
    error of later iterates (fp64 conv on the CPU) to ~1e-6, and the fused bookkeeping must equal admm_decide's."""
    torch.manual_seed(c1 + c2)
    la = 16
    codes = torch.randint(0, la, (n, c1, *sp)).float()
    sc = float(np.float32(0.173))
    x = codes.double() * sc                        # the kernel's model of the activations: fp32 scale x integer code
    w_true = torch.randn(c2, c1, 3, 3, 3) * (2.0 / (27 * c1)) ** 0.5
    b_true = torch.randn(c2) * 0.05
    y = F.conv3d(x, w_true.double(), b_true.double(), 1, 1).float()
    g_ref = w_true + 0.02 * w_true.std() * torch.randn_like(w_true)      # "first iterate"
    b_ref = b_true + 0.01 * torch.randn_like(b_true)
    r = (y.double() - F.conv3d(x, g_ref.double(), b_ref.double(), 1, 1)).float()
    xq = codes.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16).to(DEV)
    acc, _, flag = ops.gram_tc_f64(xq, torch.tensor([sc], device=DEV), r.to(DEV), None, True, att_exact=True)
    assert int(flag.item()) == 0
    # S against the fp64 im2col Gram: the integer part is exact
    cols = O.im2col(codes, 3, 3, 3, 1, 1).double() * sc
    cols = torch.cat([cols, torch.ones(1, cols.shape[1], dtype=torch.float64)], 0)
    kp = cols.shape[0]
    s_ref = cols @ cols.T
    assert (acc[:kp].cpu() - s_ref).abs().max().item() <= 1e-12 * s_ref.abs().max().item()
    t_ref = r.permute(1, 0, 2, 3, 4).reshape(c2, -1).double() @ cols.T
    assert (acc[kp:].cpu() - t_ref).abs().max().item() <= 2e-5 * t_ref.abs().max().item()
    yy = torch.tensor([(r.double() ** 2).sum().item()], dtype=torch.float64, device=DEV)
    sse = torch.zeros(1, dtype=torch.float64, device=DEV)
    dev = torch.device(DEV)
    st, st2 = ops.AdmmState(dev), ops.AdmmState(dev)
    numel = float(y.numel())
    hist, hist2 = torch.zeros(3, device=DEV), torch.zeros(3, device=DEV)
    gd, bd = g_ref.reshape(c2, -1).contiguous().to(DEV), b_ref.to(DEV)
    for i, eps in enumerate((2e-2, 5e-3, 1e-3)):
        g = g_ref + eps * w_true.std() * torch.randn_like(g_ref)
        b = b_ref + eps * 0.1 * torch.randn_like(b_ref)
        # what the statistics describe: sum((g - g_ref) X + (b - b_ref) - R)^2 with the fp32-rounded residual R
        ref = ((F.conv3d(x, (g - g_ref).double(), (b - b_ref).double(), 1, 1) - r.double()) ** 2).sum().item()
        ops.quadform_delta(acc, yy, g.reshape(c2, -1).contiguous().to(DEV), b.to(DEV), sse, gd, bd, st=st,
                           numel=numel, history=hist)
        got = sse.item()
        print(f"c1={c1} eps={eps}: sse {got:.9e} fp64 conv {ref:.9e} rel {abs(got - ref) / ref:.1e}")
        assert abs(got - ref) <= 3e-6 * ref, (got, ref)
        ops.admm_decide(st2, sse, numel, hist2)
        assert st.read() == st2.read()
    assert torch.equal(hist, hist2)
    # plain form (g_ref = None) on the statistics of y itself == the old single-CTA kernel
    acc_y, _, _ = ops.gram_tc_f64(xq, torch.tensor([sc], device=DEV), y.to(DEV), None, True, att_exact=True)
    yy_y = torch.tensor([(y.double() ** 2).sum().item()], dtype=torch.float64, device=DEV)
    g = g_ref.reshape(c2, -1).contiguous().to(DEV)
    ops.quadform_delta(acc_y, yy_y, g, bd, sse)
    ws = ops.workspace(16 + 8 * 1024, dev)
    sse_old = torch.zeros(1, dtype=torch.float64, device=DEV)
    ops.quadform_sse(acc_y, float(yy_y.item()), g, bd, sse_old, ws)
    assert abs(sse.item() - sse_old.item()) <= 1e-9 * abs(sse_old.item()) + 1e-9 * float(yy_y.item())


@pytest.mark.parametrize("n,c1,c2,sp,int_att", [(2, 32, 32, (6, 16, 8), True), (1, 64, 64, (4, 8, 16), True),
                                                 (1, 16, 24, (5, 10, 12), False), (2, 32, 16, (3, 9, 7), True)])
def test_gram_tc_dual_and_rows_only(ops, n, c1, c2, sp, int_att):
    """One Gram pass, two accumulators: the attention-weighted normal equations (bit-identical to effq_gram_tc) and the
    UNWEIGHTED S = X^ X^T with its bias row / column (exact: integer codes); then the rows-only pass for a new target
    gives T = R X^T -- together the statistics of effq_quadform_delta."""
    torch.manual_seed(c1 * 3 + c2)
    la = 16
    codes = torch.randint(0, la, (n, c1, *sp)).float()
    sc = float(np.float32(0.211))
    y = torch.randn(n, c2, *sp)
    att = (torch.rand(n, *sp) * 3).floor() + 1 if int_att else torch.rand(n, *sp) * 2 + 0.25
    xq = codes.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16).to(DEV)
    cs = torch.tensor([sc], device=DEV)
    exact = ops.att_is_exact(att.to(DEV), la - 1)
    a0, b0, stats, ws, flag = ops.gram_tc_dual(xq, cs, y.to(DEV), att.to(DEV), att_exact=exact)
    assert int(flag.item()) == 0
    a1, b1, _, flag1 = ops.gram_tc(xq, cs, y.to(DEV), att.to(DEV), True, att_exact=exact)
    assert torch.equal(a0, a1) and torch.equal(b0, b1)
    cols = O.im2col(codes, 3, 3, 3, 1, 1).double() * sc
    cols = torch.cat([cols, torch.ones(1, cols.shape[1], dtype=torch.float64)], 0)
    kp = cols.shape[0]
    s_ref = cols @ cols.T
    got = stats[:kp].cpu()
    assert (got - s_ref).abs().max().item() <= 1e-12 * s_ref.abs().max().item()
    r = torch.randn(n, c2, *sp) * 0.3
    _, flag2 = ops.gram_tc_rows_f64(xq, cs, r.to(DEV), stats, ws=ws)
    assert int(flag2.item()) == 0
    t_ref = r.permute(1, 0, 2, 3, 4).reshape(c2, -1).double() @ cols.T
    assert (stats[kp:].cpu() - t_ref).abs().max().item() <= 2e-5 * t_ref.abs().max().item()
    assert torch.equal(stats[:kp].cpu(), got)                       # the rows-only pass leaves S alone


# ---------------------------------------------------------------- proximal-step GEMM (a9)
def test_split3_bf16_is_exact(ops):
    torch.manual_seed(21)
    x = (torch.randn(37, 131) * torch.logspace(-6, 4, 131)).to(DEV)
    pl = ops.split3_bf16(x)
    assert pl.shape == (3, 37, 192) and pl.dtype == torch.bfloat16
    back = pl[0].float() + pl[1].float() + pl[2].float()            # exact in fp32: the terms do not overlap
    assert torch.equal(back[:, :131], x)
    assert torch.count_nonzero(pl[:, :, 131:]) == 0


@pytest.mark.parametrize("m,k", [(32, 865), (64, 1729), (16, 433), (128, 3457), (256, 1000), (48, 300)])
def test_solve_gemm_tc_fp32_class_accuracy(ops, m, k):
    """w* = B A^-1 from bf16 split planes: the error against fp64 must be in the class of an fp32 GEMM
    (the library SGEMM it replaces), on an SPD inverse with the dynamic range of the real systems."""
    torch.manual_seed(m + k)
    x = torch.randn(k, 2 * k, device=DEV, dtype=torch.float64)
    a = x @ x.T / (2 * k) + 0.05 * torch.eye(k, device=DEV, dtype=torch.float64)
    ainv = torch.linalg.inv(a).float()
    ainv = 0.5 * (ainv + ainv.T)
    b = (torch.randn(m, k, device=DEV) * torch.logspace(-2, 2, k, device=DEV)).float()
    ref = b.double() @ ainv.double()
    got, _ = ops.solve_gemm_tc(ops.split3_bf16(b), ops.split3_bf16(ainv), k)
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = b @ ainv
    scale = ref.abs().max().item()
    err_tc = (got.double() - ref).abs().max().item() / scale
    err_lib = (lib.double() - ref).abs().max().item() / scale
    print(f"m={m} k={k}: tensor-core split err {err_tc:.2e}, fp32 SGEMM err {err_lib:.2e}")
    assert got.shape == (m, k)
    # (the TMEM accumulator truncates instead of rounding to nearest: a small constant factor over SGEMM)
    assert err_tc <= max(4.0 * err_lib, 4e-6)
    # fused producer of the A planes: admm_rhs emits the same three terms as split3 of its fp32 output
    kk = k - 1
    b0, w0p = b, torch.randn(m, k, device=DEV)
    g, dual = torch.randn(m, kk, device=DEV), torch.randn(m, kk, device=DEV)
    out32 = torch.empty(m, k, device=DEV)
    planes = torch.empty((3, m, ops.split3_ld(k)), dtype=torch.bfloat16, device=DEV)
    ops.admm_rhs(b0, w0p, g, dual, 10.0, 1.0, out32, planes=planes)
    assert torch.equal(planes, ops.split3_bf16(out32))


# ---------------------------------------------------------------- own SPD factorisation + inverse (a9)
@pytest.mark.parametrize("n", [33, 109, 129, 257, 865, 1729, 3457])
def test_spd_inverse_own_kernels(ops, n):
    """Blocked Cholesky + block triangular inverse + W^T W through the repo's kernels (spd_inverse.py) against an
    fp64 inverse, on a matrix with the structure of the real normal matrices (Gram of non-negative features + a small
    ridge: cond ~ 1e4..1e5); error in the class of the library's fp32 Cholesky inverse.  Also: replayed launch
    sequence == first run, and a non-positive pivot is reported like LAPACK's info."""
    from efficientq_b200.spd_inverse import SpdInverter
    torch.manual_seed(n)
    v = 4 * n
    x = torch.relu(torch.randn(n, v, device=DEV, dtype=torch.float64)) * 0.4
    a64 = 2.0 * x @ x.T + (0.02 * v) * torch.eye(n, device=DEV, dtype=torch.float64)
    a = a64.float()
    ref = torch.linalg.inv(a.double())
    inv = SpdInverter(torch.device(DEV))
    got, info = inv.invert(a)
    assert int(info.item()) == 0
    lib = torch.cholesky_inverse(torch.linalg.cholesky(a))
    scale = ref.abs().max().item()
    e_own = (got.double() - ref).abs().max().item() / scale
    e_lib = (lib.double() - ref).abs().max().item() / scale
    resid = (got.double() @ a.double() - torch.eye(n, device=DEV, dtype=torch.float64)).abs().max().item()
    line = f"n={n}: A^-1 vs fp64: own kernels {e_own:.2e}, library fp32 Cholesky inverse {e_lib:.2e} | max|A^-1 A - I| {resid:.2e}"
    print(line)
    import os
    if os.path.isdir("gpurun_out"):
        open("gpurun_out/r02_spd_inverse.txt", "a").write(line + "\n")
    assert e_own <= max(4.0 * e_lib, 2e-6)
    got2, _ = inv.invert(a)                                 # second call replays the recorded launch sequence
    assert torch.equal(got, got2)
    bad = a.clone()
    bad[n // 2, n // 2] = -1.0
    _, info_bad = inv.invert(bad)
    assert int(info_bad.item()) == n // 2 + 1
