"""The N>1 path on CPU: two gloo ranks shard the calibration volumes and all-reduce exactly the
statistics efficientq_b200/dist.py lists; the result must equal the unsharded oracle.  (The
kernels themselves are CUDA-only; here the oracle stands in for the per-rank compute so that
the sharding + reduction logic is what is under test.)"""
import os

import torch
import torch.distributed as td
import torch.multiprocessing as mp

from oracle import effq_oracle as O


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from efficientq_b200.dist import DistCtx, shard_range
    ctx = DistCtx()
    g = torch.Generator().manual_seed(21)
    n, c1, c2, sp = 4, 6, 5, (6, 6, 6)
    x = torch.relu(torch.randn(n, c1, *sp, generator=g))
    w = torch.randn(c2, c1, 3, 3, 3, generator=g) * 0.1
    b = torch.randn(c2, generator=g) * 0.1
    y = torch.nn.functional.conv3d(x, w, b, 1, 1)
    att = (torch.rand(n, *sp, generator=g) * 3).floor() + 1
    lo, hi = ctx.shard(n)
    assert (lo, hi) == shard_range(n, rank, world)
    xs, ys, atts = x[lo:hi], y[lo:hi], att[lo:hi]

    # activation scale search: 2 doubles per pass
    v = xs.double()
    s = ctx.all_reduce_sum(torch.stack([v.abs().sum(), torch.tensor(float(v.numel()), dtype=torch.float64)]))
    a, a_prev, passes = (s[0] / s[1]).item(), -999.0, 0
    while abs(a - a_prev) > 1e-5:
        bq = O.discretize(v / a, 16, 0, 1)
        s = ctx.all_reduce_sum(torch.stack([(bq * v).sum(), (bq * bq).sum()]))
        a_prev, a = a, (s[0] / s[1]).item()
        passes += 1
    a_ref, _, p_ref = O.project_by_iter(x, 16, 0, 1, return_iters=True)
    assert abs(a - a_ref) < 1e-12 and passes == p_ref

    # normal-equation partial sums
    qx = a * O.discretize(v / a, 16, 0, 1).float()
    ne = O.NormalEquations(qx, ys, (3, 3, 3), 1, 1, w, b, atts)
    a0 = ctx.all_reduce_sum(ne.a0.clone())
    b0 = ctx.all_reduce_sum(ne.b0.clone())
    qx_full = a_ref * O.discretize(x.double() / a_ref, 16, 0, 1).float()
    full = O.NormalEquations(qx_full, y, (3, 3, 3), 1, 1, w, b, att)
    assert torch.allclose(a0, full.a0, rtol=1e-5, atol=1e-4) and torch.allclose(b0, full.b0, rtol=1e-5, atol=1e-4)

    # rho_scale moments and the per-iteration squared error
    m = ctx.all_reduce_sum(torch.stack([ys.double().sum(), (ys.double() ** 2).sum(),
                                        torch.tensor(float(ys.numel()), dtype=torch.float64)]))
    std = ((m[1] - m[0] ** 2 / m[2]) / (m[2] - 1)).sqrt().item()
    assert abs(std - y.std().item()) < 1e-5
    out_q = torch.nn.functional.conv3d(qx, w, b, 1, 1)
    sse = ctx.all_reduce_sum(((out_q - ys).double() ** 2).sum().reshape(1))
    ref = torch.nn.functional.mse_loss(torch.nn.functional.conv3d(qx_full, w, b, 1, 1), y).item()
    assert abs(sse.item() / y.numel() - ref) < 1e-6 * max(ref, 1e-12) + 1e-9
    if rank == 0:
        out.put("ok")
    td.destroy_process_group()


def test_two_rank_sharded_statistics_equal_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


def test_shard_range_partitions():
    from efficientq_b200.dist import shard_range
    for n in (1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class _EvalCube:
    """Five deterministic volumes; hands out a rank's round-robin share like CalibrationData.evaluation_volumes."""

    def __init__(self):
        g = torch.Generator().manual_seed(5)
        self.vols = [(f"sn{i}", torch.randn(2, 20, 16, 16, generator=g), torch.randint(0, 4, (20, 16, 16), generator=g))
                     for i in range(5)]

    def evaluation_volumes(self, split, rank=0, world=1):
        return iter(self.vols[rank::world]) if split == "val" else None


class _EvalNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(9)
        self.a = torch.nn.Conv3d(2, 3, 3, padding=1)
        self.b = torch.nn.Conv3d(2, 3, 1)

    def forward(self, x):
        return torch.stack([self.b(x), self.a(x)])


def _eval_worker(rank, world, port, out, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from efficientq_b200.dist import DistCtx
    from efficientq_b200.evaluate import PTQTester
    args = dict(num_mo=2, n_class=3, patch_size=(16, 16, 16), overlap=8, multi_label="brats", multilabel_fusetype="con")
    root = os.path.join(tmp, "sharded") if rank == 0 else None
    res = PTQTester(_EvalNet(), _EvalCube(), root, "cpu", dist=DistCtx(), **args).test_as_is("ptq")
    if rank == 0:
        ref = PTQTester(_EvalNet(), _EvalCube(), os.path.join(tmp, "single"), "cpu", **args).test_as_is("ptq")
        a = open(os.path.join(tmp, "sharded", "ptq", "val_seg.txt")).read()
        b = open(os.path.join(tmp, "single", "ptq", "val_seg.txt")).read()
        assert a == b and res == ref, (a, b)              # same report, character for character
        out.put("ok")
    else:
        out.put(res["val"]["dsc"])                        # every rank ends up with the merged metrics
    td.destroy_process_group()


def test_two_rank_sharded_evaluation_equals_unsharded(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = [q.get(timeout=5), q.get(timeout=5)]
    assert "ok" in got and any(isinstance(v, float) and 0.0 <= v <= 1.0 for v in got)
