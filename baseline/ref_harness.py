"""Runs the UNMODIFIED reference (rongzhao-zhang/EfficientQ) on the CPU: the honest CPU arm of bench.py.

The reference is pure Python.  ``ensure_ref()`` copies its ``src/`` and ``config/`` from the read-only checkout
(/root/reference, present in the build container only) into the git-ignored ``baseline/_ref/`` -- which travels to
the GPU box with the working tree -- and ``import_reference()`` imports its modules from there with four inert stub
modules (pytz, matplotlib.pyplot, nibabel, progressbar: none of them touches the arithmetic).  Nothing of the
reference is modified; ``lwq_iter`` (the ADMM iteration count) is an attribute of its module
(src/models/EfficientQConv.py:23) and is set from outside for the bounded sample.

``timed_ptq`` runs the reference's own do_ptq core (src/ptqer.py:313-364: FP forward with hooks, attention-mask
pyramid, quantizing forward -> EfficientQConv.ptq per layer) and attributes the wall-clock to phases by wrapping the
functions it calls (timing only).
"""
from __future__ import annotations

import os
import shutil
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference"


def ensure_ref() -> bool:
    """Populate baseline/_ref from the reference checkout when that exists; True if baseline/_ref is usable."""
    if os.path.isdir(os.path.join(REF_SRC, "src")):
        for sub in ("src", "config"):
            dst = os.path.join(REF_DIR, sub)
            if not os.path.isdir(dst):
                shutil.copytree(os.path.join(REF_SRC, sub), dst)
    return os.path.isfile(os.path.join(REF_DIR, "src", "models", "EfficientQConv.py"))


def import_reference():
    """The reference's modules, imported from baseline/_ref/src.  Raises if ensure_ref() never ran."""
    src = os.path.join(REF_DIR, "src")
    if not os.path.isdir(src):
        raise RuntimeError("baseline/_ref/src missing: run __graft_entry__.build() where /root/reference exists")
    if src not in sys.path:
        sys.path.insert(0, src)
    for name in ["matplotlib", "matplotlib.pyplot", "pytz", "nibabel", "progressbar"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import importlib
    import models  # noqa: F401
    mods = {n: sys.modules.get(f"models.{n}") or importlib.import_module(f"models.{n}")
            for n in ("layer_helper", "solver", "PTQConv", "EfficientQConv", "model_blk", "fold_bn", "hooks")}
    import definer
    import ptqer
    return dict(lh=mods["layer_helper"], solver=mods["solver"], ptqconv=mods["PTQConv"], effq=mods["EfficientQConv"],
                model_blk=mods["model_blk"], fold_bn=mods["fold_bn"], hooks=mods["hooks"], ptqer=ptqer, definer=definer)


class PhaseTimers:
    """Wall-clock of the reference's phases, by wrapping the functions EfficientQConv.ptq calls:
    act_search / w_project (project_by_iter on activations / weights, layer_helper.py:40-70), im2col_gram
    (QuadraSolver.__init__: im2col_loop + getA0B0, solver.py:202-314), solve (solver.py:327-345), conv_mse
    (F.conv3d + F.mse_loss, EfficientQConv.py:118-122)."""

    def __init__(self, R):
        self.R = R
        self.t = {}

    def _wrap(self, key, fn, pick=None):
        def timed(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                name = pick(*a, **k) if pick else key
                self.t[name] = self.t.get(name, 0.0) + time.perf_counter() - t0
        return timed

    def __enter__(self):
        effq, solver = self.R["effq"], self.R["solver"]
        self._saved = (effq.project_by_iter, solver.QuadraSolver.__init__, solver.QuadraSolver.solve, effq.F.conv3d,
                       effq.F.mse_loss)
        effq.project_by_iter = self._wrap(None, self._saved[0], pick=lambda v, n, lo, hi: "act_search" if lo == 0 else "w_project")
        solver.QuadraSolver.__init__ = self._wrap("im2col_gram", self._saved[1])
        solver.QuadraSolver.solve = self._wrap("solve", self._saved[2])
        effq.F.conv3d = self._wrap("conv_mse", self._saved[3])
        effq.F.mse_loss = self._wrap("conv_mse", self._saved[4])
        return self

    def __exit__(self, *exc):
        effq, solver = self.R["effq"], self.R["solver"]
        (effq.project_by_iter, solver.QuadraSolver.__init__, solver.QuadraSolver.solve, effq.F.conv3d,
         effq.F.mse_loss) = self._saved
        return False


def build_reference_model(R, args, state_fn):
    """The reference's own network for ``args`` (its definer + its EfficientQConv), seeded like the GPU arm's."""
    QConv, _, kwQ = R["definer"].get_conv_class(args)
    cube, _ = R["definer"].get_model_cube(args, QConv, kwQ)
    model = cube["model"]
    model.load_state_dict(state_fn(model), strict=False)
    model.eval()
    R["fold_bn"].search_fold_and_remove_bn(model)
    return model


def timed_ptq(R, model, data, task: str, init_stride, n_iter: int):
    """do_ptq core of the reference on ``data`` (CPU) with ``n_iter`` ADMM iterations per layer.  Returns
    (phase seconds, t_fp, t_ptq, layer_loss lines)."""
    import torch
    ptqer = R["ptqer"]
    ptqer.set_name(model)
    ptqer.set_fp(model)
    # the reference's hook stores o.detach().cpu(): a copy on its GPU run mode, an alias on CPU tensors that the next
    # in-place ReLU would overwrite -- register the copying form (same semantics as the reference on a device)
    handles = [m.register_forward_hook(lambda mod, i, o: setattr(mod, "output_fp", o.detach().clone()))
               for m in model.modules() if isinstance(m, R["ptqconv"].PTQConv)]
    t0 = time.perf_counter()
    with torch.no_grad():
        output_fp = model(data).detach()
    body = (data[:, 0] != 0.0).bool() if task == "brats" else torch.ones_like(data[:, 0]).bool()
    wmap, _ = ptqer.get_att_weight_map(output_fp, torch.ones_like(data[:, 0]).bool(), "p:0.5", task=task)
    pyr = ptqer.get_mask_pyramid(output_fp, body, wmap, init_stride, num_lvls=5, task=task)
    ptqer.set_mask(model, pyr)
    for h in handles:
        h.remove()
    layer_loss = []
    ptqer.set_anything(model, "layer_loss", layer_loss)
    for m in model.modules():
        if isinstance(m, R["ptqconv"].PTQConv):
            m.lwq_iter = n_iter
    t1 = time.perf_counter()
    ptqer.set_quantizing(model)
    timers = PhaseTimers(R)
    with timers, torch.no_grad():
        model(data)
    t2 = time.perf_counter()
    ptqer.set_quantized(model)
    return timers.t, t1 - t0, t2 - t1, layer_loss
