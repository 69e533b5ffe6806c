"""The drop-in boundary, pinned to the reference itself (tests/golden/interface.json, written by
tests/golden/make_golden.py::gen_interface from the reference's argparse parser and from the nets its own
``definer`` builds out of config/*_ptq.yaml): every command-line flag with type / default / action, and the
state-dict layout (keys, shapes, order) plus the per-layer quantizer settings of the BraTS- and LiTS-config nets --
so a reference checkpoint loads with ``strict=True`` semantics and a reference command line parses unchanged."""
import argparse
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = json.load(open(os.path.join(ROOT, "tests", "golden", "interface.json")))
EXTENSIONS = {"tune_act_iter", "w_per_channel"}                     # flags this package adds


def test_command_line_equals_reference_parser():
    from efficientq_b200 import entrance
    ours = {}
    for act in entrance.build_parser()._actions:
        if act.dest == "help":
            continue
        if not act.option_strings:
            ours[act.dest] = {"positional": True, "choices": list(act.choices) if act.choices else None}
        else:
            ours[act.dest] = {"option": act.option_strings[0], "type": act.type.__name__ if act.type else None,
                              "default": act.default, "flag": isinstance(act, argparse._StoreTrueAction),
                              "choices": list(act.choices) if act.choices else None}
    assert set(ours) - EXTENSIONS == set(REF["flags"])
    for dest, spec in REF["flags"].items():
        assert ours[dest] == spec, (dest, ours[dest], spec)


@pytest.mark.parametrize("task,lw,la", [("brats", 16, 16), ("lits", 4, 4)])
def test_net_layout_equals_reference(task, lw, la):
    from efficientq_b200 import definer, entrance
    from efficientq_b200.qconv import PTQConv
    ref = REF["nets"][task]
    a = entrance.build_parser().parse_args(["ptq", "--qlvl_w", str(lw), "--qlvl_a", str(la), "--config",
                                            os.path.join(ROOT, "config", f"{task}_ptq.yaml")])
    a = entrance.merge_config(a.config, a)
    QConv, qinfo, kwQ = definer.get_conv_class(a)
    cube, info = definer.get_model_cube(a, QConv, kwQ)
    assert (qinfo, info, cube["num_mo"], cube["nClass"], cube["nMod"], sorted(kwQ)) == \
        (ref["qinfo"], ref["model_info"], ref["num_mo"], ref["nClass"], ref["nMod"], ref["kwQ"])
    state = [[k, list(v.shape)] for k, v in cube["model"].state_dict().items()]
    assert state == ref["state"]                                     # same keys, same shapes, same order
    qmods = [[n, m.in_channels, m.out_channels, list(m.kernel_size), list(m.stride), list(m.padding),
              int(m.qlvl_w), int(m.qlvl_act), bool(m.q_act)]
             for n, m in cube["model"].named_modules() if isinstance(m, PTQConv)]
    assert qmods == ref["quantizers"]
