"""Host logic of the score-stream overlap (efficientq_b200/layer_engine.py `_split_score`): which recorded tail of an
ADMM iteration is split into (up to the projection | projection | scoring launches on the score stream).  The order of
every dependent pair must survive: search -> project -> score, and nothing but the scoring launches changes stream
(reference loop: src/models/EfficientQConv.py:99-144)."""
import ctypes as C
import types


def _engine():
    from efficientq_b200 import layer_engine
    eng = layer_engine.LayerCalibrator.__new__(layer_engine.LayerCalibrator)
    eng._score_stream = types.SimpleNamespace(cuda_stream=0xBEEF)
    eng._split_cache = None
    return eng


def _call(name, stream=0x1, tag=None):
    return (lambda *a: 0, (1, 2, C.c_void_p(stream)), tag, name)


def test_conv_free_iteration_is_split():
    eng = _engine()
    post = [_call("effq_solve_gemm_tc"), _call("effq_scale_search"), _call("effq_admm_project"),
            _call("effq_quadform_delta", tag=("quadform_k865", {}))]
    head, proj, score = eng._split_score(post)
    assert [c[3] for c in head] == ["effq_solve_gemm_tc", "effq_scale_search"]
    assert [c[3] for c in proj] == ["effq_admm_project"]
    assert [c[3] for c in score] == ["effq_quadform_delta"]
    assert score[0][1][-1].value == 0xBEEF and score[0][1][:-1] == (1, 2)       # only the stream argument changes
    assert proj[0][1][-1].value == 0x1 and post[-1][1][-1].value == 0x1         # the recorded sequence is untouched
    assert eng._split_score(post) is not None and eng._split_cache[0] is post   # cached per recorded sequence


def test_conv_scored_iteration_moves_conv_and_decide(monkeypatch):
    eng = _engine()
    post = [_call("effq_solve_gemm_tc"), _call("effq_solve_gemm_tc"), _call("effq_scale_search"),
            _call("effq_admm_project"), _call("effq_conv3d_tc"), _call("effq_admm_decide")]
    head, proj, score = eng._split_score(post)
    assert len(head) == 3 and [c[3] for c in score] == ["effq_conv3d_tc", "effq_admm_decide"]
    assert all(c[1][-1].value == 0xBEEF for c in score)
    monkeypatch.setenv("EFFQ_SCORE_STREAM_CONV", "0")
    eng._split_cache = None
    assert eng._split_score(post) is None


def test_other_tails_stay_on_one_stream():
    eng = _engine()
    # projection not directly in front of the scoring launch (e.g. a per-channel search in between): no split
    assert eng._split_score([_call("effq_admm_project"), _call("effq_scale_search_rows"), _call("effq_quadform_delta")]) is None
    eng._split_cache = None
    assert eng._split_score([_call("effq_scale_search"), _call("effq_admm_project"), _call("effq_conv3d_f32"),
                             _call("effq_admm_decide")]) is None
    eng._split_cache = None
    assert eng._split_score([]) is None


def test_timer_keeps_bracketed_scoring_on_the_main_stream():
    from efficientq_b200 import ops
    eng = _engine()
    post = [_call("effq_scale_search"), _call("effq_admm_project"), _call("effq_quadform_delta", tag=("quadform_k865", {}))]
    ops.timer.enabled, ops.timer.only = True, None
    try:
        assert eng._split_score(post) is None              # CUDA events of the timer are recorded on the main stream
        ops.timer.only = ("gram_tc_dual",)
        assert eng._split_score(post) is not None          # the timed region of bench.py brackets one other kernel only
    finally:
        ops.timer.enabled, ops.timer.only = False, None
