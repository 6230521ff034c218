"""BertAdam oracle (oracle/vb_oracle.py::bert_adam_step) against tests/golden/bert_adam.npz, which was produced by the
reference's own `BertAdam` class (optimization.py:185-304) in oracle/make_golden.py."""
import os

import numpy as np
import torch

import golden_util  # noqa: F401  (puts oracle/ on sys.path)
import adam_util
import vb_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bert_adam.npz")


def test_bert_adam_oracle_matches_reference_golden():
    gold = np.load(GOLD)
    params, grads = adam_util.scenario()
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    steps = [0] * len(params)
    clipped = 0
    for s in range(adam_util.STEPS):
        for i in range(len(params)):
            params[i], m[i], v[i], steps[i], g = vb_oracle.bert_adam_step(
                params[i], grads[s][i], m[i], v[i], steps[i], schedule="warmup_linear",
                weight_decay=adam_util.WEIGHT_DECAY[i], **adam_util.HYPER)
            clipped += int(not torch.equal(g, grads[s][i]))
            for name, t in (("p", params[i]), ("m", m[i]), ("v", v[i])):
                ref = gold[f"{name}{i}_s{s}"]
                np.testing.assert_allclose(t.numpy(), ref, rtol=1e-5, atol=1e-8, err_msg=f"{name}{i} step {s}")
    assert 0 < clipped < adam_util.STEPS * len(params)  # the clip engaged on some tensors and not on others


def test_lr_schedules_match_reference_formulas():
    # warmup_linear (optimization.py:165-174): ramp to 1 at progress == warmup, linear decay to 0 at progress == 1
    f = vb_oracle.lr_schedule
    assert f("warmup_linear", 0, 0.1, 100) == 0.0
    assert abs(f("warmup_linear", 5, 0.1, 100) - 0.5) < 1e-12
    assert abs(f("warmup_linear", 10, 0.1, 100) - 1.0) < 1e-12
    assert abs(f("warmup_linear", 55, 0.1, 100) - 0.5) < 1e-12
    assert f("warmup_linear", 150, 0.1, 100) == 0.0
    assert f("warmup_linear", 7, 0.1, -1) == 1.0           # t_total < 0: schedule not applied (optimization.py:61-62)
    assert f("warmup_constant", 50, 0.1, 100) == 1.0 and abs(f("warmup_constant", 5, 0.1, 100) - 0.5) < 1e-12
    assert abs(f("warmup_cosine", 55, 0.1, 100) - 0.5) < 1e-12
    assert f(None, 3, 0.1, 100) == 1.0


def test_product_schedules_equal_oracle_schedules():
    """visualbert_b200.optimization's schedule classes (host-side logic, no GPU) against the oracle formulas."""
    from visualbert_b200 import optimization as O
    for name in (None, "none", "warmup_linear", "warmup_constant", "warmup_cosine"):
        for warmup, t_total in ((0.1, 100), (0.25, 8), (-1, -1), (0.0, 10)):
            if warmup == 0.0 and name in ("warmup_linear", "warmup_constant", "warmup_cosine"):
                continue  # progress < 0 never holds: same as the reference, no division by the zero warmup
            sched = O.SCHEDULES[name](warmup=warmup, t_total=t_total)
            for step in (0, 1, 2, 5, 8, 10, 50, 99, 100, 150):
                assert abs(sched.get_lr(step) - vb_oracle.lr_schedule(name, step, warmup, t_total)) < 1e-12, (name, warmup, step)


def test_bert_adam_constructor_validation_matches_reference():
    import pytest
    from visualbert_b200 import BertAdam
    p = [torch.nn.Parameter(torch.zeros(2))]
    for bad in (dict(lr=-1.0), dict(lr=1e-3, schedule="nope"), dict(lr=1e-3, b1=1.0), dict(lr=1e-3, b2=-0.1), dict(lr=1e-3, e=-1.0)):
        with pytest.raises(ValueError):
            BertAdam(p, **bad)
    opt = BertAdam(p, lr=1e-3, warmup=0.1, t_total=10)
    assert opt.get_lr() == [0]  # no state yet (optimization.py:233-234)
