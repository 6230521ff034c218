"""visualbert_b200.BertAdam (vb_bert_adam_step through the C ABI) against the reference-generated golden
(tests/golden/bert_adam.npz) and against the oracle on larger / unaligned / multi-chunk tensors."""
import os

import numpy as np
import pytest
import torch

import golden_util  # noqa: F401  (puts oracle/ on sys.path)
import adam_util
import vb_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "bert_adam.npz")


def test_bert_adam_matches_reference_golden():
    from visualbert_b200 import BertAdam
    dev = torch.device("cuda:0")
    gold = np.load(GOLD)
    init, grads = adam_util.scenario()
    params = [torch.nn.Parameter(t.clone().to(dev)) for t in init]
    opt = BertAdam([{"params": params[:3], "weight_decay": 0.01}, {"params": params[3:], "weight_decay": 0.0}], **adam_util.HYPER)
    for s in range(adam_util.STEPS):
        for i, p in enumerate(params):
            p.grad = grads[s][i].clone().to(dev)
        versions = [p._version for p in params]
        opt.step()
        assert all(p._version > v for p, v in zip(params, versions))  # bf16 weight caches key on the version
        for i, p in enumerate(params):
            assert opt.state[p]["step"] == s + 1
            for name, t in (("p", p.detach()), ("m", opt.state[p]["next_m"]), ("v", opt.state[p]["next_v"])):
                np.testing.assert_allclose(t.cpu().numpy(), gold[f"{name}{i}_s{s}"], rtol=1e-5, atol=1e-8,
                                           err_msg=f"{name}{i} step {s}")
    assert abs(opt.get_lr()[0] - adam_util.HYPER["lr"] * vb_oracle.lr_schedule("warmup_linear", 4, 0.25, 8)) < 1e-12


@pytest.mark.parametrize("max_grad_norm", [1.0, -1.0])
def test_bert_adam_large_unaligned_tensors_vs_oracle(max_grad_norm):
    """Multi-chunk tensors (> VB_ADAM_CHUNK), views at odd offsets of a flat buffer (the scalar path), a parameter
    without gradient (skipped, its step counter does not advance), clipping on and off."""
    from visualbert_b200 import BertAdam
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    numels = [100003, 32768, 1, 7, 70001, 5]
    flat_p = torch.randn(sum(numels) + 3, device=dev) * 0.1
    flat_g = torch.randn(sum(numels) + 3, device=dev)
    guard_lo, guard_hi = flat_p[:1].clone(), flat_p[-2:].clone()
    params, off = [], 1  # offset 1: every view is misaligned for 16-byte access
    for n in numels:
        p = torch.nn.Parameter(flat_p[off: off + n])
        p.grad = flat_g[off: off + n] * (0.001 if n == 70001 else 1.0)
        params.append(p)
        off += n
    frozen = torch.nn.Parameter(torch.randn(10, device=dev))  # grad None
    opt = BertAdam(params + [frozen], lr=2e-3, warmup=0.1, t_total=100, weight_decay=0.01, max_grad_norm=max_grad_norm)
    ref = [(p.detach().clone(), torch.zeros_like(p), torch.zeros_like(p), 0) for p in params]
    for _ in range(3):
        g_now = [p.grad.clone() for p in params]
        opt.step()
        for i in range(len(params)):
            pr, m, v, st = ref[i]
            pr, m, v, st, _ = vb_oracle.bert_adam_step(pr, g_now[i], m, v, st, lr=2e-3, schedule="warmup_linear", warmup=0.1,
                                                       t_total=100, weight_decay=0.01, max_grad_norm=max_grad_norm)
            ref[i] = (pr, m, v, st)
    for i, p in enumerate(params):
        pr, m, v, st = ref[i]
        assert opt.state[p]["step"] == st == 3
        torch.testing.assert_close(p.detach(), pr, rtol=2e-5, atol=1e-7)
        torch.testing.assert_close(opt.state[p]["next_m"], m, rtol=2e-5, atol=1e-8)
        torch.testing.assert_close(opt.state[p]["next_v"], v, rtol=2e-5, atol=1e-10)
    assert len(opt.state[frozen]) == 0
    assert torch.equal(flat_p[:1], guard_lo) and torch.equal(flat_p[-2:], guard_hi)  # nothing outside the views was written


def test_bert_adam_in_training_step_refreshes_bf16_weights():
    """After an optimizer step the encoder must see the new weights (version-keyed bf16 caches)."""
    from visualbert_b200 import BertAdam, BertConfig, TrainVisualBERTObjective, synthetic
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    cfg = BertConfig.from_dict(synthetic.bert_config_dict(2, 128, 2, 512, vocab=512))
    model = TrainVisualBERTObjective(cfg, "pretraining", visual_embedding_dim=64).to(dev).eval()
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in
             synthetic.make_batch(4, 12, 7, 64, head="pretraining", vocab=512).items()}
    named = [(n, p) for n, p in model.named_parameters() if "pooler" not in n]
    no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
    opt = BertAdam([{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
                    {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}],
                   lr=1e-3, warmup=0.1, t_total=20)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = model(**batch)["loss"]
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0] - 0.05, losses  # the same batch six times: the loss must go down
