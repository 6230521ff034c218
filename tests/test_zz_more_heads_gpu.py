"""GPU parity of the two task heads that have no encoder-specific behaviour of their own (`vqa_advanced`, `flickr`;
M.py:1527-1554, 1568-1598): loss / accuracy against the reference golden and the oracle, gradients against the oracle.
The head code itself is pinned on the CPU (tests/test_heads_cpu.py); here it runs above the CUDA encoder."""
import pytest
import torch

import golden_util
import vb_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["small_vqa_advanced", "small_flickr"])
def test_remaining_heads_on_the_cuda_encoder(name):
    from visualbert_b200 import BertConfig, TrainVisualBERTObjective
    cfg, sd, batch, c, gold = golden_util.load(name)
    dev = torch.device("cuda:0")
    model = TrainVisualBERTObjective(BertConfig.from_dict(cfg), c["head"], visual_embedding_dim=c["Dv"])
    model.load_state_dict(sd, strict=False)
    model.to(dev).eval()
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    out = model(**batch)
    loss = out["loss"]
    assert abs(loss.item() - float(gold["loss"])) <= 1e-2 * abs(float(gold["loss"]))
    if "entity_num" in gold:
        assert float(out["entity_num"]) == float(gold["entity_num"])
        assert abs(float(out["upperbound_accuracy"]) - float(gold["upperbound_accuracy"])) < 1e-5
    sdo = {k: v.to(dev).clone().requires_grad_(True) for k, v in sd.items()}
    kw = {k: v for k, v in batch.items() if k != "position_embeddings_visual"}
    ref = vb_oracle.objective(sdo, cfg, c["head"], **kw)
    assert abs(loss.item() - ref["loss"].item()) <= 1e-2 * abs(ref["loss"].item())
    # the same arithmetic in torch bf16 is the noise floor for ill-conditioned tensors of these tiny models
    sdb = {k: v.to(dev).bfloat16().clone().requires_grad_(True) for k, v in sd.items()}
    kwb = {k: (v.bfloat16() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in kw.items()}
    refb = vb_oracle.objective(sdb, cfg, c["head"], **kwb)
    loss.backward()
    ref["loss"].backward()
    refb["loss"].float().backward()
    checked = 0
    for k in ("bert.encoder.layer.0.attention.self.query.weight", "bert.encoder.layer.1.output.dense.weight",
              "bert.embeddings.projection.weight"):
        a, b = dict(model.named_parameters())[k].grad.float().reshape(-1), sdo[k].grad.reshape(-1)
        cos = torch.dot(a, b).item() / max(a.norm().item() * b.norm().item(), 1e-30)
        err, err_bf16 = (a - b).norm().item(), (sdb[k].grad.float().reshape(-1) - b).norm().item()
        assert cos > 0.98 or err <= err_bf16, f"{k}: cosine {cos:.4f}, err {err:.3e} vs torch-bf16 {err_bf16:.3e}"
        checked += 1
    assert checked == 3
