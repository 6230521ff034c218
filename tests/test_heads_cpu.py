"""The task heads, losses and accuracies of visualbert_b200.TrainVisualBERTObjective (PyTorch code above the CUDA
encoder) checked on the CPU for ALL six training_head_types: the encoder is replaced by a stub that returns the oracle's
(sequence_output, pooled_output), the result must match the golden produced by the unmodified reference. Run twice:
fp32 encoder outputs (tight tolerance) and bf16 sequence_output (the dtype the CUDA path hands to the heads)."""
import numpy as np
import pytest
import torch

import golden_util
import vb_oracle
from visualbert_b200 import BertConfig, TrainVisualBERTObjective

HEAD_CASES = ["small_ragged_pretraining", "small_vqa", "small_nlvr", "small_multichoice", "small_vqa_advanced", "small_flickr"]


def _model_with_stub_encoder(name, seq_dtype):
    cfg, sd, batch, c, gold = golden_util.load(name)
    model = TrainVisualBERTObjective(BertConfig.from_dict(cfg), c["head"], visual_embedding_dim=c["Dv"]).eval()
    res = model.load_state_dict(sd, strict=False)
    assert set(res.missing_keys) <= {"cls.predictions.decoder.weight"} and not res.unexpected_keys, res
    kw = {k: v for k, v in batch.items() if k != "position_embeddings_visual"}
    with torch.no_grad():
        ref = vb_oracle.objective(sd, cfg, c["head"], **kw)
    seq, pooled = ref["sequence_output"].to(seq_dtype), ref["pooled_output"]
    model.bert.forward = lambda *a, **k: (seq, pooled)
    return model, batch, gold, cfg


@pytest.mark.parametrize("name", HEAD_CASES)
def test_heads_match_reference_with_fp32_encoder_outputs(name):
    model, batch, gold, cfg = _model_with_stub_encoder(name, torch.float32)
    out = model(**batch)
    assert abs(float(out["loss"]) - float(gold["loss"])) <= 2e-5 * abs(float(gold["loss"]))
    for k in ("masked_lm_loss", "next_sentence_loss", "accuracy", "upperbound_accuracy", "entity_num"):
        if k in gold:
            assert abs(float(out[k]) - float(gold[k])) <= 2e-5 * max(abs(float(gold[k])), 1e-6), k
    if "logits" in out and out["logits"] is not None:
        a, b = golden_util.subsample(out["logits"].float()), gold["logits_sub"]
        assert np.abs(a - b).max() <= 2e-5 * max(np.abs(b).max(), 1e-12)
    if "nsp" in gold:
        assert np.abs(out["seq_relationship_score"].detach().numpy() - gold["nsp"]).max() <= 2e-5 * np.abs(gold["nsp"]).max()


@pytest.mark.parametrize("name", HEAD_CASES)
def test_heads_accept_bf16_sequence_output(name):
    """dtype handling only: with a bf16 sequence_output (what the CUDA encoder returns) every head must run and land
    within bf16 rounding of the reference loss."""
    model, batch, gold, cfg = _model_with_stub_encoder(name, torch.bfloat16)
    out = model(**batch)
    assert out["loss"].dtype == torch.float32
    assert abs(float(out["loss"]) - float(gold["loss"])) <= 2e-2 * abs(float(gold["loss"]))
    out["loss"].backward()  # the head parameters receive finite gradients
    grads = [p.grad for n, p in model.named_parameters() if p.grad is not None and not n.startswith("bert.")]
    assert grads and all(torch.isfinite(g).all() for g in grads)
