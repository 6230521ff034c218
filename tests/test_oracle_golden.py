"""Pin the CPU oracle (oracle/vb_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, produced by oracle/make_golden.py). fp32 on both sides: tolerance 2e-5 relative
to the tensor's max-abs (different summation order only)."""
import numpy as np
import pytest
import torch

import golden_util
import vb_oracle

CASE_NAMES = ["cfg1_pretraining", "small_ragged_pretraining", "small_vqa", "small_nlvr", "small_multichoice",
              "base3_ragged_pretraining", "small_vcr_alignment", "small_bypass_nlvr", "small_vqa_advanced", "small_flickr"]


def _close(a, b, rtol, what):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-12)
    err = np.abs(a - b).max() / scale
    assert err < rtol, f"{what}: rel-to-max err {err:.3e} >= {rtol}"


@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_matches_reference_outputs(name):
    cfg, sd, batch, c, gold = golden_util.load(name)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    kw = {k: v for k, v in batch.items() if k != "position_embeddings_visual"}
    out = vb_oracle.objective(sd, cfg, c["head"], **kw, **c.get("flags", {}))
    _close(out["loss"].item(), gold["loss"], 2e-5, "loss")
    for k in ("masked_lm_loss", "next_sentence_loss"):
        if k in gold:
            _close(out[k].item(), gold[k], 2e-5, k)
    if "logits" in out:  # the flickr head returns scores only through its loss / accuracy
        _close(golden_util.subsample(out["logits"]), gold["logits_sub"], 2e-5, "logits")
    for k in ("accuracy", "upperbound_accuracy", "entity_num"):
        if k in gold:
            _close(float(out[k]), gold[k], 1e-6, k)
    _close(out["pooled_output"].detach().numpy(), gold["pooled"], 2e-5, "pooled")
    if "nsp" in gold:
        _close(out["seq_relationship_score"].detach().numpy(), gold["nsp"], 2e-5, "nsp")
    last = cfg["num_hidden_layers"] - 1
    _close(golden_util.subsample(out["sequence_output"]), gold[f"hidden{last}_sub"], 2e-5, "last hidden")
    out["loss"].backward()
    norms = dict(zip(gold["grad_names"].tolist(), gold["grad_norms"].tolist()))
    for k, g in norms.items():
        if k == "cls.predictions.decoder.weight":
            continue
        mine = sd[k].grad
        assert mine is not None, k
        assert abs(mine.double().norm().item() - g) <= 5e-5 * max(g, 1e-6) + 1e-9, f"grad norm {k}"
    for key in [k for k in gold if k.startswith("grad_sub::")]:
        _close(golden_util.subsample(sd[key.split("::", 1)[1]].grad), gold[key], 5e-5, key)


def test_oracle_attention_weights_mode_matches_reference():
    """output_attention_weights=True (M.py:1430-1444): only the per-layer attention maps are returned."""
    cfg, sd, batch, c, gold = golden_util.load("small_attention_weights")
    kw = {k: v for k, v in batch.items() if k != "position_embeddings_visual"}
    out = vb_oracle.objective(sd, cfg, c["head"], **kw, **c["flags"])
    assert out["loss"] is None and set(out) == {"attention_weights", "loss"}
    assert len(out["attention_weights"]) == cfg["num_hidden_layers"]
    assert list(out["attention_weights"][0].shape) == gold["attn_shape"].tolist()
    for i, w in enumerate(out["attention_weights"]):
        _close(golden_util.subsample(w), gold[f"attn{i}_sub"], 2e-5, f"attention map {i}")


def test_fp64_oracle_agrees_with_fp32_reference():
    cfg, sd, batch, c, gold = golden_util.load("small_ragged_pretraining")
    sd64 = {k: v.double() for k, v in sd.items()}
    kw = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()
          if k != "position_embeddings_visual"}
    out = vb_oracle.objective(sd64, cfg, c["head"], **kw)
    _close(out["loss"].item(), gold["loss"], 1e-5, "loss fp64 vs ref fp32")
    _close(golden_util.subsample(out["logits"]), gold["logits_sub"], 1e-5, "logits fp64 vs ref fp32")
