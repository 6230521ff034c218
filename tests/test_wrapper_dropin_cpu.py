"""Wrapper-level drop-in proof (VERDICT r1 missing #5): the reference's own AllenNLP model wrapper
`visualbert/models/model.py:191-306` (VisualBERTFixedImageEmbedding) is executed UNMODIFIED except for the one-line
import swap INTEGRATION.md §1 prescribes (`model.py:20`), with a ~40-line test-only stand-in for the `allennlp` classes
it subclasses / instantiates (AllenNLP 0.8 is not installable offline).

What runs: the wrapper's constructor (from_pretrained with the reference's keyword list, special_intialize), its
forward (image-mask construction, the 13-keyword call into TrainVisualBERTObjective.forward, metric bookkeeping) and
get_metrics(), for the nlvr / vqa / pretraining heads; the state_dict keys the training wrapper groups and restores by
name (`model_wrapper.py:106-111, 201-221`). The CUDA encoder itself cannot run here (no GPU in the build container, and
/root/reference does not exist on the GPU box), so `self.bert.bert.forward` is served by the oracle's encoder — the
same arrangement as tests/test_heads_cpu.py — and the loss is compared with the reference golden."""
import os
import sys
import types

import pytest
import torch

import golden_util
import vb_oracle

REF_MODEL_PY = "/root/reference/visualbert/models/model.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF_MODEL_PY), reason="reference checkout not present (GPU box)")


def _install_standins(monkeypatch):
    """allennlp.* names model.py:11-18 imports; only Model (an nn.Module with a `register` decorator and a vocab) and the
    two metrics the wrapper instantiates need behaviour."""
    import visualbert_b200

    class Model(torch.nn.Module):
        def __init__(self, vocab=None):
            super().__init__()
            self.vocab = vocab

        @classmethod
        def register(cls, name):
            return lambda klass: klass

    class CategoricalAccuracy:
        def __init__(self):
            self.correct, self.total = 0.0, 0.0

        def __call__(self, logits, labels):
            self.correct += (logits.argmax(-1).view(-1) == labels.view(-1)).float().sum().item()
            self.total += labels.numel()

        def get_metric(self, reset=False):
            v = self.correct / max(self.total, 1.0)
            if reset:
                self.correct = self.total = 0.0
            return v

    class Average:
        def __init__(self):
            self.s, self.n = 0.0, 0

        def __call__(self, v):
            self.s += float(v); self.n += 1

        def get_metric(self, reset=False):
            v = self.s / max(self.n, 1)
            if reset:
                self.s, self.n = 0.0, 0
            return v

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    anything = type("Anything", (), {})
    mod("allennlp")
    mod("allennlp.data")
    mod("allennlp.data.vocabulary", Vocabulary=anything)
    mod("allennlp.models")
    mod("allennlp.models.model", Model=Model)
    mod("allennlp.modules", TextFieldEmbedder=anything, Seq2SeqEncoder=anything, FeedForward=anything,
        InputVariationalDropout=anything, TimeDistributed=anything)
    mod("allennlp.training")
    mod("allennlp.training.metrics", CategoricalAccuracy=CategoricalAccuracy, Average=Average)
    mod("allennlp.modules.matrix_attention", BilinearMatrixAttention=anything)
    mod("allennlp.nn", InitializerApplicator=anything)
    mod("allennlp.nn.util", masked_softmax=None, weighted_sum=None, replace_masked_values=None)
    # the ONE-LINE swap of INTEGRATION.md §1: model.py:20 imports TrainVisualBERTObjective from here
    mod("pytorch_pretrained_bert")
    mod("pytorch_pretrained_bert.modeling", TrainVisualBERTObjective=visualbert_b200.TrainVisualBERTObjective,
        BertForMultipleChoice=anything)
    mod("pytorch_pretrained_bert.file_utils", PYTORCH_PRETRAINED_BERT_CACHE="/tmp/vb_cache")


def _load_reference_wrapper(monkeypatch):
    _install_standins(monkeypatch)
    src = open(REF_MODEL_PY).read()
    ns = {"__name__": "reference_models_model"}
    exec(compile(src, REF_MODEL_PY, "exec"), ns)   # the reference file itself, not a copy
    return ns


@pytest.mark.parametrize("case", ["small_nlvr", "small_vqa", "small_ragged_pretraining"])
def test_reference_allennlp_wrapper_runs_on_the_swapped_objective(case, tmp_path, monkeypatch):
    import visualbert_b200
    from visualbert_b200 import BertConfig
    ns = _load_reference_wrapper(monkeypatch)
    cfg, sd, batch, c, gold = golden_util.load(case)
    (tmp_path / "bert_config.json").write_text(BertConfig.from_dict(cfg).to_json_string())
    Wrapper = ns["VisualBERTFixedImageEmbedding"]
    m = Wrapper(vocab=None, bert_model_name=str(tmp_path), training_head_type=c["head"], visual_embedding_dim=c["Dv"],
                random_initialize=True, special_visual_initialize=True)
    assert isinstance(m.bert, visualbert_b200.TrainVisualBERTObjective)
    # names the training wrapper relies on (model_wrapper.py:106-111 grouping, 209-221 restore-by-name)
    keys = set(m.state_dict().keys())
    assert "bert.bert.encoder.layer.0.attention.self.query.weight" in keys and "bert.bert.pooler.dense.weight" in keys
    assert any(k.endswith("LayerNorm.weight") for k in keys) and any("bias" in k for k in keys)
    res = m.bert.load_state_dict(sd, strict=False)
    assert set(res.missing_keys) <= {"cls.predictions.decoder.weight"} and not res.unexpected_keys
    m.eval()
    # CPU stand-in for the CUDA encoder: the oracle's (sequence_output, pooled_output) for this batch
    kw = {k: v for k, v in batch.items() if k != "position_embeddings_visual"}
    with torch.no_grad():
        ref = vb_oracle.objective(sd, cfg, c["head"], **kw)
    m.bert.bert.forward = lambda *a, **k: (ref["sequence_output"], ref["pooled_output"])
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)   # model.py:263 hard-codes .cuda()
    V = batch["visual_embeddings"].shape[-2]
    image_dim = batch["image_mask"].sum(-1)   # the wrapper rebuilds image_mask from the region count (model.py:262-268)
    out = m(bert_input_ids=batch["input_ids"], bert_input_mask=batch["input_mask"], bert_input_type_ids=batch["token_type_ids"],
            image_dim_variable=image_dim, image_feat_variable=batch["visual_embeddings"],
            visual_embeddings_type=batch.get("visual_embeddings_type"), label=batch.get("label"),
            masked_lm_labels=batch.get("masked_lm_labels"), is_random_next=batch.get("is_random_next"))
    assert abs(float(out["loss"]) - float(gold["loss"])) <= 2e-5 * abs(float(gold["loss"]))
    assert out["cnn_regularization_loss"] is None
    metrics = m.get_metrics(reset=True)
    assert "accuracy" in metrics and 0.0 <= metrics["accuracy"] <= 1.0
