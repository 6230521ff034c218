"""PreTrainedBertModel.from_pretrained (M.py:486-596) on local archives: directory and .tar.gz, TF-era gamma/beta names,
the "bert." prefix rule for a bare encoder, random_initialize, and the no-network error. Host logic only (no GPU)."""
import json
import os
import tarfile

import pytest
import torch

import golden_util  # noqa: F401
from visualbert_b200 import BertVisualModel, TrainVisualBERTObjective, synthetic

CFG = synthetic.bert_config_dict(2, 128, 2, 512, vocab=512)


def _write_archive(tmp_path, sd, tf_names=False, extra_cfg=None):
    d = tmp_path / "ckpt"
    d.mkdir()
    (d / "bert_config.json").write_text(json.dumps(dict(CFG, **(extra_cfg or {}))))
    if tf_names:
        sd = {k.replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta"): v for k, v in sd.items()}
    torch.save(sd, str(d / "pytorch_model.bin"))
    return d


@pytest.mark.parametrize("tf_names", [False, True])
def test_from_pretrained_directory_and_tf_era_names(tmp_path, tf_names):
    sd = synthetic.init_state_dict(CFG, "pretraining", 64, seed=3)
    d = _write_archive(tmp_path, sd, tf_names)
    model = TrainVisualBERTObjective.from_pretrained(str(d), training_head_type="pretraining", visual_embedding_dim=64)
    got = model.state_dict()
    for k, v in sd.items():
        assert torch.equal(got[k], v), k
    assert model.cls.predictions.decoder.weight is model.bert.embeddings.word_embeddings.weight


def test_from_pretrained_tar_gz_and_cache_dir(tmp_path):
    sd = synthetic.init_state_dict(CFG, "nlvr", 64, seed=4)
    d = _write_archive(tmp_path, sd)
    tar = tmp_path / "cache" / "model.tar.gz"
    tar.parent.mkdir()
    with tarfile.open(str(tar), "w:gz") as t:
        t.add(str(d / "bert_config.json"), arcname="bert_config.json")
        t.add(str(d / "pytorch_model.bin"), arcname="pytorch_model.bin")
    model = TrainVisualBERTObjective.from_pretrained("model.tar.gz", cache_dir=str(tar.parent), training_head_type="nlvr",
                                                     visual_embedding_dim=64)
    assert torch.equal(model.state_dict()["classifier.weight"], sd["classifier.weight"])
    assert torch.equal(model.state_dict()["bert.encoder.layer.1.output.dense.weight"], sd["bert.encoder.layer.1.output.dense.weight"])


def test_bare_encoder_reads_the_bert_prefixed_entries(tmp_path):
    sd = synthetic.init_state_dict(CFG, "pretraining", 64, seed=5)
    # a bare BertVisualModel reads the visual settings from its config (TrainVisualBERTObjective writes them there, M.py:1340-1344)
    d = _write_archive(tmp_path, sd, extra_cfg=dict(visual_embedding_dim=64, bypass_transformer=False, output_attention_weights=False))
    bare = BertVisualModel.from_pretrained(str(d))
    got = bare.state_dict()
    for k, v in sd.items():
        if k.startswith("bert.") and got[k[5:]].shape == v.shape:
            assert torch.equal(got[k[5:]], v), k
    assert torch.equal(got["encoder.layer.0.attention.self.query.weight"], sd["bert.encoder.layer.0.attention.self.query.weight"])


def test_random_initialize_and_missing_archive(tmp_path):
    sd = synthetic.init_state_dict(CFG, "nlvr", 64, seed=6)
    d = _write_archive(tmp_path, sd)
    model = TrainVisualBERTObjective.from_pretrained(str(d), random_initialize=True, training_head_type="nlvr", visual_embedding_dim=64)
    assert not torch.equal(model.state_dict()["classifier.weight"], sd["classifier.weight"])
    with pytest.raises(EnvironmentError):
        TrainVisualBERTObjective.from_pretrained("bert-base-uncased", training_head_type="nlvr", visual_embedding_dim=64)
