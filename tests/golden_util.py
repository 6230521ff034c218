"""Rebuild the (config, weights, batch) of a golden case from its seeds; load the stored outputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from visualbert_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def cases():
    import make_golden
    return make_golden.CASES


def load(name):
    c = cases()[name]
    m = c["model"]
    cfg = synthetic.bert_config_dict(m["layers"], m["hidden"], m["heads"], m["inter"], vocab=m["vocab"])
    sd = synthetic.init_state_dict(cfg, c["head"], c["Dv"], seed=0,
                                   bypass_transformer=c.get("flags", {}).get("bypass_transformer", False))
    batch = synthetic.make_batch(Dv=c["Dv"], head=c["head"], seed=1234, vocab=m["vocab"], **c["batch"])
    gold = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    return cfg, sd, batch, c, gold


def subsample(t, n=4096):
    flat = t.detach().reshape(-1)
    step = max(1, flat.numel() // n)
    return flat[::step][:n].double().cpu().numpy()
