"""Per-tensor gradient diagnostics of the CUDA path vs the fp32 oracle, with the oracle-in-bf16 as noise baseline."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import golden_util, vb_oracle
from visualbert_b200 import BertConfig, TrainVisualBERTObjective

def run(name):
    cfg, sd, batch, c, gold = golden_util.load(name)
    dev = torch.device("cuda:0")
    model = TrainVisualBERTObjective(BertConfig.from_dict(cfg), c["head"], visual_embedding_dim=c["Dv"])
    model.load_state_dict(sd, strict=False); model.to(dev).eval()
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    kw = {k: v for k, v in batch.items() if k != "position_embeddings_visual"}
    out = model(**batch); out["loss"].backward()
    sdo = {k: v.to(dev).clone().requires_grad_(True) for k, v in sd.items()}
    ref = vb_oracle.objective(sdo, cfg, c["head"], **kw); ref["loss"].backward()
    sdb = {k: v.to(dev).bfloat16().clone().requires_grad_(True) for k, v in sd.items()}
    kwb = {k: (v.bfloat16() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in kw.items()}
    refb = vb_oracle.objective(sdb, cfg, c["head"], **kwb); refb["loss"].float().backward()
    print(f"=== {name}: loss mine={out['loss'].item():.6f} ref={ref['loss'].item():.6f} ref_bf16={refb['loss'].item():.6f}")
    for k, p in model.named_parameters():
        if k == "cls.predictions.decoder.weight" or sdo[k].grad is None: continue
        a = p.grad.float().reshape(-1); b = sdo[k].grad.float().reshape(-1); cb = sdb[k].grad.float().reshape(-1)
        nb = b.norm().item()
        cos = torch.dot(a, b).item() / max(a.norm().item() * nb, 1e-30)
        cosb = torch.dot(cb, b).item() / max(cb.norm().item() * nb, 1e-30)
        print(f"  {k:60s} |ref|={nb:.3e} cos={cos:.5f} relerr={(a-b).norm().item()/max(nb,1e-30):.3e} | torch-bf16: cos={cosb:.5f} relerr={(cb-b).norm().item()/max(nb,1e-30):.3e}")

for n in (sys.argv[1:] or ["small_multichoice", "small_ragged_pretraining"]):
    run(n)
