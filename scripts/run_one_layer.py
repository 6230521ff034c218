"""One BertLayer forward+backward at the benchmark shape (cfg2: B=256, S=164, H=768) through the C ABI —
the short command ncu wraps (scripts/profile_ncu.sh). Usage: python scripts/run_one_layer.py [iters] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from visualbert_b200 import BertConfig, synthetic
from visualbert_b200.modeling import BertLayer
from visualbert_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
S, H = 164, 768
dev = torch.device("cuda:0")
cfg = BertConfig.from_dict(synthetic.bert_config_dict(1, H, 12, 3072))
torch.manual_seed(0)
layer = BertLayer(cfg, 0).to(dev).train()
x = torch.randn(B, S, H, device=dev).bfloat16().requires_grad_(True)
mask = torch.ones(B, S, dtype=torch.long, device=dev)
bias = ops.mask_bias(mask, None)
for i in range(iters):
    y = layer(x, bias, seed=i + 1)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("one-layer fwd+bwd done", iters, "iters, B =", B)
