"""Per-kernel-family timing of one BertLayer fwd+bwd at the benchmark shape under the live profiler, for a set
of tuning-knob settings (env vars read by the library). Usage: python scripts/gpu_tune_layer.py [name=env,...]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    from visualbert_b200 import BertConfig, synthetic, ops, _lib
    from visualbert_b200.modeling import BertLayer
    B, S, H = int(os.environ.get("TUNE_B", 256)), int(os.environ.get("TUNE_S", 164)), 768
    dev = torch.device("cuda:0")
    cfg = BertConfig.from_dict(synthetic.bert_config_dict(1, H, 12, 3072))
    torch.manual_seed(0)
    layer = BertLayer(cfg, 0).to(dev).train(not os.environ.get("TUNE_EVAL"))
    x = torch.randn(B, S, H, device=dev).bfloat16().requires_grad_(True)
    bias = ops.mask_bias(torch.ones(B, S, dtype=torch.long, device=dev), None)
    dy = torch.randn(B, S, H, device=dev).bfloat16()
    iters = 12
    for i in range(3):
        layer(x, bias, seed=i + 1).backward(dy)
    _lib.profile_read(); _lib.profile_enable(True)
    for i in range(iters):
        layer(x, bias, seed=i + 10).backward(dy)
    prof = _lib.profile_read()
    out = {k: round(1e3 * v["ms"] / iters, 1) for k, v in prof.items() if v["launches"]}
    out["total_us"] = round(sum(out.values()), 1)
    print("TUNE " + json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("TUNE_CHILD"):
        child()
    else:
        settings = sys.argv[1:] or ["base="]
        for sname in settings:
            name, _, envs = sname.partition("=")
            env = dict(os.environ, TUNE_CHILD="1")
            for kv in filter(None, envs.split(",")):
                k, _, v = kv.partition(":")
                env[k] = v
            r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=300)
            line = [l for l in r.stdout.splitlines() if l.startswith("TUNE ")]
            print(name.ljust(14), line[0][5:] if line else ("FAILED: " + (r.stderr.strip().splitlines() or ["?"])[-1]))
            sys.stdout.flush()
