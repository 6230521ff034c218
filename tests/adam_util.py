"""Seeded BertAdam scenario shared by the golden generator (oracle/make_golden.py::make_bert_adam_golden), the CPU
oracle test and the GPU parity test: same shapes, groups, hyper-parameters and gradient stream."""
import torch

SHAPES = [(33, 17), (128,), (7, 5, 3), (1,), (4096,)]
GRAD_SCALE = (0.01, 3.0, 0.3, 5.0, 0.002)
WEIGHT_DECAY = (0.01, 0.01, 0.01, 0.0, 0.0)      # params[:3] decay group, params[3:] no-decay group
HYPER = dict(lr=5e-3, warmup=0.25, t_total=8)
STEPS = 4


def scenario():
    """-> (initial params, [per-step list of grads])"""
    g = torch.Generator().manual_seed(77)
    params = [torch.randn(*s, generator=g) * 0.05 for s in SHAPES]
    grads = [[torch.randn(*s, generator=g) * GRAD_SCALE[i] for i, s in enumerate(SHAPES)] for _ in range(STEPS)]
    return params, grads
