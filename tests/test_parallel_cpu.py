"""world_size-2 gloo test of the data-parallel plumbing (no GPU): FlatGradSync's single all-reduce equals
single-process gradients on the concatenated batch when each rank holds the same number of loss terms."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from visualbert_b200.parallel import FlatGradSync, shard_batch
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    model[2].weight = model[2].weight  # no-op; keeps parameters() order deterministic
    x = torch.randn(12, 8); y = torch.randn(12, 3)
    sync = FlatGradSync(model)
    calls = []
    orig = dist.all_reduce
    dist.all_reduce = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    sync.zero()
    b = shard_batch({"x": x, "y": y, "tag": "keep"}, rank, world)
    assert b["tag"] == "keep" and b["x"].shape[0] == 6
    torch.nn.functional.mse_loss(model(b["x"]), b["y"]).backward()
    sync.allreduce()
    flat = torch.cat([v.reshape(-1) for v in sync.views]).clone()  # the views, without the alignment padding
    dist.all_reduce = orig
    assert len(calls) == 1, "exactly one collective per step"
    for p in model.parameters():
        assert p.grad.data_ptr() >= sync.flat.data_ptr()  # grads live inside the flat buffer
        assert p.grad.data_ptr() % 256 == sync.flat.data_ptr() % 256  # every view starts on a 256-byte boundary
    if rank == 0:
        ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
        ref.load_state_dict(model.state_dict())
        torch.nn.functional.mse_loss(ref(x), y).backward()
        want = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
        q.put(float((flat - want).abs().max()))
    dist.destroy_process_group()


def test_flat_grad_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-6


def test_flat_grad_sync_single_process_and_tied_weights():
    sys.path.insert(0, ROOT)
    from visualbert_b200.parallel import FlatGradSync
    emb = torch.nn.Embedding(10, 4)
    dec = torch.nn.Linear(4, 10, bias=False)
    dec.weight = emb.weight
    m = torch.nn.ModuleList([emb, dec])
    s = FlatGradSync(m)
    assert s.flat.numel() == 40
    dec(emb(torch.tensor([1, 2]))).sum().backward()
    assert s.flat.abs().sum() > 0
    s.zero()
    assert s.flat.abs().sum() == 0 and emb.weight.grad is s.views[0]
    s.allreduce()  # no process group: no-op


def test_flat_grad_layout_keeps_qkv_adjacent_and_everything_else_aligned():
    sys.path.insert(0, ROOT)
    from visualbert_b200 import BertConfig, TrainVisualBERTObjective, synthetic
    from visualbert_b200.parallel import FlatGradSync
    cfg = BertConfig.from_dict(synthetic.bert_config_dict(2, 128, 2, 512, vocab=515))  # odd vocab: misaligning sizes
    model = TrainVisualBERTObjective(cfg, "vqa", visual_embedding_dim=64)               # classifier bias 3129 (odd)
    s = FlatGradSync(model)
    base = s.flat.data_ptr()
    for layer in model.bert.encoder.layer:
        a = layer.attention.self
        q, k, v = a.query.weight.grad, a.key.weight.grad, a.value.weight.grad
        assert k.data_ptr() == q.data_ptr() + q.numel() * 4 and v.data_ptr() == k.data_ptr() + k.numel() * 4
        bq, bk, bv = a.query.bias.grad, a.key.bias.grad, a.value.bias.grad
        assert bk.data_ptr() == bq.data_ptr() + bq.numel() * 4 and bv.data_ptr() == bk.data_ptr() + bk.numel() * 4
    for p in model.parameters():
        if p.requires_grad:
            assert (p.grad.data_ptr() - base) % 16 == 0, "16-byte vector access must stay legal"


def test_host_side_mlm_rows_equal_the_padded_label_scan():
    """BatchPrefetcher.labelled_rows (host) == nonzero over the labels padded with -1 on the region positions, which is
    what TrainVisualBERTObjective.forward scans on the device when `masked_lm_rows` is not given."""
    sys.path.insert(0, ROOT)
    from visualbert_b200 import synthetic
    from visualbert_b200.parallel import BatchPrefetcher
    for choices in (None, 3):
        b = synthetic.make_batch(5, 12, 7, 16, head="pretraining", vocab=512, ragged=True, choices=choices)
        labels = b["masked_lm_labels"].reshape(-1, 12)
        padded = torch.cat((labels, torch.full((labels.shape[0], 7), -1, dtype=labels.dtype)), dim=1)
        want = torch.nonzero(padded.reshape(-1) != -1).squeeze(1)
        assert torch.equal(BatchPrefetcher.labelled_rows(b), want)
    assert BatchPrefetcher.labelled_rows({"input_ids": torch.zeros(2, 3)}) is None


def _seed_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from visualbert_b200 import BertConfig, TrainVisualBERTObjective, synthetic
    torch.manual_seed(1234)
    cfg = synthetic.bert_config_dict(1, 64, 1, 128, vocab=64)
    m = TrainVisualBERTObjective(BertConfig.from_dict(cfg), "pretraining", visual_embedding_dim=16).bert
    q.put((rank, m.dropout_state()["seed"], m.next_seed(), m.next_seed()))
    dist.destroy_process_group()


def test_dropout_seed_follows_manual_seed_differs_per_rank_and_round_trips():
    """the encoder dropout used a constant seed — identical on every data-parallel rank, blind to
    torch.manual_seed, and not resumable. Now: base seed from torch.initial_seed(), rank mixed in per forward,
    dropout_state() / set_dropout_state() carry (seed, step)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_seed_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    (r0, base0, a0, b0), (r1, base1, a1, b1) = got
    assert base0 == base1                      # same torch.manual_seed -> same base seed on both ranks
    assert a0 != a1 and b0 != b1 and a0 != b0  # ... but different streams per rank and per step
    sys.path.insert(0, ROOT)
    from visualbert_b200 import BertConfig, TrainVisualBERTObjective, synthetic
    cfg = synthetic.bert_config_dict(1, 64, 1, 128, vocab=64)
    torch.manual_seed(1)
    m1 = TrainVisualBERTObjective(BertConfig.from_dict(cfg), "pretraining", visual_embedding_dim=16).bert
    torch.manual_seed(2)
    m2 = TrainVisualBERTObjective(BertConfig.from_dict(cfg), "pretraining", visual_embedding_dim=16).bert
    assert m1.dropout_state()["seed"] != m2.dropout_state()["seed"]
    m1.next_seed(); m1.next_seed()
    st = m1.dropout_state()
    want = m1.next_seed()
    m2.set_dropout_state(st)
    assert m2.next_seed() == want and "dropout_seed" not in m1.state_dict()
