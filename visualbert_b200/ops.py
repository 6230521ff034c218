"""torch.autograd bindings of the libvbert_b200 C ABI (include/vbert_b200.h).

PyTorch supplies device memory, the current stream and the autograd graph; every FLOP of the
encoder path runs in the sm_100a kernels. There is no fallback: on a machine without the library or
without a CUDA device these ops raise.

Activations are bf16; parameters are the model's fp32 master weights, cast to bf16 "compute weights"
by vb_cast_f32_to_bf16 (cached per parameter version). Parameter gradients come back in fp32.
"""
import ctypes

import torch

from . import _lib

_BF16 = torch.bfloat16


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _require_cuda(t, what):
    if not t.is_cuda:
        raise _lib.VBertLibraryError(
            f"{what}: tensor is on {t.device}; visualbert_b200 runs only on CUDA (sm_100a) — no CPU fallback")


# --------------------------------------------------------------------------------------------
# workspaces: backward scratch is shared by all layers of a step (same stream, sequential use)
# --------------------------------------------------------------------------------------------
_scratch_cache = {}


def _scratch(dev, M, H, I, B, A, S, need_drop):
    key = (dev, M, H, I, B, A, S)
    w = _scratch_cache.get(key)
    if w is None:
        w = dict(
            d_pre=torch.empty(M, H, device=dev, dtype=_BF16),
            d_pre_drop=None,
            d_big=torch.empty(M, max(I, 3 * H), device=dev, dtype=_BF16),
            d_x1=torch.empty(M, H, device=dev, dtype=_BF16),
            d_ctx=torch.empty(M, H, device=dev, dtype=_BF16),
            drow=torch.empty(B, A, S, device=dev, dtype=torch.float32))
        for k in [k for k in _scratch_cache if k[0] == dev]:   # keep a single shape resident PER DEVICE
            del _scratch_cache[k]
        _scratch_cache[key] = w
    if need_drop and w["d_pre_drop"] is None:
        w["d_pre_drop"] = torch.empty(M, H, device=dev, dtype=_BF16)
    return w


def cast_to_bf16(src, out=None):
    """fp32 -> bf16 through vb_cast_f32_to_bf16 (numel must be a multiple of 8)."""
    _require_cuda(src, "cast_to_bf16")
    src = src.contiguous()
    if out is None:
        out = torch.empty(src.shape, device=src.device, dtype=_BF16)
    with torch.cuda.device(src.device):   # the library launches on the CURRENT device's stream: make it the tensor's
        _lib.check(_lib.lib().vb_cast_f32_to_bf16(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                  ctypes.c_int64(src.numel()), _stream()), "vb_cast_f32_to_bf16")
    return out


def mask_bias(input_mask, image_mask):
    """(1 - cat(input_mask, image_mask)) * -10000 as fp32 [B, T+V] (reference M.py:1417, 1286-1294)."""
    _require_cuda(input_mask, "mask_bias")
    B, T = input_mask.shape
    V = 0 if image_mask is None else image_mask.shape[1]
    im = input_mask.to(torch.int64).contiguous()
    vm = None if image_mask is None else image_mask.to(torch.int64).contiguous()
    out = torch.empty(B, T + V, device=input_mask.device, dtype=torch.float32)
    with torch.cuda.device(input_mask.device):
        _lib.check(_lib.lib().vb_mask_bias(ctypes.c_void_p(im.data_ptr()), ctypes.c_void_p(_ptr(vm)),
                                           ctypes.c_void_p(out.data_ptr()), B, T, V, _stream()), "vb_mask_bias")
    return out


def _grad_targets(params, adjacent_groups=()):
    """Direct-accumulate mode — OPT-IN: only parameters a gradient owner has flagged with `_vb_direct_grad = True`
    (parallel.FlatGradSync does, for the views of its flat buffer) get their gradients written straight into `.grad`
    by the kernels (all gradient outputs of the C ABI are `+=`), with autograd receiving None. Everyone else — plain
    optimizers, DDP with gradient_as_bucket_view, tensor / post-accumulate hooks, torch.autograd.grad — goes through
    AccumulateGrad as usual: the caller allocates fresh zero buffers and returns them to autograd.
    Additionally every `.grad` must be a contiguous fp32 tensor of the parameter's shape, and the parameters of each
    adjacent group must be laid out back to back."""
    for p in params:
        g = p.grad
        if (not getattr(p, "_vb_direct_grad", False) or g is None or g.dtype != torch.float32 or not g.is_contiguous()
                or g.device != p.device or g.shape != p.shape):
            return None
    for group in adjacent_groups:
        for a, b in zip(group[:-1], group[1:]):
            if a.grad.data_ptr() + a.grad.numel() * 4 != b.grad.data_ptr():
                return None
    return [p.grad for p in params]


class WeightBank:
    """Every bf16 compute copy the CUDA path reads (packed q|k|v, attention-output, FFN matrices of all layers, the visual
    projection, the tied MLM decoder table) plus the fp32 packed q|k|v biases, refreshed from the fp32 master
    parameters by ONE vb_cast_multi launch.

    `refresh(force=True)` is called at the start of every training-mode forward: the masters may have been changed by
    anything — the reference BertAdam updates through `p.data` (optimization.py:293), which does not bump
    `Tensor._version`, so version-keyed caching alone would silently train on stale bf16 weights. In eval mode the copy
    is redone only when a (data_ptr, _version) signature changed (load_state_dict, manual edits through autograd-visible
    ops)."""

    def __init__(self):
        self.sig = None
        self.items = []       # (src parameter, dst tensor, dst_is_fp32)
        self.table = None
        self.n_chunks = 0
        self.keep = []        # owners of the dst storage
        self.generation = 0

    def _signature(self, params):
        return tuple((p.data_ptr(), p._version) for p in params)

    def bind(self, items, keep):
        """items: list of (src fp32 parameter, dst tensor [contiguous view], dst_is_fp32)."""
        import numpy as np
        self.items = items
        self.keep = keep
        arr = (_lib.CastItem * len(items))()
        chunk = 0
        for i, (src, dst, f32) in enumerate(items):
            assert src.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous() and dst.numel() == src.numel()
            arr[i].src, arr[i].dst, arr[i].numel = src.data_ptr(), dst.data_ptr(), src.numel()
            arr[i].first_chunk, arr[i].dst_fp32 = chunk, 1 if f32 else 0
            chunk += (src.numel() + _lib.VB_CAST_CHUNK - 1) // _lib.VB_CAST_CHUNK
        self.n_chunks = chunk
        raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        self.table = torch.from_numpy(raw).to(items[0][0].device)
        self.ptrs = tuple(src.data_ptr() for src, _, _ in items)
        self.sig = None

    def bound_to(self, params):
        return self.table is not None and self.ptrs == tuple(p.data_ptr() for p in params)

    def refresh(self, force):
        srcs = [s_ for s_, _, _ in self.items]
        sig = self._signature(srcs)
        if not force and sig == self.sig:
            return
        with torch.cuda.device(self.table.device):
            _lib.check(_lib.lib().vb_cast_multi(ctypes.c_void_p(self.table.data_ptr()), len(self.items), self.n_chunks, _stream()),
                       "vb_cast_multi")
        self.sig = sig
        self.generation += 1


class LayerWeights:
    """bf16 compute copies of one BertLayer's matrices. Normally views into the model's WeightBank (refreshed once per
    forward by the model); a stand-alone BertLayer refreshes them itself — on every training-mode forward, or when the
    masters' (data_ptr, _version) signature changes in eval mode."""

    def __init__(self):
        self.key = None
        self.buf = None
        self.bank = None      # set by the owning model: buffers are bank views, refresh is the bank's job

    def get(self, q, k, v, o, w1, w2, bq, bk, bv, train=False):
        if self.bank is not None:
            return self.buf
        key = tuple((p.data_ptr(), p._version) for p in (q, k, v, o, w1, w2, bq, bk, bv))
        if train or key != self.key:
            H, I = o.shape[0], w1.shape[0]
            dev = q.device
            if self.buf is None or self.buf[0].device != dev:
                self.buf = (torch.empty(3 * H, H, device=dev, dtype=_BF16), torch.empty(H, H, device=dev, dtype=_BF16),
                            torch.empty(I, H, device=dev, dtype=_BF16), torch.empty(H, I, device=dev, dtype=_BF16),
                            torch.empty(3 * H, device=dev, dtype=torch.float32))
            wqkv, wo, wi, wout, bqkv = self.buf
            with torch.no_grad():
                cast_to_bf16(q.detach(), wqkv[0:H])
                cast_to_bf16(k.detach(), wqkv[H:2 * H])
                cast_to_bf16(v.detach(), wqkv[2 * H:3 * H])
                cast_to_bf16(o.detach(), wo)
                cast_to_bf16(w1.detach(), wi)
                cast_to_bf16(w2.detach(), wout)
                torch.cat((bq.detach(), bk.detach(), bv.detach()), out=bqkv)
            self.key = key
        return self.buf


class _LayerFn(torch.autograd.Function):
    """BertLayer forward/backward (reference M.py:322-341) through vb_layer_fwd / vb_layer_bwd."""

    @staticmethod
    def forward(ctx, x, mbias, meta, qw, qb, kw, kb, vw, vb, ow, ob, g1, b1, iw, ib, dw, db, g2, b2):
        # meta: dict(heads, layer_index, hidden_dropout, attn_dropout, seed, cache=LayerWeights)
        _require_cuda(x, "bert_layer")
        B, S, H = x.shape
        I = iw.shape[0]
        M = B * S
        A = meta["heads"]
        dev = x.device
        x = x.contiguous()
        wqkv, wo, wi, wout, bqkv = meta["cache"].get(qw, kw, vw, ow, iw, dw, qb, kb, vb, train=meta.get("train", False))
        f32 = torch.float32
        acts = dict(
            qkv=torch.empty(M, 3 * H, device=dev, dtype=_BF16), ctx=torch.empty(M, H, device=dev, dtype=_BF16),
            lse=torch.empty(B, A, S, device=dev, dtype=f32), pre1=torch.empty(M, H, device=dev, dtype=_BF16),
            mean1=torch.empty(M, device=dev, dtype=f32), rstd1=torch.empty(M, device=dev, dtype=f32),
            x1=torch.empty(M, H, device=dev, dtype=_BF16), u=torch.empty(M, I, device=dev, dtype=_BF16),
            g=torch.empty(M, I, device=dev, dtype=_BF16), pre2=torch.empty(M, H, device=dev, dtype=_BF16),
            mean2=torch.empty(M, device=dev, dtype=f32), rstd2=torch.empty(M, device=dev, dtype=f32),
            keep_mask=(torch.empty(int(_lib.lib().vb_attention_keep_bytes(B, S, A)), device=dev, dtype=torch.uint8)
                       if meta["attn_dropout"] > 0 else None))
        y = torch.empty(B, S, H, device=dev, dtype=_BF16)
        d = _lib.LayerDesc(
            batch=B, seq=S, hidden=H, heads=A, inter=I, hidden_dropout=meta["hidden_dropout"],
            attn_dropout=meta["attn_dropout"], seed=meta["seed"], layer_index=meta["layer_index"],
            w_qkv=wqkv.data_ptr(), w_attn_out=wo.data_ptr(), w_inter=wi.data_ptr(), w_out=wout.data_ptr(),
            b_qkv=bqkv.data_ptr(), b_attn_out=ob.data_ptr(), ln1_gamma=g1.data_ptr(), ln1_beta=b1.data_ptr(),
            b_inter=ib.data_ptr(), b_out=db.data_ptr(), ln2_gamma=g2.data_ptr(), ln2_beta=b2.data_ptr(),
            mask_bias=mbias.data_ptr())
        a = _lib.LayerActs(**{k: _ptr(t) for k, t in acts.items()})
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().vb_layer_fwd(ctypes.byref(d), ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()),
                                               ctypes.byref(a), _stream()), "vb_layer_fwd")
        ctx.meta = meta
        ctx.acts = acts
        ctx.params = (qw, qb, kw, kb, vw, vb, ow, ob, g1, b1, iw, ib, dw, db, g2, b2)
        ctx.weights = (wqkv, wo, wi, wout, bqkv)
        ctx.weight_key = meta["cache"].key if meta["cache"].bank is None else None
        ctx.save_for_backward(x, mbias, ob, g1, b1, ib, db, g2, b2)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mbias, ob, g1, b1, ib, db, g2, b2 = ctx.saved_tensors
        meta, acts = ctx.meta, ctx.acts
        if ctx.weight_key is not None and meta["cache"].key != ctx.weight_key:
            raise RuntimeError("visualbert_b200: layer weights were modified between forward and backward")
        wqkv, wo, wi, wout, bqkv = ctx.weights
        B, S, H = x.shape
        I = wi.shape[0]
        M, A = B * S, meta["heads"]
        dev = x.device
        dy = dy.to(_BF16).contiguous()
        qw, qb, kw, kb, vw, vb, ow, ob_, g1_, b1_, iw, ib_, dw, db_, g2_, b2_ = ctx.params
        direct = _grad_targets(ctx.params, ((qw, kw, vw), (qb, kb, vb)))
        gnames = ("dw_qkv", "db_qkv", "dw_attn_out", "db_attn_out", "dln1_gamma", "dln1_beta",
                  "dw_inter", "db_inter", "dw_out", "db_out", "dln2_gamma", "dln2_beta")
        if direct is not None:
            tg = dict(zip(("qw", "qb", "kw", "kb", "vw", "vb", "ow", "ob", "g1", "b1", "iw", "ib", "dw", "db", "g2", "b2"), direct))
            parts = [tg["qw"], tg["qb"], tg["ow"], tg["ob"], tg["g1"], tg["b1"], tg["iw"], tg["ib"], tg["dw"], tg["db"], tg["g2"], tg["b2"]]
        else:
            sizes = [3 * H * H, 3 * H, H * H, H, H, H, I * H, I, H * I, H, H, H]
            flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
            parts = list(torch.split(flat, sizes))
        g = _lib.LayerGrads(**{n: t.data_ptr() for n, t in zip(gnames, parts)})
        hd = meta["hidden_dropout"] > 0
        w = _scratch(dev, M, H, I, B, A, S, hd)
        sc = _lib.LayerScratch(**{k: _ptr(t) for k, t in w.items()})
        d = _lib.LayerDesc(
            batch=B, seq=S, hidden=H, heads=A, inter=I, hidden_dropout=meta["hidden_dropout"],
            attn_dropout=meta["attn_dropout"], seed=meta["seed"], layer_index=meta["layer_index"],
            w_qkv=wqkv.data_ptr(), w_attn_out=wo.data_ptr(), w_inter=wi.data_ptr(), w_out=wout.data_ptr(),
            b_qkv=bqkv.data_ptr(), b_attn_out=ob.data_ptr(), ln1_gamma=g1.data_ptr(), ln1_beta=b1.data_ptr(),
            b_inter=ib.data_ptr(), b_out=db.data_ptr(), ln2_gamma=g2.data_ptr(), ln2_beta=b2.data_ptr(),
            mask_bias=mbias.data_ptr())
        a = _lib.LayerActs(**{k: _ptr(t) for k, t in acts.items()})
        dx = torch.empty(B, S, H, device=dev, dtype=_BF16)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().vb_layer_bwd(ctypes.byref(d), ctypes.c_void_p(x.data_ptr()), ctypes.byref(a),
                                               ctypes.c_void_p(dy.data_ptr()), ctypes.c_void_p(dx.data_ptr()),
                                               ctypes.byref(g), ctypes.byref(sc), _stream()), "vb_layer_bwd")
        ctx.acts = None
        if direct is not None:
            return (dx, None, None) + (None,) * 16
        dwqkv, dbqkv, dwo, dbo, dg1, db1, dwi, dbi, dwout, dbout, dg2, db2 = parts
        dwq, dwk, dwv = dwqkv.view(3, H, H).unbind(0)
        dbq, dbk, dbv = dbqkv.view(3, H).unbind(0)
        return (dx, None, None, dwq, dbq, dwk, dbk, dwv, dbv, dwo.view(H, H), dbo, dg1, db1,
                dwi.view(I, H), dbi, dwout.view(H, I), dbout, dg2, db2)


class EncoderPlan:
    """Host-side state of the whole-encoder call (vb_encoder_fwd / vb_encoder_bwd): the ctypes descriptor and gradient
    arrays are built once per (shape, weights) and only their per-step fields (seed, dropout) are touched afterwards."""

    def __init__(self):
        self.key = None

    def prepare(self, caches, params, B, S, H, A, I, mbias, meta):
        L = len(caches)
        weights = [c.get(*[params[16 * l + i] for i in (0, 2, 4, 6, 10, 12, 1, 3, 5)], train=meta["train"]) for l, c in enumerate(caches)]
        key = (B, S, H, A, I, L, tuple(w[0].data_ptr() for w in weights), tuple(p.data_ptr() for p in params), mbias.device)
        if key != self.key:
            self.descs = (_lib.LayerDesc * L)()
            for l in range(L):
                wqkv, wo, wi, wout, bqkv = weights[l]
                qw, qb, kw, kb, vw, vb, ow, ob, g1, b1, iw, ib, dw, db, g2, b2 = params[16 * l: 16 * l + 16]
                d = self.descs[l]
                d.batch, d.seq, d.hidden, d.heads, d.inter = B, S, H, A, I
                d.w_qkv, d.w_attn_out, d.w_inter, d.w_out = wqkv.data_ptr(), wo.data_ptr(), wi.data_ptr(), wout.data_ptr()
                d.b_qkv, d.b_attn_out, d.ln1_gamma, d.ln1_beta = bqkv.data_ptr(), ob.data_ptr(), g1.data_ptr(), b1.data_ptr()
                d.b_inter, d.b_out, d.ln2_gamma, d.ln2_beta = ib.data_ptr(), db.data_ptr(), g2.data_ptr(), b2.data_ptr()
            self.key = key
        for l in range(L):
            d = self.descs[l]
            d.hidden_dropout, d.attn_dropout, d.seed = meta["hidden_dropout"], meta["attn_dropout"], meta["seed"]
            d.layer_index = meta["layer_index0"] + l
            d.mask_bias = mbias.data_ptr()
        off = (ctypes.c_int64 * _lib.VB_ENCODER_ARENA_BUFFERS)()
        stride = int(_lib.lib().vb_encoder_arena_layout(B, S, H, A, I, 1 if meta["attn_dropout"] > 0 else 0, off))
        return weights, stride, list(off)


class _EncoderFn(torch.autograd.Function):
    """BertEncoder (M.py:344-371): all layers in ONE vb_encoder_fwd / vb_encoder_bwd call over one activation arena."""

    @staticmethod
    def forward(ctx, x, mbias, meta, *params):
        _require_cuda(x, "bert_encoder")
        B, S, H = x.shape
        L = len(params) // 16
        I = params[10].shape[0]
        A = meta["heads"]
        x = x.contiguous()
        with torch.cuda.device(x.device):
            weights, stride, off = meta["plan"].prepare(meta["caches"], params, B, S, H, A, I, mbias, meta)
            arena = torch.empty(L * stride, device=x.device, dtype=torch.uint8)
            plan = meta["plan"]
            _lib.check(_lib.lib().vb_encoder_fwd(plan.descs, L, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(arena.data_ptr()), _stream()),
                       "vb_encoder_fwd")
        n = B * S * H * 2
        outs = tuple(arena[l * stride + off[13]: l * stride + off[13] + n].view(_BF16).view(B, S, H) for l in range(L))
        ctx.meta, ctx.arena, ctx.params, ctx.weights = meta, arena, params, weights
        ctx.shape = (B, S, H, A, I, L)
        ctx.save_for_backward(x, mbias)
        ctx.mark_non_differentiable(*outs[:-1])
        ctx.set_materialize_grads(False)   # or autograd hands backward a 64 MB zero tensor for each of the L - 1 unused outputs
        return outs

    @staticmethod
    def backward(ctx, *douts):
        x, mbias = ctx.saved_tensors
        meta, params = ctx.meta, ctx.params
        B, S, H, A, I, L = ctx.shape
        M = B * S
        dev = x.device
        if douts[-1] is None:   # nothing downstream depends on the encoder output
            return (None,) * (3 + 16 * L)
        dy = douts[-1].to(_BF16).contiguous()
        groups = []
        for l in range(L):
            qw, qb, kw, kb, vw, vb = params[16 * l: 16 * l + 6]
            groups += [(qw, kw, vw), (qb, kb, vb)]
        direct = _grad_targets(params, groups)
        sizes = [3 * H * H, 3 * H, H * H, H, H, H, I * H, I, H * I, H, H, H]
        grads = (_lib.LayerGrads * L)()
        gnames = ("dw_qkv", "db_qkv", "dw_attn_out", "db_attn_out", "dln1_gamma", "dln1_beta",
                  "dw_inter", "db_inter", "dw_out", "db_out", "dln2_gamma", "dln2_beta")
        if direct is not None:
            for l in range(L):
                t = direct[16 * l: 16 * l + 16]
                ptrs = (t[0], t[1], t[6], t[7], t[8], t[9], t[10], t[11], t[12], t[13], t[14], t[15])
                for nme, tt in zip(gnames, ptrs):
                    setattr(grads[l], nme, tt.data_ptr())
            flat = None
        else:
            per = sum(sizes)
            flat = torch.zeros(L * per, device=dev, dtype=torch.float32)
            for l in range(L):
                o = l * per
                for nme, sz in zip(gnames, sizes):
                    setattr(grads[l], nme, flat.data_ptr() + 4 * o)
                    o += sz
        hd = meta["hidden_dropout"] > 0
        with torch.cuda.device(dev):
            w = _scratch(dev, M, H, I, B, A, S, hd)
            sc = _lib.LayerScratch(**{k: _ptr(t) for k, t in w.items()})
            dx = torch.empty(B, S, H, device=dev, dtype=_BF16)
            plan = meta["plan"]
            for l in range(L):  # same step, same dropout streams as the forward
                d = plan.descs[l]
                d.hidden_dropout, d.attn_dropout, d.seed = meta["hidden_dropout"], meta["attn_dropout"], meta["seed"]
                d.layer_index, d.mask_bias = meta["layer_index0"] + l, mbias.data_ptr()
            _lib.check(_lib.lib().vb_encoder_bwd(plan.descs, L, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(ctx.arena.data_ptr()),
                                                 ctypes.c_void_p(dy.data_ptr()), ctypes.c_void_p(dx.data_ptr()), grads, ctypes.byref(sc),
                                                 _stream()), "vb_encoder_bwd")
        ctx.arena = None
        if direct is not None:
            return (dx, None, None) + (None,) * (16 * L)
        out = []
        per = sum(sizes)
        for l in range(L):
            dwqkv, dbqkv, dwo, dbo, dg1, db1, dwi, dbi, dwout, dbout, dg2, db2 = torch.split(flat[l * per: (l + 1) * per], sizes)
            dwq, dwk, dwv = dwqkv.view(3, H, H).unbind(0)
            dbq, dbk, dbv = dbqkv.view(3, H).unbind(0)
            out += [dwq, dbq, dwk, dbk, dwv, dbv, dwo.view(H, H), dbo, dg1, db1, dwi.view(I, H), dbi, dwout.view(H, I), dbout, dg2, db2]
        return (dx, None, None) + tuple(out)


def bert_encoder(x, mbias, meta, params):
    """All layers at once. meta: dict(heads, layer_index0, hidden_dropout, attn_dropout, seed, train, caches=[LayerWeights],
    plan=EncoderPlan); params: 16 tensors per layer in bert_layer order. Returns the tuple of all layer outputs (only
    the last one is differentiable: a caller that needs gradients through intermediate outputs uses bert_layer)."""
    return _EncoderFn.apply(x, mbias, meta, *params)


def bert_layer(x, mbias, meta, params):
    """params: the 16 tensors of one BertLayer in reference order (q.w, q.b, k.w, k.b, v.w, v.b, attention.output
    dense.w/.b, LayerNorm.w/.b, intermediate.dense.w/.b, output.dense.w/.b, LayerNorm.w/.b)."""
    return _LayerFn.apply(x, mbias, meta, *params)


class ProjectionWeights:
    def __init__(self):
        self.key = None
        self.buf = None
        self.bank = None

    def get(self, w, train=False):
        if self.bank is not None:
            return self.buf
        key = (w.data_ptr(), w._version)
        if train or key != self.key:
            with torch.no_grad():
                self.buf = cast_to_bf16(w.detach(), self.buf if self.buf is not None and self.buf.device == w.device else None)
            self.key = key
        return self.buf


class _EmbedFn(torch.autograd.Function):
    """BertEmbeddingsWithVisualEmbedding.forward (reference M.py:1198-1257) through vb_embed_fwd / vb_embed_bwd."""

    @staticmethod
    def forward(ctx, meta, input_ids, token_type_ids, visual_type, feats, word, pos, typ, typ_vis, pos_vis, pw, pb, gamma, beta,
                vis_extra=None):
        _require_cuda(word, "bert_embeddings")
        dev = word.device
        B, T = input_ids.shape
        V = 0 if feats is None else feats.shape[1]
        H = word.shape[1]
        M = B * (T + V)
        ids = input_ids.to(torch.int64).contiguous()
        tt = token_type_ids.to(torch.int64).contiguous()
        if V > 0:
            Dv = feats.shape[2]
            vt = visual_type.to(torch.int64).contiguous()
            f = feats.reshape(B * V, Dv)
            fb = cast_to_bf16(f) if f.dtype == torch.float32 else f.to(_BF16).contiguous()
            wp = meta["cache"].get(pw, train=meta.get("train", False))
            vis_proj = torch.empty(B * V, H, device=dev, dtype=_BF16)
            xb = None if vis_extra is None else vis_extra.detach().reshape(B * V, H).to(_BF16).contiguous()
        else:
            Dv, vt, fb, wp, vis_proj, xb = 0, None, None, None, None, None
        pre = torch.empty(M, H, device=dev, dtype=_BF16)
        mean = torch.empty(M, device=dev, dtype=torch.float32)
        rstd = torch.empty(M, device=dev, dtype=torch.float32)
        y = torch.empty(B, T + V, H, device=dev, dtype=_BF16)
        d = _lib.EmbedDesc(
            batch=B, text_len=T, num_regions=V, hidden=H, visual_dim=Dv, vocab=word.shape[0], max_pos=pos.shape[0],
            n_types=typ.shape[0], eps=1e-12, dropout=meta["dropout"], seed=meta["seed"],
            input_ids=ids.data_ptr(), token_type_ids=tt.data_ptr(), visual_type=_ptr(vt), visual_feats=_ptr(fb),
            w_proj=_ptr(wp), b_proj=_ptr(pb), word=word.data_ptr(), pos=pos.data_ptr(), type=typ.data_ptr(),
            pos_vis=pos_vis.data_ptr(), type_vis=typ_vis.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
            visual_addend=_ptr(xb))
        a = _lib.EmbedActs(vis_proj=_ptr(vis_proj), pre=pre.data_ptr(), mean=mean.data_ptr(), rstd=rstd.data_ptr())
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().vb_embed_fwd(ctypes.byref(d), ctypes.c_void_p(y.data_ptr()), ctypes.byref(a), _stream()),
                       "vb_embed_fwd")
        ctx.meta = meta
        ctx.shape = (B, T, V, H, Dv)
        ctx.feats_need_grad = feats is not None and feats.requires_grad
        ctx.feats_dtype = None if feats is None else feats.dtype
        ctx.feats_shape = None if feats is None else feats.shape
        ctx.extra = None if (vis_extra is None or not vis_extra.requires_grad) else (vis_extra.shape, vis_extra.dtype)
        ctx.params = (word, pos, typ, typ_vis, pos_vis, pw, pb, gamma, beta)
        ctx.save_for_backward(ids, tt, vt, fb, wp, pb, word, pos, typ, typ_vis, pos_vis, gamma, beta, pre, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        ids, tt, vt, fb, wp, pb, word, pos, typ, typ_vis, pos_vis, gamma, beta, pre, mean, rstd = ctx.saved_tensors
        B, T, V, H, Dv = ctx.shape
        meta = ctx.meta
        dev = word.device
        M = B * (T + V)
        dy = dy.to(_BF16).contiguous()
        f32 = torch.float32
        direct = _grad_targets(ctx.params) if V > 0 else None
        if direct is not None:
            dword, dpos, dtyp, dtyp_vis, dpos_vis, dpw, dpb, dgamma, dbeta = direct
        else:
            dword = torch.zeros_like(word, dtype=f32)
            dpos = torch.zeros_like(pos, dtype=f32)
            dtyp = torch.zeros_like(typ, dtype=f32)
            dtyp_vis = torch.zeros_like(typ_vis, dtype=f32)
            dpos_vis = torch.zeros_like(pos_vis, dtype=f32)
            dgamma = torch.zeros_like(gamma, dtype=f32)
            dbeta = torch.zeros_like(beta, dtype=f32)
            dpw = torch.zeros(H, Dv, device=dev, dtype=f32) if V > 0 else None
            dpb = torch.zeros(H, device=dev, dtype=f32) if V > 0 else None
        d_pre = torch.empty(M, H, device=dev, dtype=_BF16)
        if V > 0:
            d_vis = torch.empty(B * V, H, device=dev, dtype=_BF16)
            d_feats = torch.empty(B * V, Dv, device=dev, dtype=_BF16) if ctx.feats_need_grad else None
        else:
            d_vis = d_feats = None
        d = _lib.EmbedDesc(
            batch=B, text_len=T, num_regions=V, hidden=H, visual_dim=Dv, vocab=word.shape[0], max_pos=pos.shape[0],
            n_types=typ.shape[0], eps=1e-12, dropout=meta["dropout"], seed=meta["seed"],
            input_ids=ids.data_ptr(), token_type_ids=tt.data_ptr(), visual_type=_ptr(vt), visual_feats=_ptr(fb),
            w_proj=_ptr(wp), b_proj=_ptr(pb), word=word.data_ptr(), pos=pos.data_ptr(), type=typ.data_ptr(),
            pos_vis=pos_vis.data_ptr(), type_vis=typ_vis.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr())
        a = _lib.EmbedActs(vis_proj=0, pre=pre.data_ptr(), mean=mean.data_ptr(), rstd=rstd.data_ptr())
        g = _lib.EmbedGrads(
            dword=dword.data_ptr(), dpos=dpos.data_ptr(), dtype=dtyp.data_ptr(), dpos_vis=dpos_vis.data_ptr(),
            dtype_vis=dtyp_vis.data_ptr(), dw_proj=_ptr(dpw), db_proj=_ptr(dpb), dgamma=dgamma.data_ptr(),
            dbeta=dbeta.data_ptr(), d_pre=d_pre.data_ptr(), d_vis=_ptr(d_vis), d_feats=_ptr(d_feats))
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().vb_embed_bwd(ctypes.byref(d), ctypes.byref(a), ctypes.c_void_p(dy.data_ptr()),
                                               ctypes.byref(g), _stream()), "vb_embed_bwd")
        dfe = None
        if d_feats is not None:
            dfe = d_feats.view(ctx.feats_shape).to(ctx.feats_dtype)
        # the gradient of anything added to the projected region rows is d_vis itself
        dex = None if ctx.extra is None else d_vis.view(ctx.extra[0]).to(ctx.extra[1])
        if direct is not None:
            return (None, None, None, None, dfe) + (None,) * 9 + (dex,)
        return (None, None, None, None, dfe, dword, dpos, dtyp, dtyp_vis, dpos_vis, dpw, dpb, dgamma, dbeta, dex)


def bert_embeddings(meta, input_ids, token_type_ids, visual_type, feats, word, pos, typ, typ_vis, pos_vis, pw, pb, gamma, beta,
                    vis_extra=None):
    """vis_extra: optional [B, V, H] term added to the projected region rows before the LayerNorm (differentiable)."""
    return _EmbedFn.apply(meta, input_ids, token_type_ids, visual_type, feats, word, pos, typ, typ_vis, pos_vis, pw, pb,
                          gamma, beta, vis_extra)


# --------------------------------------------------------------------------------------------
# masked-LM head on the library's kernels: decoder GEMMs on tcgen05, fused cross-entropy
# --------------------------------------------------------------------------------------------
def _gemm(dev, **kw):
    a = _lib.GemmArgs()
    for k, v in kw.items():
        setattr(a, k, v)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().vb_gemm(ctypes.byref(a), _stream()), "vb_gemm")


class DecoderWeights:
    """bf16 copy of the tied word-embedding matrix [vocab, H] and the fp32 bias padded to a multiple of 16 columns
    (padding columns get -30000 so they vanish in the softmax), refreshed when the masters change."""

    def __init__(self):
        self.key = None
        self.table = None
        self.bias = None
        self.bank = None

    def alloc(self, E):
        V = E.shape[0]
        Vp = (V + 15) // 16 * 16
        if self.table is None or self.table.device != E.device or self.table.shape[0] != Vp:
            self.table = torch.zeros(Vp, E.shape[1], device=E.device, dtype=_BF16)
            self.bias = torch.full((Vp,), -30000.0, device=E.device, dtype=torch.float32)
        return self.table, self.bias

    def get(self, E, bias, train=False):
        if self.bank is not None:
            return self.table, self.bias
        key = (E.data_ptr(), E._version, bias.data_ptr(), bias._version)
        if train or key != self.key:
            V = E.shape[0]
            Vp = (V + 15) // 16 * 16
            with torch.no_grad():
                # [Vp, H] with zero rows V..Vp-1: the GEMMs address all Vp rows of the table
                if self.table is None or self.table.device != E.device or self.table.shape[0] != Vp:
                    self.table = torch.zeros(Vp, E.shape[1], device=E.device, dtype=_BF16)
                cast_to_bf16(E.detach(), self.table[:V])
                if self.bias is None or self.bias.device != E.device or self.bias.numel() != Vp:
                    self.bias = torch.full((Vp,), -30000.0, device=E.device, dtype=torch.float32)
                self.bias[:V].copy_(bias.detach())
            self.key = key
        return self.table, self.bias


class _MlmDecoderFn(torch.autograd.Function):
    """logits[n, Vp] = t[n, H] @ E[V, H]^T + bias (BertLMPredictionHead decoder, reference M.py:403-421) — forward,
    input-gradient and weight-gradient GEMMs all on gemm_tcgen05_kernel; the weight gradient accumulates straight
    into the (tied) word-embedding gradient when that buffer exists."""

    @staticmethod
    def forward(ctx, t, E, bias, cache, train=False):
        _require_cuda(t, "mlm_decoder")
        n, H = t.shape
        V = E.shape[0]
        Vp = (V + 15) // 16 * 16
        table, bias_p = cache.get(E, bias, train=train)
        t = t.to(_BF16).contiguous()
        logits = torch.empty(n, Vp, device=t.device, dtype=_BF16)
        _gemm(t.device, A=t.data_ptr(), lda=H, B=table.data_ptr(), ldb=H, M=n, N=Vp, K=H, D=logits.data_ptr(), ldd=Vp,
              bias=bias_p.data_ptr())
        ctx.save_for_backward(t, table)
        ctx.E, ctx.bias, ctx.V = E, bias, V
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        t, table = ctx.saved_tensors
        E, bias, V = ctx.E, ctx.bias, ctx.V
        n, H = t.shape
        Vp = dlogits.shape[1]
        dlogits = dlogits.contiguous()
        dt = torch.empty(n, H, device=t.device, dtype=_BF16)
        _gemm(t.device, A=dlogits.data_ptr(), lda=Vp, B=table.data_ptr(), ldb=H, b_mn_major=1, M=n, N=H, K=Vp, D=dt.data_ptr(), ldd=H)
        direct = _grad_targets((E, bias))
        if direct is not None:
            dE, db = direct
            db_pad = torch.zeros(Vp, device=t.device, dtype=torch.float32)
        else:
            dE = torch.zeros(V, H, device=t.device, dtype=torch.float32)
            db_pad = torch.zeros(Vp, device=t.device, dtype=torch.float32)
        _gemm(t.device, A=dlogits.data_ptr(), lda=Vp, a_mn_major=1, B=t.data_ptr(), ldb=H, b_mn_major=1, M=V, N=H, K=n,
              D=dE.data_ptr(), ldd=H, d_fp32=1, splits=1)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().vb_colsum_bf16(ctypes.c_void_p(dlogits.data_ptr()), ctypes.c_int64(Vp), ctypes.c_void_p(db_pad.data_ptr()),
                                                 n, Vp, _stream()), "vb_colsum_bf16")
        if direct is not None:
            db.add_(db_pad[:V])
            return dt, None, None, None, None
        return dt, dE, db_pad[:V].clone(), None, None


def mlm_decoder(t, E, bias, cache, train=False):
    return _MlmDecoderFn.apply(t, E, bias, cache, train)


class _CrossEntropyFn(torch.autograd.Function):
    """mean over rows of (logsumexp(logits[:, :V]) - logits[row, label]); the backward overwrites the logits with
    their gradient in place (vb_cross_entropy_bwd)."""

    @staticmethod
    def forward(ctx, logits, labels, V):
        n, Vp = logits.shape
        labels = labels.to(torch.int64).contiguous()
        lse = torch.empty(n, device=logits.device, dtype=torch.float32)
        rows = torch.empty(n, device=logits.device, dtype=torch.float32)
        with torch.cuda.device(logits.device):
            _lib.check(_lib.lib().vb_cross_entropy_fwd(ctypes.c_void_p(logits.data_ptr()), ctypes.c_int64(Vp),
                                                       ctypes.c_void_p(labels.data_ptr()), n, V, ctypes.c_void_p(lse.data_ptr()),
                                                       ctypes.c_void_p(rows.data_ptr()), _stream()), "vb_cross_entropy_fwd")
        ctx.logits, ctx.labels, ctx.lse, ctx.V = logits, labels, lse, V
        return rows.mean()

    @staticmethod
    def backward(ctx, g):
        logits, labels, lse, V = ctx.logits, ctx.labels, ctx.lse, ctx.V
        n, Vp = logits.shape
        scale = (g.float() / n).reshape(1).contiguous()
        with torch.cuda.device(logits.device):
            _lib.check(_lib.lib().vb_cross_entropy_bwd(ctypes.c_void_p(logits.data_ptr()), ctypes.c_int64(Vp),
                                                       ctypes.c_void_p(labels.data_ptr()), n, V, Vp, ctypes.c_void_p(lse.data_ptr()),
                                                       ctypes.c_void_p(scale.data_ptr()), _stream()), "vb_cross_entropy_bwd")
        ctx.logits = None
        return logits, None, None


def cross_entropy_rows(logits, labels, V):
    return _CrossEntropyFn.apply(logits, labels, V)
