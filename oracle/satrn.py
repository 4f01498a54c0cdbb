"""CPU restatement of the reference's formula-recognition hot path (the ORACLE).

TEST INFRASTRUCTURE ONLY -- the checker, never the thing shipped or measured.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``p4-fr-sorry-math-but-love-you_b200``) never does and has no CPU fallback.

It is a *functional* restatement in plain fp32 torch-CPU ops (F.conv2d,
F.linear, softmax ...) that works straight on a reference-layout ``state_dict``
(SURVEY.md 8b / App. A.2); every function cites the reference lines it follows.
All citations are into /root/reference/networks/EfficientSATRN.py unless a
file name is given.

Parity status:
  * decoder, 2-D PE, SATRN encoder layers, greedy loop, best-first "beam"
    search, teacher-forced branch: pinned against the reference's own modules
    run through ``oracle/ref_shim.py`` (tests/test_oracle_vs_reference.py, and
    the committed fixtures under tests/golden/ made by oracle/make_golden.py).
  * conv trunk blocks: the arithmetic lives in third-party ``timm==0.4.9``
    (requirements.txt:15; model ``tf_efficientnetv2_s_in21ft1k``; call sites
    :66,:74,:84), which is absent from /root/reference and from this image.
    Its published block definitions are restated here.  The
    reference holds no test or golden vector for it -> PARITY UNPINNED at the
    timm boundary (only the parameter count / state_dict layout are checked).
"""
from __future__ import annotations

import heapq
import itertools
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SOS_ID, EOS_ID, PAD_ID = 0, 1, 2  # utils/data_utils.py:6-9,36-41

# (kind, repeats, kernel, stride, expand, out_ch, se_ratio)  -- SURVEY App. A.1
EFFNETV2_S_ARCH = [
    ("cn", 2, 3, 1, 1, 24, 0.0),
    ("er", 4, 3, 2, 4, 48, 0.0),
    ("er", 4, 3, 2, 4, 64, 0.0),
    ("ir", 6, 3, 2, 4, 128, 0.25),
    ("ir", 9, 3, 1, 6, 160, 0.25),
    ("ir", 15, 3, 2, 6, 256, 0.25),
]


@dataclass
class ModelSpec:
    """Dimensions the constructors read from FLAGS (:667-688) and the vocab."""

    network: str = "EfficientSATRN"
    height: int = 128
    width: int = 256
    in_ch: int = 1
    enc_hidden: int = 512
    enc_filter: int = 512
    enc_layers: int = 2
    enc_heads: int = 8
    dec_src: int = 512
    dec_hidden: int = 256
    dec_filter: int = 1024
    dec_layers: int = 3
    dec_heads: int = 8
    num_classes: int = 245
    pe1d_len: int = 500  # :401

    @property
    def feat_hw(self) -> Tuple[int, int]:
        if self.network == "LiteSATRN":
            return self.height // 16, self.width // 16  # LiteSATRN.py:281-283
        return self.height // 32, self.width // 32  # :300


def efficient_satrn_spec(**kw) -> ModelSpec:
    return ModelSpec(**kw)


# ---------------------------------------------------------------------------
# Trunk block table (shared by the shape list, the forward and the product's
# weight packer tests)
# ---------------------------------------------------------------------------
def trunk_blocks(stem_chs: int = 24):
    """Yield (prefix, kind, cin, cout, k, stride, expand, se_reduce) per block."""
    cin = stem_chs
    out = []
    for si, (kind, reps, k, stride, expand, cout, se) in enumerate(EFFNETV2_S_ARCH):
        for r in range(reps):
            s = stride if r == 0 else 1
            out.append(("eff_block.%d.%d" % (si, r), kind, cin, cout, k, s, expand, int(cin * se)))
            cin = cout
    return out


def _bn_entries(shapes, p, c):
    shapes[p + ".weight"] = (c,)
    shapes[p + ".bias"] = (c,)
    shapes[p + ".running_mean"] = (c,)
    shapes[p + ".running_var"] = (c,)
    shapes[p + ".num_batches_tracked"] = ()


def _mha_entries(shapes, p, qc, kc, d):
    for n, cin in (("q_linear", qc), ("k_linear", kc), ("v_linear", kc)):
        shapes[p + "." + n + ".weight"] = (d, cin)
        shapes[p + "." + n + ".bias"] = (d,)
    shapes[p + ".out_linear.weight"] = (qc, d)
    shapes[p + ".out_linear.bias"] = (qc,)


def param_shapes(spec: ModelSpec) -> Dict[str, tuple]:
    """state_dict layout, in the reference's registration order (SURVEY 8b,
    App. A.2; module definitions :63-79, :90-109, :231-257, :349-372,
    :429-461)."""
    s: Dict[str, tuple] = {}
    e = "encoder.shallow_cnn."
    if spec.network == "EfficientSATRN":
        s[e + "conv_stem.weight"] = (24, spec.in_ch, 3, 3)
        _bn_entries(s, e + "bn1", 24)
        for pfx, kind, cin, cout, k, stride, expand, se_r in trunk_blocks():
            p = e + pfx
            mid = cin * expand
            if kind == "cn":
                s[p + ".conv.weight"] = (cout, cin, k, k)
                _bn_entries(s, p + ".bn1", cout)
            elif kind == "er":
                s[p + ".conv_exp.weight"] = (mid, cin, k, k)
                _bn_entries(s, p + ".bn1", mid)
                s[p + ".conv_pwl.weight"] = (cout, mid, 1, 1)
                _bn_entries(s, p + ".bn2", cout)
            else:
                s[p + ".conv_pw.weight"] = (mid, cin, 1, 1)
                _bn_entries(s, p + ".bn1", mid)
                s[p + ".conv_dw.weight"] = (mid, 1, k, k)
                _bn_entries(s, p + ".bn2", mid)
                s[p + ".se.conv_reduce.weight"] = (se_r, mid, 1, 1)
                s[p + ".se.conv_reduce.bias"] = (se_r,)
                s[p + ".se.conv_expand.weight"] = (mid, se_r, 1, 1)
                s[p + ".se.conv_expand.bias"] = (mid,)
                s[p + ".conv_pwl.weight"] = (cout, mid, 1, 1)
                _bn_entries(s, p + ".bn3", cout)
        s[e + "conv_last.weight"] = (spec.enc_hidden, 256, 1, 1)
        _bn_entries(s, e + "bn2", spec.enc_hidden)
    elif spec.network == "LiteSATRN":
        # LiteSATRN.py:21-70  ShallowCNN: 4 x (conv3x3 p1, BN, ReLU, maxpool2)
        h = spec.enc_hidden
        chans = [(spec.in_ch, h // 2), (h // 2, h), (h, h), (h, h)]
        for i, (ci, co) in enumerate(chans):
            s[e + "conv%d.weight" % i] = (co, ci, 3, 3)
            _bn_entries(s, e + "batch_norm%d" % i, co)
    else:
        raise NotImplementedError(spec.network)

    pe = "encoder.positional_encoding."
    H = spec.enc_hidden
    s[pe + "dense0.weight"] = (H // 2, H)
    s[pe + "dense0.bias"] = (H // 2,)
    s[pe + "dense1.weight"] = (2 * H, H // 2)
    s[pe + "dense1.bias"] = (2 * H,)
    for i in range(spec.enc_layers):
        p = "encoder.attention_layers.%d." % i
        s[p + "norm.weight"] = (H,)
        s[p + "norm.bias"] = (H,)
        _mha_entries(s, p + "attention_layer", H, H, H)
        s[p + "conv0.weight"] = (spec.enc_filter, H, 1, 1)
        _bn_entries(s, p + "norm0", spec.enc_filter)
        s[p + "depthwise.weight"] = (spec.enc_filter, 1, 3, 3)
        s[p + "depthwise.bias"] = (spec.enc_filter,)
        _bn_entries(s, p + "depthwise_norm", spec.enc_filter)
        s[p + "conv1.weight"] = (H, spec.enc_filter, 1, 1)
        _bn_entries(s, p + "norm1", H)

    D = spec.dec_hidden
    s["decoder.embedding.weight"] = (spec.num_classes + 1, D)
    for i in range(spec.dec_layers):
        p = "decoder.attention_layers.%d." % i
        _mha_entries(s, p + "self_attention_layer", D, D, D)
        s[p + "self_attention_norm.weight"] = (D,)
        s[p + "self_attention_norm.bias"] = (D,)
        _mha_entries(s, p + "attention_layer", D, spec.dec_src, D)
        s[p + "attention_norm.weight"] = (D,)
        s[p + "attention_norm.bias"] = (D,)
        s[p + "feedforward_layer.linear0.weight"] = (spec.dec_filter, D)
        s[p + "feedforward_layer.linear0.bias"] = (spec.dec_filter,)
        s[p + "feedforward_layer.linear1.weight"] = (D, spec.dec_filter)
        s[p + "feedforward_layer.linear1.bias"] = (D,)
        s[p + "feedforward_norm.weight"] = (D,)
        s[p + "feedforward_norm.bias"] = (D,)
    s["decoder.generator.weight"] = (spec.num_classes, D)
    s["decoder.generator.bias"] = (spec.num_classes,)
    return s


# ---------------------------------------------------------------------------
# Positional tables (not in the state_dict: built on the host by the reference)
# ---------------------------------------------------------------------------
def pe2d_table(length: int, hidden: int) -> torch.Tensor:
    """:111-127  sin||cos concatenated, hidden//2 timescales."""
    position = torch.arange(length).float()
    nts = hidden // 2
    log_inc = math.log(1.0e4 / 1.0) / (torch.FloatTensor([nts]) - 1)
    inv = 1.0 * torch.exp(torch.FloatTensor(torch.arange(nts) * -log_inc))
    scaled = position.unsqueeze(1) * inv.unsqueeze(0)
    return torch.cat((torch.sin(scaled), torch.cos(scaled)), dim=1)  # [length, hidden]


def pe1d_table(channels: int, max_len: int = 500) -> torch.Tensor:
    """:408-418  interleaved sin (even) / cos (odd)."""
    pos = torch.arange(max_len).float().unsqueeze(1)
    i = torch.arange(channels).float().unsqueeze(0)
    rates = 1 / torch.pow(10000, (2 * (i // 2)) / channels)
    pe = pos * rates
    pe[:, 0::2] = torch.sin(pe[:, 0::2])
    pe[:, 1::2] = torch.cos(pe[:, 1::2])
    return pe  # [max_len, channels]


# ---------------------------------------------------------------------------
# Encoder
# ---------------------------------------------------------------------------
class _Calib:
    """BN calibration recorder used only by ``synth_state_dict``."""

    def __init__(self, sd):
        self.sd = sd


def _quantize_mantissa(x: torch.Tensor, bits: int = 10) -> torch.Tensor:
    m, e = torch.frexp(x.double())
    m = torch.round(m * (1 << bits)) / (1 << bits)
    return torch.ldexp(m, e)


class TrainMode:
    """Marker passed in the ``calib`` slot: BatchNorm in TRAIN mode (batch statistics, running statistics updated in
    place with ``momentum`` exactly as nn.BatchNorm2d does) -- the mode the reference trains in
    (train_modules/train_single_opt.py:54 ``model.train()``)."""

    def __init__(self, momentum: float = 0.1):
        self.momentum = momentum


def _bn(x, sd, p, eps, calib: Optional[_Calib] = None):
    if isinstance(calib, TrainMode):
        if p + ".num_batches_tracked" in sd:
            sd[p + ".num_batches_tracked"] += 1
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            True, calib.momentum, eps)
    if calib is not None:
        # own synthetic-checkpoint procedure: running stats := batch stats of the
        # eval-mode forward so far, in fp64, rounded to 10 mantissa bits so the
        # generated checkpoint is bit-reproducible on any host.
        xd = x.double()
        mean = xd.mean(dim=(0, 2, 3))
        var = xd.var(dim=(0, 2, 3), unbiased=False).clamp_min(1e-6)
        sd[p + ".running_mean"] = _quantize_mantissa(mean).to(x.dtype)
        sd[p + ".running_var"] = _quantize_mantissa(var).to(x.dtype)
    return F.batch_norm(
        x, sd[p + ".running_mean"].to(x.dtype), sd[p + ".running_var"].to(x.dtype),
        sd[p + ".weight"].to(x.dtype), sd[p + ".bias"].to(x.dtype), False, 0.0, eps)


def _conv_same(x, w, stride, groups=1, bias=None):
    """timm 0.4.9 layers/padding.py + conv2d_same.py: static symmetric pad for
    stride 1, TF-style dynamic asymmetric pad for stride 2."""
    k = w.shape[-1]
    if stride == 1:
        return F.conv2d(x, w, bias, 1, (k - 1) // 2, 1, groups)
    ih, iw = x.shape[-2:]
    ph = max((math.ceil(ih / stride) - 1) * stride + k - ih, 0)
    pw = max((math.ceil(iw / stride) - 1) * stride + k - iw, 0)
    x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2])
    return F.conv2d(x, w, bias, stride, 0, 1, groups)


def trunk_forward(sd, spec: ModelSpec, x, calib=None, taps=None):
    """EfficientNet.forward :81-87 with timm blocks restated (App. A.1)."""
    e = "encoder.shallow_cnn."
    W = lambda n: sd[n].to(x.dtype)
    if spec.network == "LiteSATRN":
        for i in range(4):  # LiteSATRN.py:53-70
            x = F.conv2d(x, W(e + "conv%d.weight" % i), None, 1, 1)
            x = F.relu(_bn(x, sd, e + "batch_norm%d" % i, 1e-5, calib))
            x = F.max_pool2d(x, 2, 2)
        return x
    x = F.conv2d(x, W(e + "conv_stem.weight"), None, 2, 0)  # :67-69 padding 0
    x = F.silu(_bn(x, sd, e + "bn1", 1e-3, calib))
    if taps is not None:
        taps["stem"] = x
    for pfx, kind, cin, cout, k, stride, expand, se_r in trunk_blocks():
        p = e + pfx
        res = x if (cin == cout and stride == 1) else None
        if kind == "cn":
            y = F.silu(_bn(_conv_same(x, W(p + ".conv.weight"), stride), sd, p + ".bn1", 1e-3, calib))
        elif kind == "er":
            y = F.silu(_bn(_conv_same(x, W(p + ".conv_exp.weight"), stride), sd, p + ".bn1", 1e-3, calib))
            y = _bn(F.conv2d(y, W(p + ".conv_pwl.weight")), sd, p + ".bn2", 1e-3, calib)
        else:
            y = F.silu(_bn(F.conv2d(x, W(p + ".conv_pw.weight")), sd, p + ".bn1", 1e-3, calib))
            y = _conv_same(y, W(p + ".conv_dw.weight"), stride, groups=y.shape[1])
            y = F.silu(_bn(y, sd, p + ".bn2", 1e-3, calib))
            s = y.mean((2, 3), keepdim=True)
            s = F.silu(F.conv2d(s, W(p + ".se.conv_reduce.weight"), W(p + ".se.conv_reduce.bias")))
            s = F.conv2d(s, W(p + ".se.conv_expand.weight"), W(p + ".se.conv_expand.bias"))
            y = y * torch.sigmoid(s)
            y = _bn(F.conv2d(y, W(p + ".conv_pwl.weight")), sd, p + ".bn3", 1e-3, calib)
        x = y + res if res is not None else y
        if taps is not None:
            taps[pfx] = x
    x = F.conv2d(x, W(e + "conv_last.weight"))  # :75-79
    x = F.silu(_bn(x, sd, e + "bn2", 1e-5, calib))
    return x


def pe2d_forward(sd, spec: ModelSpec, x):
    """PositionalEncoding.forward :135-154 (dropout = identity in eval)."""
    b, c, h, w = x.shape
    p = "encoder.positional_encoding."
    W = lambda n: sd[n].to(x.dtype)
    h_enc = pe2d_table(h, c).to(x).unsqueeze(1).unsqueeze(0)  # [1,h,1,c]  (tables are host-built; .to(x): device + dtype)
    w_enc = pe2d_table(w, c).to(x).unsqueeze(0).unsqueeze(0)  # [1,1,w,c]
    g = torch.mean(x, [2, 3])
    g = F.relu(F.linear(g, W(p + "dense0.weight"), W(p + "dense0.bias")))
    g = torch.sigmoid(F.linear(g, W(p + "dense1.weight"), W(p + "dense1.bias")))
    g = torch.reshape(g, [-1, 2, 1, c])
    pos = g[:, 0:1, :, :] * h_enc + g[:, 1:2, :, :] * w_enc  # [b,h,w,c]
    return pos.permute(0, 3, 1, 2) + x


def mha(sd, p, q_in, k_in, v_in, heads, mask=None):
    """MultiHeadAttention.forward :198-228 + ScaledDotProductAttention :164-172.
    temperature = sqrt(heads*head_dim) and is a DIVISION (:166,:187-189)."""
    W = lambda n: sd[p + n].to(q_in.dtype)
    b, ql, kl = q_in.size(0), q_in.size(1), k_in.size(1)
    d = sd[p + ".q_linear.weight"].shape[0]
    hd = d // heads
    q = F.linear(q_in, W(".q_linear.weight"), W(".q_linear.bias")).view(b, ql, heads, hd).transpose(1, 2)
    k = F.linear(k_in, W(".k_linear.weight"), W(".k_linear.bias")).view(b, kl, heads, hd).transpose(1, 2)
    v = F.linear(v_in, W(".v_linear.weight"), W(".v_linear.bias")).view(b, kl, heads, hd).transpose(1, 2)
    attn = torch.matmul(q, k.transpose(2, 3)) / (float(d) ** 0.5)
    if mask is not None:
        attn = attn.masked_fill(mask.unsqueeze(1), float("-inf"))
    attn = torch.softmax(attn, dim=-1)
    out = torch.matmul(attn, v).transpose(1, 2).contiguous().view(b, ql, d)
    return F.linear(out, W(".out_linear.weight"), W(".out_linear.bias"))


def encoder_layer_forward(sd, spec: ModelSpec, i: int, x, calib=None):
    """EncoderLayer.forward :259-281, including the raw-reshape un-flatten
    (:269) and the twice-used LayerNorm (:265,:268)."""
    p = "encoder.attention_layers.%d." % i
    W = lambda n: sd[p + n].to(x.dtype)
    b, c, h, w = x.shape
    flat = x.view(b, c, h * w).transpose(1, 2)
    y = F.layer_norm(flat, (c,), W("norm.weight"), W("norm.bias"), 1e-5)
    y = mha(sd, p + "attention_layer", y, y, y, spec.enc_heads)
    y = F.layer_norm(y + flat, (c,), W("norm.weight"), W("norm.bias"), 1e-5)
    y = y.reshape(-1, c, h, w)  # reinterpretation, NOT a transpose back
    y = F.relu(_bn(F.conv2d(y, W("conv0.weight")), sd, p + "norm0", 1e-5, calib))
    y = F.conv2d(y, W("depthwise.weight"), W("depthwise.bias"), 1, 1, 1, y.shape[1])
    y = F.relu(_bn(y, sd, p + "depthwise_norm", 1e-5, calib))
    y = F.relu(_bn(F.conv2d(y, W("conv1.weight")), sd, p + "norm1", 1e-5, calib))
    return y + x


def encoder_forward(sd, spec: ModelSpec, images, calib=None, taps=None):
    """SATRNEncoder.forward :311-323 -> src [B, h*w, C] (contiguous here)."""
    x = trunk_forward(sd, spec, images, calib, taps)
    if taps is not None:
        taps["trunk"] = x
    x = pe2d_forward(sd, spec, x)
    if taps is not None:
        taps["pe2d"] = x
    for i in range(spec.enc_layers):
        x = encoder_layer_forward(sd, spec, i, x, calib)
        if taps is not None:
            taps["enc_layer%d" % i] = x
    b, c, h, w = x.shape
    return x.view(b, c, h * w).transpose(1, 2).contiguous()


# ---------------------------------------------------------------------------
# Decoder
# ---------------------------------------------------------------------------
def _embed(sd, tokens):
    """text_embedding :480-483  Embedding * sqrt(hidden)."""
    e = F.embedding(tokens, sd["decoder.embedding.weight"])
    return e * math.sqrt(e.size(-1))


def _ffn(sd, p, x):
    """Feedforward.forward :339-346 -- ReLU after BOTH linears."""
    x = F.relu(F.linear(x, sd[p + ".linear0.weight"], sd[p + ".linear0.bias"]))
    return F.relu(F.linear(x, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"]))


def _ln(sd, p, x):
    return F.layer_norm(x, (x.size(-1),), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def decoder_layer_as_written(sd, spec, i, tgt, tgt_prev, src, tgt_mask):
    """TransformerDecoderLayer.forward :374-397, op for op."""
    p = "decoder.attention_layers.%d." % i
    kv = tgt if tgt_prev is None else torch.cat([tgt_prev, tgt], 1)
    att = mha(sd, p + "self_attention_layer", tgt, kv, kv, spec.dec_heads, tgt_mask)
    out = _ln(sd, p + "self_attention_norm", att + tgt)
    att = mha(sd, p + "attention_layer", out, src, src, spec.dec_heads)
    out = _ln(sd, p + "attention_norm", att + out)
    ff = _ffn(sd, p + "feedforward_layer", out)
    return _ln(sd, p + "feedforward_norm", ff + out)


def decode_greedy_as_written(sd, spec: ModelSpec, src, steps: int):
    """SATRNDecoder.forward inference branch :528-566, op for op (re-projects
    K/V of the whole history and of ``src`` every step, as the reference does).
    Returns logits [B, steps, V]."""
    b = src.size(0)
    pe = pe1d_table(spec.dec_hidden, spec.pe1d_len).to(src.device)
    target = torch.full((b,), SOS_ID, dtype=torch.long, device=src.device)
    features = [None] * spec.dec_layers
    out = []
    for t in range(steps):
        tgt = _embed(sd, target.unsqueeze(1)) + pe[t].unsqueeze(0).unsqueeze(1)
        for l in range(spec.dec_layers):
            tgt = decoder_layer_as_written(sd, spec, l, tgt, features[l], src, None)
            features[l] = tgt if features[l] is None else torch.cat([features[l], tgt], 1)
        logit = F.linear(tgt, sd["decoder.generator.weight"], sd["decoder.generator.bias"])
        target = torch.argmax(logit[:, -1, :], dim=-1)
        out.append(logit)
    return torch.cat(out, dim=1)


class DecoderState:
    """Per-image incremental state of the exact recurrence:
    per layer, K/V rows of the cached layer OUTPUTS, plus cross K/V of src."""

    def __init__(self, sd, spec: ModelSpec, src):
        self.sd, self.spec = sd, spec
        self.b = src.size(0)
        self.t = 0
        L, D, H = spec.dec_layers, spec.dec_hidden, spec.dec_heads
        self.k_self = [torch.empty(self.b, 0, D) for _ in range(L)]
        self.v_self = [torch.empty(self.b, 0, D) for _ in range(L)]
        self.k_cross, self.v_cross = [], []
        for l in range(L):
            p = "decoder.attention_layers.%d.attention_layer" % l
            self.k_cross.append(F.linear(src, sd[p + ".k_linear.weight"], sd[p + ".k_linear.bias"]))
            self.v_cross.append(F.linear(src, sd[p + ".v_linear.weight"], sd[p + ".v_linear.bias"]))
        self.pe = pe1d_table(D, spec.pe1d_len)


def _attend(q, k, v, heads, temperature):
    b, ql, d = q.shape
    hd = d // heads
    qh = q.view(b, ql, heads, hd).transpose(1, 2)
    kh = k.view(b, -1, heads, hd).transpose(1, 2)
    vh = v.view(b, -1, heads, hd).transpose(1, 2)
    a = torch.softmax(torch.matmul(qh, kh.transpose(2, 3)) / temperature, dim=-1)
    return torch.matmul(a, vh).transpose(1, 2).contiguous().view(b, ql, d)


def decode_step_cached(st: DecoderState, tokens: torch.Tensor, position: Optional[int] = None,
                       append: bool = True):
    """One decode step (App. A.4): same mathematics as
    ``decoder_layer_as_written`` with the K/V rows of earlier layer outputs
    cached instead of re-projected.  tokens [B] -> logits [B, V]."""
    sd, spec = st.sd, st.spec
    t = st.t if position is None else position
    temp = float(spec.dec_hidden) ** 0.5
    x = _embed(sd, tokens.unsqueeze(1)) + st.pe[t].unsqueeze(0).unsqueeze(1)  # [B,1,D]
    new_kv = []
    for l in range(spec.dec_layers):
        p = "decoder.attention_layers.%d." % l
        sa = p + "self_attention_layer"
        lin = lambda n, z: F.linear(z, sd[sa + n + ".weight"], sd[sa + n + ".bias"])
        q = lin(".q_linear", x)
        k = torch.cat([st.k_self[l], lin(".k_linear", x)], 1)  # current layer INPUT is the last key
        v = torch.cat([st.v_self[l], lin(".v_linear", x)], 1)
        a = _attend(q, k, v, spec.dec_heads, temp)
        u = _ln(sd, p + "self_attention_norm", lin(".out_linear", a) + x)
        ca = p + "attention_layer"
        q2 = F.linear(u, sd[ca + ".q_linear.weight"], sd[ca + ".q_linear.bias"])
        c = _attend(q2, st.k_cross[l], st.v_cross[l], spec.dec_heads, temp)
        w = _ln(sd, p + "attention_norm",
                F.linear(c, sd[ca + ".out_linear.weight"], sd[ca + ".out_linear.bias"]) + u)
        y = _ln(sd, p + "feedforward_norm", _ffn(sd, p + "feedforward_layer", w) + w)
        new_kv.append((lin(".k_linear", y), lin(".v_linear", y)))  # cache rows of the OUTPUT
        x = y
    if append:
        for l, (kk, vv) in enumerate(new_kv):
            st.k_self[l] = torch.cat([st.k_self[l], kk], 1)
            st.v_self[l] = torch.cat([st.v_self[l], vv], 1)
        st.t += 1
    logits = F.linear(x[:, 0], sd["decoder.generator.weight"], sd["decoder.generator.bias"])
    return logits, new_kv


def decode_greedy(sd, spec: ModelSpec, src, steps: int, forced_tokens: Optional[torch.Tensor] = None):
    """Greedy loop :539-558 on the cached recurrence.  ``forced_tokens``
    [B, steps] (optional) feeds given tokens instead of the argmax (forced
    decoding).  Returns (logits [B,steps,V], tokens [B,steps])."""
    st = DecoderState(sd, spec, src)
    tok = torch.full((src.size(0),), SOS_ID, dtype=torch.long)
    logits, toks = [], []
    for t in range(steps):
        lg, _ = decode_step_cached(st, tok)
        nxt = torch.argmax(lg, dim=-1)
        logits.append(lg)
        toks.append(nxt)
        tok = forced_tokens[:, t] if forced_tokens is not None else nxt
    return torch.stack(logits, 1), torch.stack(toks, 1)


def teacher_forced(sd, spec: ModelSpec, src, text):
    """SATRNDecoder.forward train branch :488-495 with masks :469-478.
    text [B, L] (= expected[:, :-1]) -> logits [B, L, V]."""
    L = text.size(1)
    pe = pe1d_table(spec.dec_hidden, spec.pe1d_len)
    tgt = _embed(sd, text) + pe[:L].unsqueeze(0)
    pad = text == PAD_ID
    pad[:, 0] = False
    mask = pad.unsqueeze(1) | torch.triu(torch.ones(L, L), diagonal=1).bool().unsqueeze(0)
    for l in range(spec.dec_layers):
        tgt = decoder_layer_as_written(sd, spec, l, tgt, None, src, mask)
    return F.linear(tgt, sd["decoder.generator.weight"], sd["decoder.generator.bias"])


# ---------------------------------------------------------------------------
# Best-first "beam" search (:708-867 + postprocessing/decoding.py:56-91)
# ---------------------------------------------------------------------------
@dataclass(eq=False)
class _Node:
    prev: Optional["_Node"]
    token: int
    logp: float  # Python float == fp64 (:821, decoding.py:80)
    length: int
    kv: List  # per layer (K rows [1,t,D], V rows [1,t,D]) of ANCESTOR outputs

    def __lt__(self, other):  # decoding.py:83-84
        return self.length < other.length


def beam_search(sd, spec: ModelSpec, src, beam_width: int = 5, max_sequence: int = 230):
    """EfficientSATRN.beam_search :708-867 with topk=1: per-sample best-first
    search over a global priority queue; score = -logp/len (:747,:824);
    budget (max_sequence-1) expansions (:754,:831); stops at the first popped
    EOS (:764-767); row starts with SOS, PAD-padded/truncated to max_sequence
    (:857-865).  Queue ordering = Python tuple order on (score, node) with
    node.__lt__ on len, ties beyond that resolved by heap order exactly as
    queue.PriorityQueue (heapq) does.  Returns LongTensor [B, max_sequence]."""
    outs = []
    temp = float(spec.dec_hidden) ** 0.5
    for bi in range(src.size(0)):
        st = DecoderState(sd, spec, src[bi:bi + 1])
        empty = [(torch.empty(1, 0, spec.dec_hidden), torch.empty(1, 0, spec.dec_hidden))
                 for _ in range(spec.dec_layers)]
        root = _Node(None, SOS_ID, 0.0, 1, empty)
        heap = [(-(root.logp / float(root.length)), root)]
        end = None
        num_steps = 0
        while True:
            if num_steps >= (max_sequence - 1) * beam_width:
                break
            score, n = heapq.heappop(heap)
            if n.token == EOS_ID and n.prev is not None:
                end = n
                break
            for l in range(spec.dec_layers):
                st.k_self[l], st.v_self[l] = n.kv[l]
            lg, new_kv = decode_step_cached(st, torch.tensor([n.token]), position=n.length - 1,
                                            append=False)
            child_kv = [(torch.cat([n.kv[l][0], new_kv[l][0]], 1),
                         torch.cat([n.kv[l][1], new_kv[l][1]], 1)) for l in range(spec.dec_layers)]
            logp = F.log_softmax(lg.view(1, 1, -1), dim=-1)
            lp, idx = torch.topk(logp, beam_width)
            for k in range(beam_width):
                child = _Node(n, int(idx[0, 0, k]), n.logp + lp[0, 0, k].item(), n.length + 1, child_kv)
                heapq.heappush(heap, (-(child.logp / float(child.length)), child))
            num_steps += beam_width
        if end is None:
            _, end = heapq.heappop(heap)
        utt = []
        n = end
        while n is not None:
            utt.append(n.token)
            n = n.prev
        utt = utt[::-1]
        if len(utt) < max_sequence:
            utt = utt + [PAD_ID] * (max_sequence - len(utt))
        else:
            utt = utt[:max_sequence]
        outs.append(utt)
    return torch.tensor(outs)


# ---------------------------------------------------------------------------
# Whole model
# ---------------------------------------------------------------------------
def forward_greedy(sd, spec: ModelSpec, images, steps: int, as_written: bool = False):
    """EfficientSATRN.forward :697-706 with is_train=False -> logits [B,steps,V]."""
    src = encoder_forward(sd, spec, images)
    if as_written:
        return decode_greedy_as_written(sd, spec, src, steps)
    return decode_greedy(sd, spec, src, steps)[0]


def greedy_tokens(logits):
    """postprocessing/decoding.py:38-40  topk(1) over the vocab axis."""
    _, seq = torch.topk(logits.transpose(1, 2), 1, dim=1)
    return seq.squeeze(1)


def expected_tokens(batch: int, max_sequence: int = 230, fill_id: int = 158):
    """inference_modules/inference_single.py:52 + data/dataset.py:111-115:
    [SOS] + max_sequence x id('\\sin') + [EOS]  ->  [B, max_sequence+2]."""
    row = [SOS_ID] + [fill_id] * max_sequence + [EOS_ID]
    return torch.tensor([row] * batch, dtype=torch.long)


def min_margins(logits):
    """Per-image minimum top1-top2 logit gap over all steps."""
    top2 = torch.topk(logits, 2, dim=-1).values
    return (top2[..., 0] - top2[..., 1]).min(dim=1).values
