"""CPU restatement of the reference's SwinTRN encoder (networks/SWIN.py) -- ORACLE.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The decoder of SWIN
(networks/SWIN.py:922-1021, a copy of the SATRN decoder classes) is served by
oracle/satrn.py's decoder functions with a SWIN ModelSpec.  All citations are into
/root/reference/networks/SWIN.py.  Pinned against the reference's own
SwinTransformer through oracle/ref_shim.py (tests/golden/swin_seed0.npz).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from . import satrn

# SWIN.__init__ (:1028-1031): Swin-B / 384, patch 4, window 12, APE
IMG, PATCH, EMBED, WINDOW = 384, 4, 128, 12
DEPTHS, HEADS = (2, 2, 18, 2), (4, 8, 16, 32)
HEAD_CLASSES = 21841  # unused classification head that is still in the state_dict


def swin_spec(num_classes: int = 245) -> satrn.ModelSpec:
    """configs/SWIN.yaml decoder block (the yaml's SATRN.encoder block is ignored by SWIN)."""
    return satrn.ModelSpec(network="SWIN", height=IMG, width=IMG, in_ch=3, enc_hidden=1024, enc_filter=0,
                           enc_layers=0, enc_heads=0, dec_src=1024, dec_hidden=512, dec_filter=512, dec_layers=4,
                           dec_heads=8, num_classes=num_classes)


def relative_position_index(ws: int = WINDOW) -> torch.Tensor:
    """:116-135"""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shift_attn_mask(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """:286-310  (0 / -100 mask of the shifted-window blocks)"""
    img = torch.zeros((1, H, W, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, h, w, :] = cnt
            cnt += 1
    mw = img.view(1, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, float(-100.0)).masked_fill(m == 0, float(0.0))


def stage_layout():
    """(stage, dim, resolution, depth, heads, [(window, shift) per block])"""
    out = []
    res = IMG // PATCH
    for i, (depth, heads) in enumerate(zip(DEPTHS, HEADS)):
        dim, r = EMBED * 2 ** i, res // 2 ** i
        blocks = []
        for j in range(depth):
            ws, shift = WINDOW, (0 if j % 2 == 0 else WINDOW // 2)
            if r <= ws:  # :262-265
                ws, shift = r, 0
            blocks.append((ws, shift))
        out.append((i, dim, r, depth, heads, blocks))
    return out


def encoder_param_shapes() -> Dict[str, tuple]:
    """state_dict entries of SWIN.encoder in registration order (buffers included)."""
    s: Dict[str, tuple] = {}
    e = "encoder."
    n_patches = (IMG // PATCH) ** 2
    s[e + "absolute_pos_embed"] = (1, n_patches, EMBED)
    s[e + "patch_embed.proj.weight"] = (EMBED, 3, PATCH, PATCH)
    s[e + "patch_embed.proj.bias"] = (EMBED,)
    s[e + "patch_embed.norm.weight"] = (EMBED,)
    s[e + "patch_embed.norm.bias"] = (EMBED,)
    for i, dim, r, depth, heads, blocks in stage_layout():
        for j, (ws, shift) in enumerate(blocks):
            p = "%slayers.%d.blocks.%d." % (e, i, j)
            if shift > 0:
                s[p + "attn_mask"] = ((r // ws) ** 2, ws * ws, ws * ws)
            s[p + "norm1.weight"] = (dim,)
            s[p + "norm1.bias"] = (dim,)
            s[p + "attn.relative_position_bias_table"] = ((2 * ws - 1) ** 2, heads)
            s[p + "attn.relative_position_index"] = (ws * ws, ws * ws)
            s[p + "attn.qkv.weight"] = (3 * dim, dim)
            s[p + "attn.qkv.bias"] = (3 * dim,)
            s[p + "attn.proj.weight"] = (dim, dim)
            s[p + "attn.proj.bias"] = (dim,)
            s[p + "norm2.weight"] = (dim,)
            s[p + "norm2.bias"] = (dim,)
            s[p + "mlp.fc1.weight"] = (4 * dim, dim)
            s[p + "mlp.fc1.bias"] = (4 * dim,)
            s[p + "mlp.fc2.weight"] = (dim, 4 * dim)
            s[p + "mlp.fc2.bias"] = (dim,)
        if i < len(DEPTHS) - 1:
            p = "%slayers.%d.downsample." % (e, i)
            s[p + "reduction.weight"] = (2 * dim, 4 * dim)
            s[p + "norm.weight"] = (4 * dim,)
            s[p + "norm.bias"] = (4 * dim,)
    s[e + "norm.weight"] = (EMBED * 8,)
    s[e + "norm.bias"] = (EMBED * 8,)
    s[e + "head.weight"] = (HEAD_CLASSES, EMBED * 8)
    s[e + "head.bias"] = (HEAD_CLASSES,)
    return s


def param_shapes(spec: satrn.ModelSpec) -> Dict[str, tuple]:
    s = encoder_param_shapes()
    dec = satrn.param_shapes(satrn.ModelSpec(network="LiteSATRN", enc_hidden=32, enc_filter=32, enc_layers=0,
                                             dec_src=spec.dec_src, dec_hidden=spec.dec_hidden,
                                             dec_filter=spec.dec_filter, dec_layers=spec.dec_layers,
                                             dec_heads=spec.dec_heads, num_classes=spec.num_classes))
    for k, v in dec.items():
        if k.startswith("decoder."):  # SWIN's Feedforward is an nn.Sequential (:827-841): layers.0 / layers.3
            s[k.replace("feedforward_layer.linear0", "feedforward_layer.layers.0")
               .replace("feedforward_layer.linear1", "feedforward_layer.layers.3")] = v
    return s


def decoder_view(sd):
    """The same tensors under the SATRN decoder names, so oracle.satrn's decoder functions apply."""
    out = dict(sd)
    for k, v in sd.items():
        if "feedforward_layer.layers." in k:
            out[k.replace("feedforward_layer.layers.0", "feedforward_layer.linear0")
                 .replace("feedforward_layer.layers.3", "feedforward_layer.linear1")] = v
    return out


def synth_state_dict(spec: satrn.ModelSpec, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded synthetic SwinTRN checkpoint (no BatchNorm anywhere -> no calibration needed)."""
    g = torch.Generator().manual_seed(3000 + seed)
    sd = {}
    for name, shape in param_shapes(spec).items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "relative_position_index":
            t = relative_position_index(int(math.isqrt(shape[0])))
        elif leaf == "attn_mask":
            i = int(name.split(".")[2])
            r = (IMG // PATCH) // 2 ** i
            t = shift_attn_mask(r, r, WINDOW, WINDOW // 2)
        elif "norm" in name and leaf == "weight":
            t = torch.rand(shape, generator=g) * 0.4 + 0.8
        elif "norm" in name and leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif leaf == "relative_position_bias_table":
            t = torch.randn(shape, generator=g) * 0.5
        elif leaf == "absolute_pos_embed":
            t = torch.randn(shape, generator=g) * 0.2
        elif name == "decoder.embedding.weight":
            t = torch.randn(shape, generator=g) * 0.1
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            gain = 4.0 if (".q_linear." in name or ".k_linear." in name or ".qkv." in name) else 1.0
            t = torch.randn(shape, generator=g) * (gain / fan_in) ** 0.5
        sd[name] = t
    sd["decoder.generator.bias"][satrn.EOS_ID] += 1.0
    return sd


def synth_images(batch: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(4000 + seed)
    return torch.randn(batch, 3, IMG, IMG, generator=g)


def _window_attention(sd, p, x, heads, ws, mask):
    """WindowAttention.forward :149-193: q scaled by head_dim**-0.5, + relative position bias, + mask."""
    B_, N, C = x.shape
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"]).reshape(B_, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * ((C // heads) ** -0.5)
    attn = q @ k.transpose(-2, -1)
    bias = sd[p + "relative_position_bias_table"][sd[p + "relative_position_index"].view(-1)].view(N, N, -1)
    attn = attn + bias.permute(2, 0, 1).contiguous().unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, N, N)
    attn = torch.softmax(attn, dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def _block(sd, p, x, r, heads, ws, shift):
    """SwinTransformerBlock.forward :314-362 (drop_path = identity in eval)."""
    B, L, C = x.shape
    shortcut = x
    y = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5).view(B, r, r, C)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    win = y.view(B, r // ws, ws, r // ws, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws * ws, C)
    aw = _window_attention(sd, p + "attn.", win, heads, ws, sd[p + "attn_mask"] if shift > 0 else None)
    y = aw.view(B, r // ws, r // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, r, r, C)
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    x = shortcut + y.view(B, L, C)
    h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))  # exact GELU (:30,:43)
    return x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def encoder_forward(sd, images, taps=None):
    """SwinTransformer.forward_features :725-736 -> [B, 144, 1024]."""
    e = "encoder."
    x = F.conv2d(images, sd[e + "patch_embed.proj.weight"], sd[e + "patch_embed.proj.bias"], PATCH).flatten(2).transpose(1, 2)
    x = F.layer_norm(x, (EMBED,), sd[e + "patch_embed.norm.weight"], sd[e + "patch_embed.norm.bias"], 1e-5)
    x = x + sd[e + "absolute_pos_embed"]
    if taps is not None:
        taps["embed"] = x
    for i, dim, r, depth, heads, blocks in stage_layout():
        for j, (ws, shift) in enumerate(blocks):
            x = _block(sd, "%slayers.%d.blocks.%d." % (e, i, j), x, r, heads, ws, shift)
            if taps is not None:
                taps["block%d.%d" % (i, j)] = x
        if i < len(DEPTHS) - 1:  # PatchMerging :404-421
            p = "%slayers.%d.downsample." % (e, i)
            B, L, C = x.shape
            y = x.view(B, r, r, C)
            y = torch.cat([y[:, 0::2, 0::2, :], y[:, 1::2, 0::2, :], y[:, 0::2, 1::2, :], y[:, 1::2, 1::2, :]], -1)
            y = y.view(B, -1, 4 * C)
            y = F.layer_norm(y, (4 * C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
            x = F.linear(y, sd[p + "reduction.weight"])
            if taps is not None:
                taps["merge%d" % i] = x
    return F.layer_norm(x, (EMBED * 8,), sd[e + "norm.weight"], sd[e + "norm.bias"], 1e-5)
