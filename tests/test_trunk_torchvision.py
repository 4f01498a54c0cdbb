"""Independent pin of the restated timm trunk (oracle/satrn.py::trunk_forward).

timm 0.4.9 is absent from the reference tree and from this image, so the 40 EfficientNetV2-S blocks of the
oracle are a restatement.  torchvision's ``efficientnet_v2_s().features[1:7]`` is an independently written
implementation of the same published topology (same 19 846 552 parameters): loading the synthetic checkpoint
into it -- with timm's TF-"same" asymmetric padding put in front of the five stride-2 convolutions, the only
place the two libraries differ -- and comparing block by block checks the oracle's block arithmetic
(ConvBnAct / EdgeResidual / InvertedResidual + SqueezeExcite, BN eps 1e-3, SiLU, skip rules) against code
that neither the builder nor the reference wrote.  CPU only.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import satrn, synth

torchvision = pytest.importorskip("torchvision")


def _tf_same_pad_hook(conv):
    """timm layers/padding.py + conv2d_same.py: dynamic asymmetric padding for stride-2 convolutions."""
    k, s = conv.kernel_size[0], conv.stride[0]
    conv.padding = (0, 0)

    def hook(_m, args):
        x = args[0]
        ih, iw = x.shape[-2:]
        ph = max((math.ceil(ih / s) - 1) * s + k - ih, 0)
        pw = max((math.ceil(iw / s) - 1) * s + k - iw, 0)
        return (F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]),)

    conv.register_forward_pre_hook(hook)


def _load_bn(bn, sd, p):
    bn.weight.data.copy_(sd[p + ".weight"]); bn.bias.data.copy_(sd[p + ".bias"])
    bn.running_mean.copy_(sd[p + ".running_mean"]); bn.running_var.copy_(sd[p + ".running_var"])
    assert bn.eps == 1e-3


def _torchvision_trunk(sd):
    from torchvision.models import efficientnet_v2_s
    feats = efficientnet_v2_s(weights=None).features
    stages = [feats[i] for i in range(1, 7)]
    e = "encoder.shallow_cnn."
    it = iter(satrn.trunk_blocks())
    for stage in stages:
        for blk in stage:
            pfx, kind, cin, cout, k, stride, expand, se_r = next(it)
            p, b = e + pfx, blk.block
            if kind == "cn":
                b[0][0].weight.data.copy_(sd[p + ".conv.weight"]); _load_bn(b[0][1], sd, p + ".bn1")
                convs = [b[0][0]]
            elif kind == "er":
                b[0][0].weight.data.copy_(sd[p + ".conv_exp.weight"]); _load_bn(b[0][1], sd, p + ".bn1")
                b[1][0].weight.data.copy_(sd[p + ".conv_pwl.weight"]); _load_bn(b[1][1], sd, p + ".bn2")
                convs = [b[0][0]]
            else:
                b[0][0].weight.data.copy_(sd[p + ".conv_pw.weight"]); _load_bn(b[0][1], sd, p + ".bn1")
                b[1][0].weight.data.copy_(sd[p + ".conv_dw.weight"]); _load_bn(b[1][1], sd, p + ".bn2")
                b[2].fc1.weight.data.copy_(sd[p + ".se.conv_reduce.weight"]); b[2].fc1.bias.data.copy_(sd[p + ".se.conv_reduce.bias"])
                b[2].fc2.weight.data.copy_(sd[p + ".se.conv_expand.weight"]); b[2].fc2.bias.data.copy_(sd[p + ".se.conv_expand.bias"])
                b[3][0].weight.data.copy_(sd[p + ".conv_pwl.weight"]); _load_bn(b[3][1], sd, p + ".bn3")
                convs = [b[1][0]]
            for c in convs:
                if c.stride[0] == 2:
                    _tf_same_pad_hook(c)
    assert next(it, None) is None
    return torch.nn.Sequential(*stages).eval()


def test_oracle_trunk_blocks_match_torchvision(spec, ckpt0):
    tv = _torchvision_trunk(ckpt0)
    n_tv = sum(p.numel() for p in tv.parameters())
    assert n_tv == 19_846_552
    x = synth.synth_images(spec, 2, 0)
    taps = {}
    with torch.no_grad():
        satrn.trunk_forward(ckpt0, spec, x, taps=taps)
        y = taps["stem"]                      # the stem is reference code (:67-73), the blocks are timm's
        worst = 0.0
        names = [b[0] for b in satrn.trunk_blocks()]
        blocks = [blk for stage in tv for blk in stage]
        assert len(blocks) == len(names) == 40
        for name, blk in zip(names, blocks):
            y = blk(y)
            want = taps[name]
            assert y.shape == want.shape, name
            err = (y - want).abs().max().item() / max(want.abs().max().item(), 1e-6)
            worst = max(worst, err)
            assert err <= 2e-5, (name, err)   # fp32 round-off of two different conv call orders
            y = want                          # re-seed so a block's error cannot hide behind the next one
    print("worst per-block relative error vs torchvision: %.2e" % worst)
