"""CPU yardstick for the bf16 mode of the EfficientSATRN encoder (test infrastructure: imports oracle/).

Runs the oracle trunk + encoder layers under several rounding policies and reports, against the fp32 oracle on the
same synthetic checkpoint: relative L2 error of the encoder memory, forced-decoding logit error and the free-running
token agreement of a 231-step greedy decode (fp32 decoder fed the perturbed memory).  Policies:

  autocast   torch.autocast("cpu", bfloat16) over the oracle encoder: what stock PyTorch mixed precision gives
  frx_r1     the round-1 CUDA pipeline: bf16 operands / fp32 accumulation, EVERY activation stored in bf16 (residual
             stream included), SE gate applied in place (a second bf16 rounding of the expanded map)
  frx_r2     round 2: the narrow block outputs (the residual stream, 24..256 channels) stay fp32; only GEMM operands
             and the wide expanded maps are bf16; the SE gate is applied once
  operands   bf16 GEMM operands only (weights + A operand rounded), everything else fp32: the floor of any pipeline
             that feeds bf16 into the tensor cores

The library's encoder now feeds IEEE fp16 operands (common.cuh, eh_t); ``--dtype fp16`` applies the same policies with
fp16 rounding.  Measured on 16 images, seed 0 (memory rel-L2 / forced max-rel / forced argmax / free-running tokens):
  bf16: operands 0.0797 / 0.105 / .. / 0.81,   frx_r1 0.094 / 0.14 / 0.85 / 0.86 (measured on the GPU in round 1)
  fp16: operands 0.0101 / 0.016 / 0.976 / 0.938, frx_r1 0.0115 / 0.021 / 0.977 / 0.920

    python tools/bf16_yardstick.py [--batch 16] [--seed 0] [--steps 231] [--dtype bf16|fp16]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import satrn, synth  # noqa: E402


OPERAND_DTYPE = [torch.bfloat16]


def set_operand_dtype(dtype):
    """Rounding applied by every policy: torch.bfloat16 (default) or torch.float16 (what the library's encoder uses)."""
    OPERAND_DTYPE[0] = dtype


def rb(x):
    return x.to(OPERAND_DTYPE[0]).to(torch.float32)


class Policy:
    def __init__(self, name, store_wide=True, store_res=True, se_twice=True, enc_bf16=True):
        self.name, self.store_wide, self.store_res, self.se_twice, self.enc_bf16 = name, store_wide, store_res, se_twice, enc_bf16

    def wide(self, x):  # expanded / intermediate maps
        return rb(x) if self.store_wide else x

    def res(self, x):   # block outputs (residual stream)
        return rb(x) if self.store_res else x


def bn(x, sd, p, eps):
    return satrn._bn(x, sd, p, eps)


def trunk_emulated(sd, spec, x, pol):
    e = "encoder.shallow_cnn."
    W = lambda n: rb(sd[n])
    conv = lambda a, w, stride=1, groups=1: satrn._conv_same(rb(a), w, stride, groups)
    x = F.conv2d(x, sd[e + "conv_stem.weight"], None, 2, 0)          # stem: fp32 arithmetic on the fp32 image
    x = pol.res(F.silu(bn(x, sd, e + "bn1", 1e-3)))
    for pfx, kind, cin, cout, k, stride, expand, se_r in satrn.trunk_blocks():
        p = e + pfx
        res = x if (cin == cout and stride == 1) else None
        if kind == "cn":
            y = F.silu(bn(conv(x, W(p + ".conv.weight"), stride), sd, p + ".bn1", 1e-3))
        elif kind == "er":
            y = pol.wide(F.silu(bn(conv(x, W(p + ".conv_exp.weight"), stride), sd, p + ".bn1", 1e-3)))
            y = bn(F.conv2d(rb(y), W(p + ".conv_pwl.weight")), sd, p + ".bn2", 1e-3)
        else:
            y = pol.wide(F.silu(bn(F.conv2d(rb(x), W(p + ".conv_pw.weight")), sd, p + ".bn1", 1e-3)))
            y = satrn._conv_same(y, sd[p + ".conv_dw.weight"], stride, groups=y.shape[1])   # fp32 taps, fp32 sum
            y = F.silu(bn(y, sd, p + ".bn2", 1e-3))
            s = y.mean((2, 3), keepdim=True)
            s = F.silu(F.conv2d(s, sd[p + ".se.conv_reduce.weight"], sd[p + ".se.conv_reduce.bias"]))
            s = F.conv2d(s, sd[p + ".se.conv_expand.weight"], sd[p + ".se.conv_expand.bias"])
            if pol.se_twice:
                y = pol.wide(pol.wide(y) * torch.sigmoid(s))
            else:
                y = pol.wide(y * torch.sigmoid(s))
            y = bn(F.conv2d(rb(y), W(p + ".conv_pwl.weight")), sd, p + ".bn3", 1e-3)
        x = pol.res(y + res if res is not None else y)
    x = F.conv2d(rb(x), W(e + "conv_last.weight"))
    return F.silu(bn(x, sd, e + "bn2", 1e-5))


def encoder_layer_emulated(sd, spec, i, x):
    p = "encoder.attention_layers.%d." % i
    W = lambda n: sd[p + n]
    b, c, h, w = x.shape
    flat = x.view(b, c, h * w).transpose(1, 2)
    y = rb(F.layer_norm(flat, (c,), W("norm.weight"), W("norm.bias"), 1e-5))
    a = p + "attention_layer"
    heads, d = spec.enc_heads, c
    lin = lambda t, n: F.linear(t, rb(sd[a + n + ".weight"]), sd[a + n + ".bias"])
    q, k, v = lin(y, ".q_linear"), lin(y, ".k_linear"), lin(y, ".v_linear")
    o = rb(satrn._attend(q, k, v, heads, float(d) ** 0.5))
    y = lin(o, ".out_linear")
    y = rb(F.layer_norm(y + flat, (c,), W("norm.weight"), W("norm.bias"), 1e-5))
    y = y.reshape(-1, c, h, w)
    y = rb(F.relu(bn(F.conv2d(y, rb(W("conv0.weight"))), sd, p + "norm0", 1e-5)))
    y = F.conv2d(y, W("depthwise.weight"), W("depthwise.bias"), 1, 1, 1, y.shape[1])
    y = rb(F.relu(bn(y, sd, p + "depthwise_norm", 1e-5)))
    y = F.relu(bn(F.conv2d(y, rb(W("conv1.weight"))), sd, p + "norm1", 1e-5))
    return y + x


def encoder_emulated(sd, spec, images, pol):
    x = trunk_emulated(sd, spec, images, pol)
    x = satrn.pe2d_forward(sd, spec, x)
    for i in range(spec.enc_layers):
        x = encoder_layer_emulated(sd, spec, i, x) if pol.enc_bf16 else satrn.encoder_layer_forward(sd, spec, i, x)
    b, c, h, w = x.shape
    return x.view(b, c, h * w).transpose(1, 2).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--steps", type=int, default=231)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    a = ap.parse_args()
    set_operand_dtype(torch.float16 if a.dtype == "fp16" else torch.bfloat16)
    torch.set_num_threads(os.cpu_count() or 1)
    spec = satrn.ModelSpec()
    sd = synth.synth_state_dict(spec, a.seed)
    x = synth.synth_images(spec, a.batch, a.seed)
    with torch.no_grad():
        ref = satrn.encoder_forward(sd, spec, x)
        lg_ref, tok_ref = satrn.decode_greedy(sd, spec, ref, a.steps)
        mems = {}
        with torch.autocast("cpu", dtype=torch.bfloat16):
            mems["autocast"] = satrn.encoder_forward(sd, spec, x).float()
        mems["frx_r1"] = encoder_emulated(sd, spec, x, Policy("frx_r1"))
        mems["frx_r2"] = encoder_emulated(sd, spec, x, Policy("frx_r2", store_res=False, se_twice=False))
        mems["operands"] = encoder_emulated(sd, spec, x, Policy("operands", store_wide=False, store_res=False, se_twice=False))
        print("%-10s %12s %14s %14s %12s" % ("policy", "mem rel-L2", "forced max-rel", "forced argmax", "free tokens"))
        for name, m in mems.items():
            lg_f, _ = satrn.decode_greedy(sd, spec, m, a.steps, forced_tokens=tok_ref)
            _, tok = satrn.decode_greedy(sd, spec, m, a.steps)
            d = lg_f - lg_ref
            print("%-10s %12.4f %14.4f %14.4f %12.4f" % (
                name, (m - ref).norm().item() / ref.norm().item(), d.abs().max().item() / lg_ref.abs().max().item(),
                (lg_f.argmax(-1) == lg_ref.argmax(-1)).float().mean().item(), (tok == tok_ref).float().mean().item()))


if __name__ == "__main__":
    main()
