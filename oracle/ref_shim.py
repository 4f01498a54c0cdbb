"""Import shim for the REAL reference (test infrastructure, container-only).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and this particular
module is only usable where ``/root/reference`` exists (the build container),
never on the GPU box.  It is used by ``oracle/make_golden.py`` to produce the
committed fixtures under ``tests/golden/`` and by the CPU tests that pin the
restatement (``oracle/satrn.py``) against the reference itself.

What it does:
  1. puts ``/root/reference`` on ``sys.path``;
  2. stubs the packages that are not installed here (``albumentations``,
     ``editdistance``, ``timm``) in ``sys.modules``;
  3. replaces ``timm.create_model`` by a builder that returns the restated
     timm-0.4.9 ``tf_efficientnetv2_s`` ``.blocks`` (App. A.1) -- the
     third-party arithmetic that is NOT under /root/reference
     (``requirements.txt:15``, call site ``networks/EfficientSATRN.py:66,74``);
  4. imports ``utils`` before ``networks`` (circular import);
  5. patches ``PositionEncoder1D.forward`` (``networks/EfficientSATRN.py:420-426``)
     to use ``x.device`` instead of ``x.get_device()`` so the CPU path runs
      -- a semantic no-op.
"""
import math
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("FRX_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "networks"))


# --------------------------------------------------------------------------
# timm 0.4.9 restatement (module form, so that state_dict names match the
# reference checkpoint layout: encoder.shallow_cnn.eff_block.<stage>.<idx>.*)
# --------------------------------------------------------------------------
class _Conv2dSame(nn.Conv2d):
    """TF-style dynamic 'same' padding (timm layers/conv2d_same.py)."""

    def forward(self, x):
        ih, iw = x.shape[-2:]
        kh, kw = self.weight.shape[-2:]
        sh, sw = self.stride
        ph = max((math.ceil(ih / sh) - 1) * sh + kh - ih, 0)
        pw = max((math.ceil(iw / sw) - 1) * sw + kw - iw, 0)
        if ph > 0 or pw > 0:
            x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2])
        return F.conv2d(x, self.weight, self.bias, self.stride, (0, 0), self.dilation, self.groups)


def _conv(cin, cout, k, stride=1, groups=1, bias=False):
    if stride == 1:
        return nn.Conv2d(cin, cout, k, stride=1, padding=(k - 1) // 2, groups=groups, bias=bias)
    return _Conv2dSame(cin, cout, k, stride=stride, padding=0, groups=groups, bias=bias)


def _bn(c):
    return nn.BatchNorm2d(c, eps=1e-3)


class _ConvBnAct(nn.Module):
    def __init__(self, cin, cout, k, stride, skip):
        super().__init__()
        self.has_residual = skip and stride == 1 and cin == cout
        self.conv = _conv(cin, cout, k, stride)
        self.bn1 = _bn(cout)
        self.act1 = nn.SiLU(inplace=True)

    def forward(self, x):
        y = self.act1(self.bn1(self.conv(x)))
        return y + x if self.has_residual else y


class _EdgeResidual(nn.Module):
    def __init__(self, cin, cout, k, stride, expand):
        super().__init__()
        mid = cin * expand
        self.has_residual = cin == cout and stride == 1
        self.conv_exp = _conv(cin, mid, k, stride)
        self.bn1 = _bn(mid)
        self.act1 = nn.SiLU(inplace=True)
        self.conv_pwl = _conv(mid, cout, 1)
        self.bn2 = _bn(cout)

    def forward(self, x):
        y = self.act1(self.bn1(self.conv_exp(x)))
        y = self.bn2(self.conv_pwl(y))
        return y + x if self.has_residual else y


class _SqueezeExcite(nn.Module):
    def __init__(self, chs, reduced):
        super().__init__()
        self.conv_reduce = nn.Conv2d(chs, reduced, 1, bias=True)
        self.act1 = nn.SiLU(inplace=True)
        self.conv_expand = nn.Conv2d(reduced, chs, 1, bias=True)

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        s = self.conv_expand(self.act1(self.conv_reduce(s)))
        return x * s.sigmoid()


class _InvertedResidual(nn.Module):
    def __init__(self, cin, cout, k, stride, expand, se_ratio):
        super().__init__()
        mid = cin * expand
        self.has_residual = cin == cout and stride == 1
        self.conv_pw = _conv(cin, mid, 1)
        self.bn1 = _bn(mid)
        self.act1 = nn.SiLU(inplace=True)
        self.conv_dw = _conv(mid, mid, k, stride, groups=mid)
        self.bn2 = _bn(mid)
        self.act2 = nn.SiLU(inplace=True)
        self.se = _SqueezeExcite(mid, int(cin * se_ratio))
        self.conv_pwl = _conv(mid, cout, 1)
        self.bn3 = _bn(cout)

    def forward(self, x):
        y = self.act1(self.bn1(self.conv_pw(x)))
        y = self.act2(self.bn2(self.conv_dw(y)))
        y = self.se(y)
        y = self.bn3(self.conv_pwl(y))
        return y + x if self.has_residual else y


# (kind, repeats, kernel, stride, expand, out_ch, se_ratio)  -- SURVEY App. A.1
EFFNETV2_S_ARCH = [
    ("cn", 2, 3, 1, 1, 24, 0.0),
    ("er", 4, 3, 2, 4, 48, 0.0),
    ("er", 4, 3, 2, 4, 64, 0.0),
    ("ir", 6, 3, 2, 4, 128, 0.25),
    ("ir", 9, 3, 1, 6, 160, 0.25),
    ("ir", 15, 3, 2, 6, 256, 0.25),
]


def build_effnetv2_s_blocks(stem_chs=24):
    stages = []
    cin = stem_chs
    for kind, reps, k, stride, expand, cout, se in EFFNETV2_S_ARCH:
        blocks = []
        for r in range(reps):
            s = stride if r == 0 else 1
            if kind == "cn":
                blocks.append(_ConvBnAct(cin, cout, k, s, skip=True))
            elif kind == "er":
                blocks.append(_EdgeResidual(cin, cout, k, s, expand))
            else:
                blocks.append(_InvertedResidual(cin, cout, k, s, expand, se))
            cin = cout
        stages.append(nn.Sequential(*blocks))
    return nn.Sequential(*stages)


class _TimmModelStub:
    def __init__(self):
        self.blocks = build_effnetv2_s_blocks()


_LOADED = None


def load_reference():
    """Return a namespace with the reference's own modules imported."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    stub("albumentations")
    stub("albumentations.pytorch", ToTensorV2=object)
    stub("editdistance", eval=lambda a, b: 0)
    timm = stub("timm", create_model=lambda name, pretrained=False, **kw: _TimmModelStub())
    stub("timm.models")
    layers = stub(
        "timm.models.layers",
        DropPath=nn.Identity,
        to_2tuple=lambda x: x if isinstance(x, tuple) else (x, x),
        trunc_normal_=nn.init.trunc_normal_,
    )
    timm.models = sys.modules["timm.models"]
    timm.models.layers = layers

    import utils  # noqa: F401  (must come first: circular import chain)
    import networks  # noqa: F401
    import postprocessing  # noqa: F401

    def _pe_forward(self, x, point=-1):
        # networks/EfficientSATRN.py:420-426 with x.get_device() -> x.device
        if point == -1:
            out = x + self.position_encoder[:, : x.size(1), :].to(x.device)
            out = self.dropout(out)
        else:
            out = x + self.position_encoder[:, point, :].unsqueeze(1).to(x.device)
        return out

    for modname in ("networks.EfficientSATRN", "networks.LiteSATRN", "networks.SWIN"):
        mod = sys.modules.get(modname)
        if mod is not None and hasattr(mod, "PositionEncoder1D"):
            mod.PositionEncoder1D.forward = _pe_forward

    ns = types.SimpleNamespace(
        utils=sys.modules["utils"],
        networks=sys.modules["networks"],
        postprocessing=sys.modules["postprocessing"],
        satrn=sys.modules["networks.EfficientSATRN"],
        lite=sys.modules.get("networks.LiteSATRN"),
        swin=sys.modules.get("networks.SWIN"),
    )
    _LOADED = ns
    return ns


class cpu_get_device:
    """Context manager: ``tensor.get_device()`` returns ``tensor.device`` so the
    reference's ``.to(input.get_device())`` (:774) works on CPU."""

    def __enter__(self):
        self._orig = torch.Tensor.get_device
        torch.Tensor.get_device = lambda t: t.device
        return self

    def __exit__(self, *a):
        torch.Tensor.get_device = self._orig
        return False


class _Loader:
    def __init__(self, dataset):
        self.dataset = dataset


def reference_loader():
    """Object with the ``.dataset.token_to_id`` the beam search reads (:717-719)."""
    return _Loader(reference_vocab())


class VocabDataset:
    """Stand-in for the dataset object the constructors read
    (``networks/EfficientSATRN.py:679-692``): only the two vocab dicts."""

    def __init__(self, token_to_id, id_to_token):
        self.token_to_id = token_to_id
        self.id_to_token = id_to_token


def reference_flags(network="EfficientSATRN", height=128, width=256, rgb=1, dropout=0.1):
    """FLAGS = the reference yaml with input_size / data.rgb overridden
    (SURVEY 8d)."""
    import yaml

    ref = load_reference()
    with open(os.path.join(REFERENCE_ROOT, "configs", network + ".yaml")) as f:
        d = yaml.safe_load(f)
    d["input_size"] = {"height": height, "width": width}
    d["data"]["rgb"] = rgb
    d["dropout_rate"] = dropout
    return ref.utils.Flags(d).get()


def reference_vocab():
    ref = load_reference()
    t2i, i2t = ref.utils.load_vocab([os.path.join(REFERENCE_ROOT, "configs", "tokens.txt")])
    return VocabDataset(t2i, i2t)
