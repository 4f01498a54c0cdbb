"""Shared test helpers (test infrastructure)."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TOKENS = ["<SOS>", "<EOS>", "<PAD>"]


class Vocab:
    """Stand-in for the dataset object the constructors read
    (networks/EfficientSATRN.py:679-692): 245 classes, SOS=0, EOS=1, PAD=2."""

    def __init__(self, n=245):
        names = TOKENS + ["tok%d" % i for i in range(n - 3)]
        self.token_to_id = {t: i for i, t in enumerate(names)}
        self.id_to_token = {i: t for i, t in enumerate(names)}


def flags_dict(height=128, width=256, rgb=1):
    """configs/EfficientSATRN.yaml with input_size / data.rgb overridden."""
    return {
        "network": "EfficientSATRN",
        "input_size": {"height": height, "width": width},
        "SATRN": {
            "encoder": {"hidden_dim": 512, "filter_dim": 512, "layer_num": 2, "head_num": 8},
            "decoder": {"src_dim": 512, "hidden_dim": 256, "filter_dim": 1024, "layer_num": 3, "head_num": 8},
        },
        "data": {"rgb": rgb},
        "dropout_rate": 0.1,
    }


def make_model(state_dict=None, precision="fp32", **kw):
    import frx
    flags = frx.Flags(flags_dict()).get()
    return frx.EfficientSATRN(flags, Vocab(), state_dict, None, precision=precision, **kw)


def lite_flags_dict(height=128, width=256, rgb=1):
    """configs/LiteSATRN.yaml with data.rgb overridden."""
    return {
        "network": "LiteSATRN",
        "input_size": {"height": height, "width": width},
        "SATRN": {
            "encoder": {"hidden_dim": 256, "filter_dim": 256, "layer_num": 1, "head_num": 4},
            "decoder": {"src_dim": 256, "hidden_dim": 128, "filter_dim": 512, "layer_num": 2, "head_num": 4},
        },
        "data": {"rgb": rgb},
        "dropout_rate": 0.1,
    }


def make_lite_model(state_dict=None, **kw):
    import frx
    return frx.LiteSATRN(frx.Flags(lite_flags_dict()).get(), Vocab(), state_dict, None, **kw)


def swin_flags_dict():
    """configs/SWIN.yaml (the SATRN.encoder block is ignored by SWIN)."""
    return {
        "network": "SWIN",
        "input_size": {"height": 384, "width": 384},
        "SATRN": {
            "encoder": {"hidden_dim": 300, "filter_dim": 600, "layer_num": 6, "head_num": 8},
            "decoder": {"src_dim": 1024, "hidden_dim": 512, "filter_dim": 512, "layer_num": 4, "head_num": 8},
        },
        "data": {"rgb": 3},
        "dropout_rate": 0.1,
    }


def make_swin_model(state_dict=None, **kw):
    import frx
    return frx.SWIN(frx.Flags(swin_flags_dict()).get(), Vocab(), state_dict, **kw)
