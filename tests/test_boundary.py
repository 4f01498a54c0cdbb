"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/frx.h declares, the nn.Module front-end has the reference's
state_dict layout, and the product fails loudly without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from helpers import ROOT, Vocab, flags_dict, make_model
from oracle import ref_shim, satrn

import frx


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "frx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(frx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = frx.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(frx._lib.SYMBOLS) == declared
    assert lib.frx_version().decode().endswith("sm_100a")


def test_library_has_no_torch_dependency():
    import subprocess
    out = subprocess.run(["ldd", frx.library_path()], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out and "libcudart.so" not in out


def test_sass_is_sm100_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", frx.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_state_dict_layout_matches_reference_layout(spec, ckpt0):
    model = make_model()
    sd = model.state_dict()
    shapes = satrn.param_shapes(spec)
    assert list(sd.keys()) == list(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    assert sd["encoder.shallow_cnn.bn1.num_batches_tracked"].dtype == torch.int64
    # a reference-layout checkpoint loads strictly, and round-trips
    res = model.load_state_dict(ckpt0, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(model.state_dict()["decoder.generator.weight"], ckpt0["decoder.generator.weight"])
    assert sum(p.numel() for p in model.parameters()) == 27_221_141
    assert model.decoder.st_id == 0 and model.decoder.layer_num == 3 and model.decoder.manager is None
    assert isinstance(model.criterion, torch.nn.CrossEntropyLoss) and model.criterion.ignore_index == 2


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_checkpoint_interchange_with_reference(ckpt0):
    ref = ref_shim.load_reference()
    theirs = ref.networks.EfficientSATRN(ref_shim.reference_flags(), ref_shim.reference_vocab())
    ours = make_model()
    assert list(theirs.state_dict().keys()) == list(ours.state_dict().keys())
    ours.load_state_dict(theirs.state_dict(), strict=True)
    theirs.load_state_dict(ours.state_dict(), strict=True)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_flags_match_reference_loader():
    import yaml
    ref = ref_shim.load_reference()
    path = os.path.join(ref_shim.REFERENCE_ROOT, "configs", "EfficientSATRN.yaml")
    a = ref.utils.Flags(yaml.safe_load(open(path))).get()
    b = frx.Flags(path).get()
    assert a.SATRN.decoder.hidden_dim == b.SATRN.decoder.hidden_dim == 256
    assert a.optimizer.lr == b.optimizer.lr == 5e-4
    assert a.input_size.height == b.input_size.height and a.data.rgb == b.data.rgb


def test_positional_tables_match_oracle():
    from frx.networks import pe1d_table, pe2d_table
    assert torch.equal(pe1d_table(256), satrn.pe1d_table(256))
    assert torch.equal(pe2d_table(4, 512), satrn.pe2d_table(4, 512))
    assert torch.equal(pe2d_table(8, 512), satrn.pe2d_table(8, 512))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(ckpt0):
    model = make_model(ckpt0).eval()
    x = torch.zeros(2, 1, 128, 256)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(x, satrn.expected_tokens(2, 4), False, 0.0)
    cfg = frx._lib.FrxConfig()
    cfg.max_batch, cfg.max_steps = 2, 4
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        frx._lib.Handle(cfg)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setenv("FRX_LIBRARY", "/nonexistent/libfrx.so")
    monkeypatch.setattr(frx._lib, "_LIB", None)
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        frx.load_library()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "p4-fr-sorry-math-but-love-you_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_decode_rejects_unknown_method(ckpt0):
    with pytest.raises(NotImplementedError):
        frx.decode(make_model(), torch.zeros(2, 1, 128, 256), method="sampling")
