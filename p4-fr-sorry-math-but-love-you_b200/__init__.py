"""B200-native EfficientSATRN hot path behind the reference's nn.Module API.

The directory name follows the build contract (``p4-fr-sorry-math-but-love-you_b200``)
and is not a valid Python identifier; import it through the alias module
``frx`` at the repository root (``import frx``), which loads this package.

Public surface (mirrors /root/reference):
  networks.EfficientSATRN / EfficientSATRN_encoder / EfficientSATRN_decoder
      (networks/EfficientSATRN.py:664-952)
  decoding.decode                (postprocessing/decoding.py:6-53)
  ensemble.make_decoder_values   (utils/ensemble_utils.py:46-120)
  flags.Flags                    (utils/flags.py:32-45)
The compute lives in ``lib/libfrx.so`` (C ABI in include/frx.h).  There is no
CPU or PyTorch fallback: importing works anywhere, but creating a model handle
without the compiled library or without a B200 raises RuntimeError.
"""
from . import _lib, data, decoding, ensemble, flags, layout, networks, sharding, synthetic  # noqa: F401
from ._lib import library_path, load_library  # noqa: F401
from .decoding import decode  # noqa: F401
from .flags import Flags  # noqa: F401
from .networks import EfficientSATRN, EfficientSATRN_decoder, EfficientSATRN_encoder, LiteSATRN, SWIN  # noqa: F401

__all__ = ["EfficientSATRN", "LiteSATRN", "SWIN", "EfficientSATRN_encoder", "EfficientSATRN_decoder", "decode", "Flags",
           "load_library", "library_path"]
