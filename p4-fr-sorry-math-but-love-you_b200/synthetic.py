"""Seeded random-init weights in the reference checkpoint layout, for
benchmarks and demos (there is no network for trained checkpoints; BASELINE.json
names "random-init weights").  Variance-preserving normal init so activations
neither vanish nor blow up through the 40-block trunk with identity BatchNorm
statistics (a default ``nn.Module`` init collapses them to 1e-7)."""
import torch

from . import layout


def synthetic_state_dict(dims, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = layout.encoder_shapes(dims)
    shapes.update(layout.decoder_shapes(dims))
    sd = {}
    for name, shape in shapes.items():
        leaf = name.rsplit(".", 1)[-1]
        is_norm = any(t in name for t in (".bn", "norm"))
        if leaf == "num_batches_tracked":
            t = torch.ones((), dtype=torch.long)
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif is_norm and leaf == "weight":
            t = torch.rand(shape, generator=g) * 0.4 + 0.8
        elif is_norm and leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif name.endswith("embedding.weight"):
            t = torch.randn(shape, generator=g) * 0.1
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            gain = 2.0 if len(shape) == 4 and ".se." not in name else 1.0
            if ".q_linear." in name or ".k_linear." in name:
                gain = 4.0
            t = torch.randn(shape, generator=g) * (gain / fan_in) ** 0.5
        sd[name] = t
    return sd
