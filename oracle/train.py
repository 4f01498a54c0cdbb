"""CPU restatement of the reference's teacher-forced TRAINING step (the ORACLE for config 4).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows train_modules/train_single_opt.py:72-112:
``model.train()`` forward with teacher forcing (networks/EfficientSATRN.py:488-495; BatchNorm batch statistics),
``CrossEntropyLoss(ignore_index=PAD)`` on ``output.transpose(1, 2)`` vs ``expected[:, 1:]`` (:82-86,
EfficientSATRN.py:690-692), ``loss.backward()``, ``clip_grad_norm_(params, max_grad_norm)`` (:95), AdamW step
(utils/utils.py:91-92; configs/EfficientSATRN.yaml: lr 5e-4, weight_decay 1e-6, max_grad_norm 2.0).
Dropout is 0 in every parity run (``dropout_rate: 0``): the reference draws its masks from torch's global RNG, which no
other implementation can reproduce.  Pinned against the reference's own modules by tests/test_oracle_train.py
(build container) and the fixtures oracle/make_golden.py --train writes.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import satrn

NON_PARAM = ("running_mean", "running_var", "num_batches_tracked")


def is_param(name: str) -> bool:
    return not name.endswith(NON_PARAM)


def train_forward(sd, spec: satrn.ModelSpec, images, expected, momentum: float = 0.1):
    """EfficientSATRN.forward(input, expected, True, 1.0) in train mode -> logits [B, L, V] (L = expected.size(1)-1)."""
    mode = satrn.TrainMode(momentum)
    src = satrn.encoder_forward(sd, spec, images, calib=mode)
    return satrn.teacher_forced(sd, spec, src, expected[:, :-1])


def loss_fn(logits, expected):
    """train_single_opt.py:82-86 with criterion = CrossEntropyLoss(ignore_index=PAD) (EfficientSATRN.py:690-692)."""
    return F.cross_entropy(logits.transpose(1, 2), expected[:, 1:], ignore_index=satrn.PAD_ID)


def synth_batch(spec: satrn.ModelSpec, batch: int, max_len: int, seed: int):
    """Synthetic training batch: randn images; targets [SOS] + n tokens uniform in [3, V-2] + [EOS], n
    uniform in [4, max_len-1], PAD-padded to max_len + 1 columns."""
    g = torch.Generator().manual_seed(3000 + seed)
    images = torch.randn(batch, spec.in_ch, spec.height, spec.width, generator=g)
    expected = torch.full((batch, max_len + 1), satrn.PAD_ID, dtype=torch.long)
    for b in range(batch):
        n = int(torch.randint(4, max_len, (1,), generator=g))
        expected[b, 0] = satrn.SOS_ID
        expected[b, 1:1 + n] = torch.randint(3, spec.num_classes - 1, (n,), generator=g)
        expected[b, 1 + n] = satrn.EOS_ID
    return images, expected


class Trainer:
    """State of the reference's single-optimizer loop: parameters (leaf tensors over the state_dict), AdamW."""

    def __init__(self, sd: Dict[str, torch.Tensor], spec: satrn.ModelSpec, lr: float = 5e-4, weight_decay: float = 1e-6,
                 max_grad_norm: float = 2.0):
        self.spec, self.max_grad_norm = spec, max_grad_norm
        # the synthetic checkpoint keeps its calibrated running statistics in fp64; a module would hold them in fp32
        self.sd = {k: (v.float() if v.is_floating_point() else v).clone() for k, v in sd.items()}
        for k in self.sd:
            if is_param(k):
                self.sd[k].requires_grad_(True)
        self.params = [v for k, v in self.sd.items() if is_param(k)]
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay)

    def forward_backward(self, images, expected):
        """-> (loss, {name: grad}) with the gradients left in .grad (no clipping)."""
        self.opt.zero_grad()
        with torch.enable_grad():
            logits = train_forward(self.sd, self.spec, images, expected)
            loss = loss_fn(logits, expected)
            loss.backward()
        return loss.item(), {k: v.grad for k, v in self.sd.items() if is_param(k)}

    def step(self, images, expected):
        """One iteration of train_single_opt.py:72-112 -> (loss, grad_norm before clipping)."""
        loss, _ = self.forward_backward(images, expected)
        gn = torch.nn.utils.clip_grad_norm_(self.params, max_norm=self.max_grad_norm)
        self.opt.step()
        return loss, float(gn)

    def state_dict(self):
        return {k: v.detach() for k, v in self.sd.items()}
