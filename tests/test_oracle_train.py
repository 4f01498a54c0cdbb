"""Pin the train-step oracle (oracle/train.py) against the fixtures the REAL reference produced
(oracle/make_golden.py --train: the reference's modules in train mode, its criterion, clip_grad_norm_ and AdamW) --
CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import satrn, synth, train

GOLD = os.path.join(ROOT, "tests", "golden", "efficientsatrn_train.npz")


@pytest.mark.parametrize("seed", [0, 1])
def test_train_step_oracle_matches_reference(spec, seed):
    g = np.load(GOLD)
    tr = train.Trainer(synth.synth_state_dict(spec, seed), spec)
    names = [k for k in tr.sd if train.is_param(k)]
    assert names == list(g["names"])           # named_parameters() order == state_dict order of the parameters
    for it in range(3):
        x, e = train.synth_batch(spec, 4, 24, 10 * seed + it)
        if it == 0:
            loss, grads = tr.forward_backward(x, e)
            l2 = np.array([grads[n].norm().item() for n in names])
            ref = g["grad_l2_seed%d" % seed]
            assert np.all(np.abs(l2 - ref) <= 2e-4 * ref + 1e-7), np.abs(l2 - ref).max()
            gn = float(torch.nn.utils.clip_grad_norm_(tr.params, max_norm=tr.max_grad_norm))
            tr.opt.step()
        else:
            loss, gn = tr.step(x, e)
        # step 0 is exact; later steps carry Adam's first-step amplification of round-off (update = lr * sign(g) for every
        # element, however small its gradient), so two correct implementations drift apart at the 1e-3 level
        tol_l, tol_g = (1e-6, 1e-5) if it == 0 else (1e-3, 2e-2)
        assert abs(loss - g["loss_seed%d" % seed][it]) <= tol_l * abs(loss), (it, loss)
        assert abs(gn - g["grad_norm_seed%d" % seed][it]) <= tol_g * gn, (it, gn)


def test_train_forward_is_teacher_forced_with_batch_statistics(spec, ckpt0):
    """Train mode differs from eval only through BatchNorm (batch statistics) when dropout is 0."""
    x, e = train.synth_batch(spec, 2, 12, 5)
    sd = {k: (v.float() if v.is_floating_point() else v).clone() for k, v in ckpt0.items()}
    with torch.no_grad():
        lt = train.train_forward(sd, spec, x, e)
        le = satrn.teacher_forced(ckpt0, spec, satrn.encoder_forward(ckpt0, spec, x), e[:, :-1])
    assert lt.shape == le.shape == (2, 12, spec.num_classes)
    assert (lt - le).abs().max() > 1e-3
    assert int(sd["encoder.shallow_cnn.bn1.num_batches_tracked"]) == int(ckpt0["encoder.shallow_cnn.bn1.num_batches_tracked"]) + 1


LITE_GOLD = os.path.join(ROOT, "tests", "golden", "litesatrn_train.npz")


@pytest.mark.parametrize("seed", [0, 1])
def test_lite_train_step_oracle_matches_reference(seed):
    """The same training step for LiteSATRN -- the student of the reference's distillation loop
    (train_modules/train_distillation.py:95-96) -- against oracle/make_golden.py --train --lite (the real
    networks.LiteSATRN in train mode)."""
    from oracle.make_golden import LITE_SPEC
    lspec = satrn.ModelSpec(**LITE_SPEC)
    g = np.load(LITE_GOLD)
    tr = train.Trainer(synth.synth_state_dict(lspec, seed, calib_batch=4), lspec)
    names = [k for k in tr.sd if train.is_param(k)]
    assert names == list(g["names"])
    for it in range(3):
        x, e = train.synth_batch(lspec, 4, 24, 10 * seed + it)
        if it == 0:
            loss, grads = tr.forward_backward(x, e)
            l2 = np.array([grads[n].norm().item() for n in names])
            ref = g["grad_l2_seed%d" % seed]
            assert np.all(np.abs(l2 - ref) <= 2e-4 * ref + 1e-7), np.abs(l2 - ref).max()
            gn = float(torch.nn.utils.clip_grad_norm_(tr.params, max_norm=tr.max_grad_norm))
            tr.opt.step()
        else:
            loss, gn = tr.step(x, e)
        tol_l, tol_g = (1e-6, 1e-5) if it == 0 else (1e-3, 2e-2)
        assert abs(loss - g["loss_seed%d" % seed][it]) <= tol_l * abs(loss), (it, loss)
        assert abs(gn - g["grad_norm_seed%d" % seed][it]) <= tol_g * gn, (it, gn)
