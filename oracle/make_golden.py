"""Generate tests/golden/*.npz from the REAL reference (container-only).

TEST INFRASTRUCTURE ONLY.  Run as ``python -m oracle.make_golden`` in the build
container where /root/reference exists.  The reference's own modules
(networks.EfficientSATRN.EfficientSATRN, postprocessing.decode, .beam_search)
are imported through oracle/ref_shim.py, loaded with the seeded synthetic
checkpoint (oracle/synth.py, ``load_state_dict(strict=True)``), and run on CPU
fp32.  Inputs are NOT stored: they are regenerated from the seed
(oracle.synth.synth_images / synth_state_dict are bit-reproducible); a sha256
of the checkpoint bytes is stored so a host that generates different weights
is detected.
"""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, satrn, synth  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def state_dict_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    ref = ref_shim.load_reference()
    spec = satrn.ModelSpec()
    for seed, batch in ((0, 4), (1, 2)):
        sd = synth.synth_state_dict(spec, seed)
        model = ref.networks.EfficientSATRN(ref_shim.reference_flags(), ref_shim.reference_vocab()).eval()
        model.load_state_dict(sd, strict=True)
        images = synth.synth_images(spec, batch, seed)
        expected = satrn.expected_tokens(batch)
        # greedy through the reference's own entry point (decoding.py:34-40)
        logits = model(images, expected, False, 0.0)
        tokens = ref.postprocessing.decode(model, images, expected=expected, method="greedy")
        memory = model.encoder(images).contiguous()
        out = dict(
            digest=np.array(state_dict_digest(sd)),
            memory=memory.numpy(),
            logits=logits.numpy().astype(np.float32),
            tokens=tokens.numpy(),
        )
        if seed == 0:
            loader = ref_shim.reference_loader()
            for bw in (4, 8):
                with ref_shim.cpu_get_device():
                    out["beam%d" % bw] = ref.postprocessing.decode(
                        model, images, data_loader=loader, expected=expected, method="beam",
                        beam_width=bw).numpy()
            # teacher-forced branch (:488-495), eval mode so dropout is identity
            g = torch.Generator().manual_seed(7)
            text = torch.randint(3, 244, (batch, 24), generator=g)
            text[:, 0] = satrn.SOS_ID
            text[0, 15:] = satrn.PAD_ID
            text[2, 9:] = satrn.PAD_ID
            exp_tf = torch.cat([text, torch.full((batch, 1), satrn.EOS_ID)], 1)
            out["tf_text"] = text.numpy()
            out["tf_logits"] = model(images, exp_tf, True, 1.0).numpy()
        path = os.path.join(GOLDEN_DIR, "efficientsatrn_seed%d.npz" % seed)
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB", "digest", out["digest"])


LITE_SPEC = dict(network="LiteSATRN", enc_hidden=256, enc_filter=256, enc_layers=1, enc_heads=4,
                 dec_src=256, dec_hidden=128, dec_filter=512, dec_layers=2, dec_heads=4)


def main_lite():
    """LiteSATRN (networks/LiteSATRN.py) greedy fixtures: 100 % reference code (no third-party arithmetic)."""
    torch.set_grad_enabled(False)
    ref = ref_shim.load_reference()
    spec = satrn.ModelSpec(**LITE_SPEC)
    sd = synth.synth_state_dict(spec, 0, calib_batch=4)
    model = ref.networks.LiteSATRN(ref_shim.reference_flags("LiteSATRN"), ref_shim.reference_vocab()).eval()
    model.load_state_dict(sd, strict=True)
    batch, steps = 3, 120
    images = synth.synth_images(spec, batch, 0)
    expected = satrn.expected_tokens(batch, steps - 1)
    logits = model(images, expected, False, 0.0)
    tokens = ref.postprocessing.decode(model, images, expected=expected, method="greedy")
    out = dict(digest=np.array(state_dict_digest(sd)), memory=model.encoder(images).contiguous().numpy(),
               logits=logits.numpy(), tokens=tokens.numpy())
    path = os.path.join(GOLDEN_DIR, "litesatrn_seed0.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main_swin():
    """SwinTRN (networks/SWIN.py) greedy fixtures from the reference's own SWIN class (100 % reference code)."""
    import yaml
    from oracle import swin
    torch.set_grad_enabled(False)
    ref = ref_shim.load_reference()
    torch.hub.load_state_dict_from_url = lambda *a, **k: {"model": {}}   # SWIN.py:1033 downloads ImageNet weights
    d = yaml.safe_load(open(os.path.join(ref_shim.REFERENCE_ROOT, "configs", "SWIN.yaml")))
    d["dropout_rate"] = 0.1
    spec = swin.swin_spec()
    sd = swin.synth_state_dict(spec, 0)
    model = ref.networks.SWIN(ref.utils.Flags(d).get(), ref_shim.reference_vocab(), checkpoint=None).eval()
    model.load_state_dict(sd, strict=True)
    batch, steps = 2, 40
    images = swin.synth_images(batch, 0)
    expected = satrn.expected_tokens(batch, steps - 1)
    logits = model(images, expected, False, 0.0)
    memory = model.encoder(images)
    out = dict(digest=np.array(state_dict_digest(sd)), memory_sub=memory[:, ::4].contiguous().numpy(),
               logits=logits.numpy(), tokens=logits.argmax(-1).numpy())
    path = os.path.join(GOLDEN_DIR, "swin_seed0.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")



def compile_rules(rules, tokens):
    """Per-token tables equivalent to the reference's RULES dict (postprocessing/postprocessing.py:12-156) for a
    given token list: bit 0 cannot_initial, 1 next_underbar, 2 next_lbracket, 3 cannot_next_underbar,
    4 cannot_next_lbracket; limit[v] = limit_params[token] where limit_series[token] is true, else 0."""
    V = len(tokens)
    flags = np.zeros(V, np.int32)
    limit = np.zeros(V, np.int32)
    for bit, key in enumerate(("cannot_initial", "next_underbar", "next_lbracket", "cannot_next_underbar",
                               "cannot_next_lbracket")):
        for t in rules[key]:
            flags[tokens.index(t)] |= 1 << bit
    for v, t in enumerate(tokens):
        if rules["limit_series"].get(t, False):
            limit[v] = rules["limit_params"][t]
    return flags, limit


def main_manager():
    """Greedy decode under the reference's DecodingManager (postprocessing/postprocessing.py:182-405,
    EfficientSATRN.py:536-564): masked-softmax outputs and tokens of the REAL reference."""
    import importlib
    torch.set_grad_enabled(False)
    ref = ref_shim.load_reference()
    pp = importlib.import_module("postprocessing.postprocessing")
    spec = satrn.ModelSpec()
    batch = 4
    manager = pp.get_decoding_manager(os.path.join(ref_shim.REFERENCE_ROOT, "configs", "tokens.txt"), batch_size=batch)
    # The reference calls `self.manager.reset()` after the loop (:564) although reset() requires `sequence_length`
    # (postprocessing.py:237) -- a TypeError once all steps are done.  The wrapper supplies the missing argument;
    # nothing the loop produced depends on it.
    _reset = manager.reset
    manager.reset = lambda sequence_length=None: _reset(sequence_length)
    sd = synth.synth_state_dict(spec, 0)
    model = ref.networks.EfficientSATRN(ref_shim.reference_flags(), ref_shim.reference_vocab(), None, manager).eval()
    model.load_state_dict(sd, strict=True)
    images = synth.synth_images(spec, batch, 0)
    expected = satrn.expected_tokens(batch)
    with ref_shim.cpu_get_device():
        probs = model(images, expected, False, 0.0)            # [B, 231, 245] masked softmax (:553-555)
    tokens = probs.argmax(-1)
    flags, limit = compile_rules(manager.rules, manager.tokens)
    keep = np.array([0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 230])
    out = dict(digest=np.array(state_dict_digest(sd)), tokens=tokens.numpy(), probs_steps=keep,
               probs=probs[:, keep].numpy().astype(np.float32), probs_max=probs.max(-1).values.numpy().astype(np.float32),
               vocab=np.array(manager.tokens), flags=flags, limit=limit)
    path = os.path.join(GOLDEN_DIR, "efficientsatrn_seed0_manager.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    plain = np.load(os.path.join(GOLDEN_DIR, "efficientsatrn_seed0.npz"))["tokens"]
    print("tokens changed by the manager: %.1f %%" % (100.0 * (plain != tokens.numpy()).mean()))


def main_ensemble():
    """utils/ensemble_utils.py:46-120 (make_decoder_values) of the REAL reference on two seeded checkpoints: the ensemble's
    averaged distributions [B, T, V] and token sequences, without and with the DecodingManager.  The function is called
    as written; its `torch` name is proxied only to capture the `decoded_values` tensor it hands to torch.topk."""
    import importlib
    import types
    torch.set_grad_enabled(False)
    ref = ref_shim.load_reference()
    eu = importlib.import_module("utils.ensemble_utils")
    pp = importlib.import_module("postprocessing.postprocessing")
    spec = satrn.ModelSpec()
    batch, max_sequence = 3, 23
    flags, vocab = ref_shim.reference_flags(), ref_shim.reference_vocab()
    sds = [synth.synth_state_dict(spec, seed) for seed in (0, 1)]
    images = synth.synth_images(spec, batch, 5)
    decoders, memories = [], []
    for sd in sds:
        enc = sys.modules[ref.networks.EfficientSATRN.__module__].EfficientSATRN_encoder(flags, vocab).eval()
        enc.load_state_dict({k: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
        dec = sys.modules[ref.networks.EfficientSATRN.__module__].EfficientSATRN_decoder(flags, vocab).eval()
        dec.load_state_dict({k: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
        with ref_shim.cpu_get_device():
            memories.append(enc(images))
        decoders.append(dec)
    captured = []

    class TorchProxy:
        def __getattr__(self, name):
            return getattr(torch, name)

        def topk(self, x, **kw):
            captured.append(x.clone())
            return torch.topk(x, **kw)

    eu.torch = TorchProxy()
    eu.id_to_string = lambda sequences, loader, do_eval=0: [" ".join(str(int(t)) for t in row) for row in sequences]
    parser = types.SimpleNamespace(max_sequence=max_sequence)
    dec_loader = [(["img%d" % i for i in range(batch)], memories)]
    out = dict(digest0=np.array(state_dict_digest(sds[0])), digest1=np.array(state_dict_digest(sds[1])),
               memory0=memories[0].numpy(), memory1=memories[1].numpy())
    for tag, manager in (("plain", None), ("managed", pp.get_decoding_manager(
            os.path.join(ref_shim.REFERENCE_ROOT, "configs", "tokens.txt"), batch_size=batch))):
        del captured[:]
        with ref_shim.cpu_get_device():
            res = eu.make_decoder_values(decoders, parser, None, dec_loader, manager, torch.device("cpu"))
        probs = captured[0]
        if probs.dim() == 4:
            probs = probs.squeeze(2)
        out["probs_" + tag] = probs.numpy().astype(np.float32)            # [B, 24, 245]
        out["tokens_" + tag] = np.array([[int(t) for t in r[1].split()] for r in res], np.int64)
        print(tag, "tokens[0]:", out["tokens_" + tag][0][:12], "probs", out["probs_" + tag].shape)
        if manager is not None:
            fl, lim = compile_rules(manager.rules, manager.tokens)
            out["flags"], out["limit"], out["vocab"] = fl, lim, np.array(manager.tokens)
    eu.torch = torch
    path = os.path.join(GOLDEN_DIR, "efficientsatrn_ensemble.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main_train():
    """Three iterations of the reference's single-optimizer training step (train_modules/train_single_opt.py:72-112)
    on the REAL reference modules in train mode: teacher forcing 1.0, CrossEntropyLoss(ignore_index=PAD),
    clip_grad_norm_(2.0), AdamW(lr 5e-4, weight_decay 1e-6).  Every nn.Dropout is set to p = 0 (the decoder's Feedforward
    hard-codes 0.1 whatever FLAGS.dropout_rate says, EfficientSATRN.py:327,368-370; dropout masks come from torch's global
    RNG and cannot be reproduced by another implementation).  Stores per-step loss / gradient norm and, for step 0, the L2
    norm of every parameter's gradient."""
    from oracle import train
    ref = ref_shim.load_reference()
    lite = "--lite" in sys.argv     # LiteSATRN student of the distillation loop (train_distillation.py:95-96): same step
    spec = satrn.ModelSpec(**LITE_SPEC) if lite else satrn.ModelSpec()
    out = {}
    for seed in (0, 1):
        if lite:
            sd = synth.synth_state_dict(spec, seed, calib_batch=4)
            model = ref.networks.LiteSATRN(ref_shim.reference_flags("LiteSATRN", dropout=0.0), ref_shim.reference_vocab())
        else:
            sd = synth.synth_state_dict(spec, seed)
            model = ref.networks.EfficientSATRN(ref_shim.reference_flags(dropout=0.0), ref_shim.reference_vocab())
        model.load_state_dict(sd, strict=True)
        model.train()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        params = [p for p in model.parameters() if p.requires_grad]
        opt = ref.utils.get_optimizer("AdamW", params, lr=5e-4, weight_decay=1e-6)
        losses, norms = [], []
        for it in range(3):
            x, e = train.synth_batch(spec, 4, 24, 10 * seed + it)
            logits = model(x, e, True, 1.0)
            loss = model.criterion(logits.transpose(1, 2), e[:, 1:])
            opt.zero_grad()
            loss.backward()
            if it == 0:
                names = [n for n, p in model.named_parameters() if p.requires_grad]
                out["names"] = np.array(names)
                out["grad_l2_seed%d" % seed] = np.array([p.grad.norm().item() for p in params], np.float64)
                out["logits0_seed%d" % seed] = logits.detach().numpy()[:, ::4].astype(np.float32)
            norms.append(float(torch.nn.utils.clip_grad_norm_(params, max_norm=2.0)))
            opt.step()
            losses.append(loss.item())
        out["loss_seed%d" % seed] = np.array(losses, np.float64)
        out["grad_norm_seed%d" % seed] = np.array(norms, np.float64)
        print("seed", seed, "loss", losses, "grad norm", norms)
    path = os.path.join(GOLDEN_DIR, "litesatrn_train.npz" if lite else "efficientsatrn_train.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    if "--train" in sys.argv:
        main_train()
        sys.exit(0)
    if "--manager" in sys.argv:
        main_manager()
        sys.exit(0)
    if "--ensemble" in sys.argv:
        main_ensemble()
        sys.exit(0)
    if "--swin" in sys.argv:
        main_swin()
        sys.exit(0)
    if "--lite" in sys.argv:
        main_lite()
        sys.exit(0)
    main()
