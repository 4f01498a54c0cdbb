"""CPU restatement of the reference's DecodingManager (rule-constrained greedy decoding).

TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

Follows /root/reference/postprocessing/postprocessing.py:
  * ``DecodingManager.sift``     :193-233  softmax -> zero the black-listed classes -> argmax -> record
  * ``MemoryNode.record``        :303-325  run-length of the current token, cumulative '{' / '}' counts
  * ``MemoryNode._look_back``    :327-391  the black-list as a function of (current token, run length, brackets)
and its call site in the greedy loop, /root/reference/networks/EfficientSATRN.py:536-564 (the model then returns the
MASKED SOFTMAX rows instead of logits, and the next input token is the constrained argmax).

The rule tables are the per-token compilation of the reference's RULES dict (``oracle.make_golden.compile_rules``):
``flags`` bit 0 cannot_initial, 1 next_underbar, 2 next_lbracket, 3 cannot_next_underbar, 4 cannot_next_lbracket;
``limit[v]`` = maximum run length of token v (0 = unlimited).  Pinned against the reference's own output in
tests/golden/efficientsatrn_seed0_manager.npz (tests/test_oracle.py).
"""
from typing import List

import torch

from oracle import satrn

BIT_CANNOT_INITIAL, BIT_NEXT_UNDERBAR, BIT_NEXT_LBRACKET, BIT_NO_UNDERBAR, BIT_NO_LBRACKET = 1, 2, 4, 8, 16


class Rules:
    def __init__(self, vocab: List[str], flags, limit):
        self.vocab = list(vocab)
        self.flags = [int(x) for x in flags]
        self.limit = [int(x) for x in limit]
        ids = {t: i for i, t in enumerate(self.vocab)}
        self.sos, self.eos, self.empty = ids["<SOS>"], ids["<EOS>"], ids[""]
        self.lbrace, self.rbrace, self.underbar = ids["{"], ids["}"], ids["_"]

    def as_manager(self):
        """Duck-typed stand-in for the reference's DecodingManager (``.rules`` / ``.tokens``), i.e. what a caller of
        ``frx.EfficientSATRN(..., decoding_manager=...)`` passes."""
        v = self.vocab

        def having(bit):
            return [v[i] for i, f in enumerate(self.flags) if f & bit]

        rules = {
            "cannot_initial": having(BIT_CANNOT_INITIAL), "next_underbar": having(BIT_NEXT_UNDERBAR),
            "next_lbracket": having(BIT_NEXT_LBRACKET), "cannot_next_underbar": having(BIT_NO_UNDERBAR),
            "cannot_next_lbracket": having(BIT_NO_LBRACKET),
            "limit_series": {t: self.limit[i] > 0 for i, t in enumerate(v)},
            "limit_params": {t: self.limit[i] for i, t in enumerate(v) if self.limit[i] > 0},
        }

        class _Manager:
            pass

        m = _Manager()
        m.rules, m.tokens = rules, list(v)
        return m


class Node:
    """MemoryNode (:272-325): what one sample has generated so far."""

    def __init__(self, rules: Rules):
        self.r = rules
        self.cur, self.run, self.nl, self.nr = rules.sos, 1, 0, 0

    def record(self, tok: int):                              # :303-325
        self.run = self.run + 1 if tok == self.cur else 1
        if tok == self.r.lbrace:
            self.nl += 1
        elif tok == self.r.rbrace:
            self.nr += 1
        self.cur = tok

    def blacklist(self) -> List[int]:                        # :327-391
        r = self.r
        bl = {r.sos, r.empty}
        if self.nl == self.nr:
            bl.add(r.rbrace)
        if self.cur == r.eos:
            return sorted(bl)
        V = len(r.vocab)
        if self.cur == r.sos:
            bl.update(v for v in range(V) if r.flags[v] & BIT_CANNOT_INITIAL)
            return sorted(bl)
        f = r.flags[self.cur]
        if f & BIT_NEXT_UNDERBAR:
            bl.update(v for v in range(V) if v != r.underbar)
            return sorted(bl)
        if f & BIT_NEXT_LBRACKET:
            bl.update(v for v in range(V) if v != r.lbrace)
            return sorted(bl)
        if f & BIT_NO_UNDERBAR:
            bl.add(r.underbar)
        if f & BIT_NO_LBRACKET:
            bl.add(r.lbrace)
        if r.limit[self.cur] > 0 and self.run >= r.limit[self.cur]:
            bl.add(self.cur)
        return sorted(bl)


def sift(nodes: List[Node], logits: torch.Tensor):
    """DecodingManager.sift (:193-233): logits [B, V] -> (targets [B], masked softmax [B, V])."""
    probs = torch.softmax(logits, dim=-1)
    mask = torch.zeros_like(probs, dtype=torch.bool)
    for b, n in enumerate(nodes):
        mask[b, n.blacklist()] = True
    probs = probs.masked_fill(mask, 0)
    targets = torch.argmax(probs, dim=-1)
    for t, n in zip(targets.tolist(), nodes):
        n.record(t)
    return targets, probs


def decode_greedy_managed(sd, spec, src, steps: int, rules: Rules):
    """The greedy loop with a manager (EfficientSATRN.py:536-564) on the cached recurrence.
    Returns (masked softmax [B, steps, V], tokens [B, steps])."""
    st = satrn.DecoderState(sd, spec, src)
    nodes = [Node(rules) for _ in range(src.size(0))]
    tok = torch.full((src.size(0),), satrn.SOS_ID, dtype=torch.long)
    outs, toks = [], []
    for _ in range(steps):
        lg, _ = satrn.decode_step_cached(st, tok)
        tok, probs = sift(nodes, lg)
        outs.append(probs)
        toks.append(tok)
    return torch.stack(outs, 1), torch.stack(toks, 1)
