"""Time the sampling / perplexity tail: one `ekv_sample_top_p` launch against the ATen composition it replaces
(the reference's logits_adapter arithmetic + torch.multinomial), and `ekv_token_nll` against cross_entropy.

    python tools/time_sampling.py  ->  one JSON line per shape (CUDA events on the current stream, median of 50)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easykv_b200 import _lib, sampling  # noqa: E402


def aten_tail(logits, temperature, top_p):
    """The reference's op sequence (easykv/easykv.py:115-134, :258), written out for timing only."""
    prob = torch.softmax(logits / temperature, dim=-1)
    sp, si = torch.sort(prob, descending=True, dim=-1)
    cs = torch.cumsum(sp, dim=-1)
    sp[(cs - sp) > top_p] = 0.0
    sp.div_(sp.sum(dim=-1, keepdim=True))
    final = torch.gather(sp, -1, torch.sort(si, descending=False, dim=-1)[1])
    raw = torch.softmax(logits, dim=-1)
    return torch.multinomial(final, num_samples=1), raw


def timed(fn, n=50):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


def main():
    lib = _lib.load()
    for rows, vocab in [(1, 32000), (64, 32000), (1, 128256), (64, 128256)]:
        x = torch.randn(rows, vocab, device="cuda") * 2.5
        l0 = lib.ekv_launch_count()
        sampling.sample_top_p(x, 0.8, 0.9)
        launches = lib.ekv_launch_count() - l0
        ours = timed(lambda: sampling.sample_top_p(x, 0.8, 0.9))
        aten = timed(lambda: aten_tail(x, 0.8, 0.9))
        print(json.dumps(dict(op="sample_top_p", rows=rows, vocab=vocab, us=round(ours, 1), aten_us=round(aten, 1),
                              our_launches=int(launches), note="ours includes the exponential_ draw; ATen's includes 3 host syncs")))
    for rows, vocab in [(64, 32000), (512, 32000)]:
        x = torch.randn(rows, vocab, device="cuda") * 2.5
        t = torch.randint(0, vocab, (rows,), device="cuda")
        ours = timed(lambda: sampling.token_nll(x, t))
        aten = timed(lambda: torch.nn.functional.cross_entropy(x, t, reduction="none"))
        gbs = rows * vocab * 4 / ours / 1e3
        print(json.dumps(dict(op="token_nll", rows=rows, vocab=vocab, us=round(ours, 1), aten_us=round(aten, 1), read_gbs=round(gbs, 1))))


if __name__ == "__main__":
    main()
