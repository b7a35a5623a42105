"""End-to-end run of the user surface on a full-size random-weight model (development aid; bench.py is the
contract): BASELINE configs[1] — Llama-2-7B shape, fp16, 4K prompt, mode='auto', budget 1024, stride 64, roco —
through `enable_fixed_kv` / `model.easykv_generate` bound to the INSTALLED transformers model classes.

    python tools/e2e_generate.py [--layers 32] [--prompt 4096] [--new 64] [--arch llama|mistral]

Prints one JSON line: prefill seconds, decode tokens/s (whole model: projections, MLP, sampling and the Python
driver included — at batch 1 that is launch- and weight-bound, the hot path is a few percent of it), retained
cache, number of eviction events.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import easykv_b200  # noqa: E402


class Tok:
    eos_token_id = -1

    def decode(self, ids, skip_special_tokens=True):
        return " ".join(str(int(i)) for i in ids)

    def convert_ids_to_tokens(self, ids):
        return [str(int(i)) for i in ids]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="llama")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--prompt", type=int, default=4096)
    ap.add_argument("--new", type=int, default=64)
    ap.add_argument("--budget", type=float, default=1024, help="integer = absolute, fraction < 1 = ratio of the prompt (encoding / ppl)")
    ap.add_argument("--mode", default="auto")
    ap.add_argument("--keep-attention", action="store_true")
    ap.add_argument("--stride", type=int, default=64)
    ap.add_argument("--policy", default="roco")
    ap.add_argument("--no-graph", action="store_true", help="eager decode loop (no CUDA-graph replay of the steady decode step)")
    args = ap.parse_args()
    import transformers
    if args.arch == "llama":
        cfg = transformers.LlamaConfig(hidden_size=4096, intermediate_size=11008, num_hidden_layers=args.layers,
                                       num_attention_heads=32, num_key_value_heads=32, vocab_size=32000,
                                       max_position_embeddings=8192, attn_implementation="eager")
        cls = transformers.LlamaForCausalLM
    else:
        cfg = transformers.MistralConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=args.layers,
                                         num_attention_heads=32, num_key_value_heads=8, vocab_size=32000,
                                         max_position_embeddings=32768, sliding_window=None, attn_implementation="eager")
        cls = transformers.MistralForCausalLM
    torch.manual_seed(0)
    with torch.device("cuda"):
        model = cls(cfg).half().eval()
    ids = torch.randint(3, 32000, (1, args.prompt), generator=torch.Generator().manual_seed(1)).cuda()
    with contextlib.redirect_stdout(io.StringIO()):
        easykv_b200.enable_fixed_kv(model, Tok(), mode=args.mode, stride=args.stride)
    budget = args.budget if args.budget < 1 else int(args.budget)
    gen = dict(temperature=1e-9, top_p=1.0, budget=budget, kv_policy=args.policy, keep_attention=args.keep_attention,
               cuda_graph=not args.no_graph)
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):           # untimed warm-up (cuBLAS handles, kernel attributes)
        model.easykv_generate(input_ids=ids[:, :max(2 * args.stride, 256)], generation_config=dict(gen, budget=budget if budget < 1 else min(budget, 64), max_new_tokens=2))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        model.easykv_generate(input_ids=ids, generation_config=dict(gen, max_new_tokens=args.new, record_timing=True))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    sess = model.easykv_last
    out = {"prefill_only": sess.t_prompt_done - t0, "printed": buf.getvalue().strip()}
    dec = t1 - sess.t_prompt_done
    tt = sess.token_times
    gaps = sorted((b - a) * 1e3 for a, b in zip(tt[:-1], tt[1:]))
    pct = lambda f: round(gaps[min(len(gaps) - 1, int(f * len(gaps)))], 3) if gaps else None   # noqa: E731
    first = [round((b - a) * 1e3, 2) for a, b in zip(tt[:6], tt[1:7])]
    print(json.dumps(dict(arch=args.arch, token_gap_ms=dict(p10=pct(0.1), p50=pct(0.5), p90=pct(0.9), max=pct(1.0), first=first), layers=args.layers, prompt=args.prompt, new_tokens=args.new, budget=budget, mode=args.mode,
                          keep_attention=args.keep_attention, stride=args.stride, policy=args.policy, prefill_s=round(out["prefill_only"], 3),
                          decode_tokens_per_s=round(args.new / dec, 2), decode_ms_per_token=round(dec / args.new * 1e3, 2),
                          retained=sess.cache.n[0], eviction_events=len(sess.events), printed=out["printed"],
                          launches=int(sess.cache.lib.ekv_launch_count()), graphed_steps=sess.graphed_steps, graphed_chunks=sess.graphed_chunks, graph_capture_s=round(sess.graph_capture_s, 3),
                          steady_ms_per_token=round((dec - sess.graph_capture_s) / args.new * 1e3, 2),
                          graph_error=sess.graph_error)))


if __name__ == "__main__":
    main()
