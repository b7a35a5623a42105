"""Small invocation of the sampling / perplexity tail kernels for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_sampling.py
Covers the shared-memory path, the global-workspace path (large vocabulary), a vocabulary that is not a multiple of the
block size, the tie-cut slow path, the draw and the NLL kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easykv_b200 import sampling  # noqa: E402

torch.manual_seed(0)
x = torch.randn(3, 5003, device="cuda") * 2
x[0] = -20.0
x[0, 100:110] = 3.0                                   # ten equal tokens holding ~all the mass: top_p 0.55 keeps six
tok, prob, raw = sampling.sample_top_p(x, 1.0, 0.55, want_prob=True, want_raw=True)
assert int((prob[0] > 0).sum()) == 6 and bool((prob.gather(-1, tok) > 0).all())
y = torch.randn(2, 152064, device="cuda") * 2         # global-memory workspace path
tok, prob, _ = sampling.sample_top_p(y, 0.8, 0.9, want_prob=True)
assert abs(float(prob[1].sum()) - 1) < 1e-4
tok = sampling.sample_top_p(x, 1e-9, 1.0)[0]           # draw only, no prob output
assert torch.equal(tok[1:, 0], x[1:].argmax(-1))       # (row 0 holds ten equal maxima: any of them is a valid draw)
nll = sampling.token_nll(x, torch.tensor([0, 5002, 17], device="cuda"))
ref = torch.nn.functional.cross_entropy(x, torch.tensor([0, 5002, 17], device="cuda"), reduction="none")
assert torch.allclose(nll, ref, rtol=1e-5, atol=1e-5)
torch.cuda.synchronize()
print("sampling tail ok")
