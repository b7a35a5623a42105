"""How long bench.py's per-step input refresh (pool -> q / k_new / v_new) takes on its own (CUDA events, graph replay)."""
import torch, time
dev = "cuda"
L, B, H, D, POOL = 32, 64, 32, 128, 4
q = torch.zeros(L, B, H, 1, D, device=dev, dtype=torch.float16)
pool = torch.randn(POOL, *q.shape, device=dev, dtype=torch.float16)
cursor = torch.zeros(1, dtype=torch.int64, device=dev)
def variant_a():
    cursor.add_(1).remainder_(POOL)
    q.copy_(pool.index_select(0, cursor)[0])
flat = pool.view(POOL, -1)
def variant_b():
    cursor.add_(1).remainder_(POOL)
    torch.index_select(flat, 0, cursor, out=q.view(1, -1))
onehot = torch.zeros(POOL, device=dev, dtype=torch.float16)
def variant_c():
    cursor.add_(1).remainder_(POOL)
    torch.gather(flat, 0, cursor.view(1, 1).expand(1, flat.shape[1]), out=q.view(1, -1))
for name, fn in (("index_select+copy", variant_a), ("index_select out=", variant_b), ("gather out=", variant_c)):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(name, f"{e0.elapsed_time(e1) / 20 * 1e3:.1f} us per refresh of {q.numel() * 2 / 1e6:.1f} MB")
