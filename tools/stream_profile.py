import sys; sys.path.insert(0, "tools"); import sweep, torch
sweep.run_stream_case("7B b64 streaming", 64, 32, 32, 1088)
# breakdown: rope_cache alone
from easykv_b200.cache import BudgetedKVCache
import ctypes as C
from easykv_b200 import _lib
B,H,Hkv,n,d=64,32,32,1088,128
cache=BudgetedKVCache(1,B,H,Hkv,d,n+1,dtype=torch.float16,arith=1); cache.enable_streaming()
cache.load_prefill(0, torch.randn(B,Hkv,n,d,device="cuda").half(), torch.randn(B,Hkv,n,d,device="cuda").half(), n, [1.0]*n)
inv = 1.0 / (10000.0 ** (torch.arange(0, d, 2, device="cuda").float() / d))
emb = torch.cat([torch.outer(torch.arange(n + 8, device="cuda").float(), inv)] * 2, dim=-1)
cos, sin = emb.cos().half(), emb.sin().half()
shape=cache._shape(0,0); io=cache._io(0)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
for rep in range(2):
    e0.record()
    for _ in range(10):
        cache.lib.ekv_rope_cache(C.byref(shape), C.byref(io), cache.K_raw[0].data_ptr(), cos.data_ptr(), sin.data_ptr(), torch.cuda.current_stream().cuda_stream)
    e1.record(); torch.cuda.synchronize()
print("rope_cache us", e0.elapsed_time(e1)*100, "GB/s", 2*B*Hkv*n*d*2/ (e0.elapsed_time(e1)*1e-4) /1e9)
k_raw=torch.randn(B,1,Hkv*d,device="cuda").half().view(B,1,Hkv,d).transpose(1,2)
slots=torch.randint(0,n,(B,Hkv,1),device="cuda")
e0.record()
for _ in range(10):
    cache.K_raw[0].scatter_(2, slots[..., None].expand(B, Hkv, 1, d), k_raw)
e1.record(); torch.cuda.synchronize()
print("scatter us", e0.elapsed_time(e1)*100)
