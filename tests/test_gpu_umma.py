"""The Blackwell tensor-core primitives the strided-prefill chunk kernel is built from (csrc/ekv_umma.cuh): tensor-map TMA
tiles with the 128-byte swizzle, hand-built shared-memory / instruction descriptors for K-major and MN-major operands,
tcgen05.mma into tensor memory, tcgen05.ld back — pinned against torch through the C ABI (ekv_debug_umma_probe)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_umma_probe_matches_torch(ekv_lib, dtype):
    from easykv_b200 import _lib
    g = torch.Generator().manual_seed(5)
    K = torch.randn(128, 128, generator=g).to(dtype).cuda()
    V = torch.randn(128, 128, generator=g).to(dtype).cuda()
    Q = (torch.randn(64, 128, generator=g) * 0.3).to(dtype).cuda()
    Pt = torch.rand(128, 64, generator=g).to(dtype).cuda()
    St = torch.full((128, 64), float("nan"), device="cuda")
    Ot = torch.full((128, 64), float("nan"), device="cuda")
    code = {torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}[dtype]
    _lib.check(ekv_lib.ekv_debug_umma_probe(code, K.data_ptr(), V.data_ptr(), Q.data_ptr(), Pt.data_ptr(), St.data_ptr(),
                                            Ot.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    St_ref = K.float() @ Q.float().T                      # [128 keys, 64 rows]
    Ot_ref = V.float().T @ Pt.float()                     # [128 dims, 64 rows]
    assert (St - St_ref).abs().max().item() <= 2e-4 * St_ref.abs().max().item() + 1e-5, (St - St_ref).abs().max().item()
    assert (Ot - Ot_ref).abs().max().item() <= 2e-4 * Ot_ref.abs().max().item() + 1e-5, (Ot - Ot_ref).abs().max().item()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_umma_probe_16_rows_matches_torch(ekv_lib, dtype):
    """The GQA decode kernel's operand shapes: N = 16 rows, P^T as an un-swizzled MN-major operand (dtype | 0x100)."""
    from easykv_b200 import _lib
    g = torch.Generator().manual_seed(6)
    K = torch.randn(128, 128, generator=g).to(dtype).cuda()
    V = torch.randn(128, 128, generator=g).to(dtype).cuda()
    Q = (torch.randn(16, 128, generator=g) * 0.3).to(dtype).cuda()
    Pt = torch.rand(128, 16, generator=g).to(dtype).cuda()
    St = torch.full((128, 16), float("nan"), device="cuda")
    Ot = torch.full((128, 16), float("nan"), device="cuda")
    code = {torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}[dtype] | 0x100
    _lib.check(ekv_lib.ekv_debug_umma_probe(code, K.data_ptr(), V.data_ptr(), Q.data_ptr(), Pt.data_ptr(), St.data_ptr(),
                                            Ot.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    St_ref = K.float() @ Q.float().T
    Ot_ref = V.float().T @ Pt.float()
    assert (St - St_ref).abs().max().item() <= 2e-4 * St_ref.abs().max().item() + 1e-5, (St - St_ref).abs().max().item()
    assert (Ot - Ot_ref).abs().max().item() <= 2e-4 * Ot_ref.abs().max().item() + 1e-5, (Ot - Ot_ref).abs().max().item()
