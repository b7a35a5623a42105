"""Parity against the reference RUNNING ON THE B200: tests/golden_gpu/*.npz are runs of the unmodified reference
(oracle/_ref, vendored by tools/vendor_ref.sh) in fp16 / bf16 on CUDA, recorded on the GPU box by
oracle/gen_golden_gpu.py — cuBLAS GEMMs, ATen's CUDA softmax / division kernels, CUDA topk.  They are replayed here
teacher-forced through the C ABI with `arith=1`, the product's default arithmetic flavour (what `easykv_generate` and
bench.py run), on every kernel family: eviction ids identical to the reference's, attention outputs within 1e-3 (fp16),
the exported cache bit-identical to the reference's final cache."""
import pytest
import torch

from oracle import replay

pytestmark = pytest.mark.gpu

CASES = replay.list_golden_gpu()


def test_gpu_goldens_present():
    assert len(CASES) >= 8


def _check(rep, name):
    assert rep.n_events > 0
    assert not rep.victim_mismatch, rep.victim_mismatch[:2]
    for f, l, ref, got, margin in rep.tie_ambiguous:      # exact ties: torch.topk's pick is unspecified
        assert min(margin) == 0.0
    assert len(rep.tie_ambiguous) <= 1 or "bf16" in name
    assert rep.final_cache_equal
    assert rep.max_out_err <= (8e-3 if "bf16" in name else 1e-3), rep.max_out_err


@pytest.mark.parametrize("arith", [1, 0], ids=["aten_cuda", "aten_cpu"])
@pytest.mark.parametrize("name", CASES)
def test_reference_on_b200_replay(ekv_lib, name, arith):
    import engines as E
    shadow = lambda *a: replay.OracleEngine(*a, scale_mul=bool(arith))
    rep = replay.replay(name, lambda *a: E.CudaEngine(*a, arith=arith), resync=True, shadow=shadow,
                        golden_dir=replay.GOLDEN_GPU_DIR)
    _check(rep, name)


@pytest.mark.parametrize("chunk_variant", [2], ids=["mma_sync_chunks"])
@pytest.mark.parametrize("name", [n for n in CASES if "enc" in n or "auto" in n or "ppl" in n])
def test_reference_on_b200_replay_two_pass_chunk_kernels(ekv_lib, name, chunk_variant):
    """The fallback chunk path (mma.sync two-pass kernels) on the same traces."""
    import engines as E
    ekv_lib.ekv_debug_set_chunk_variant(chunk_variant)
    try:
        rep = replay.replay(name, lambda *a: E.CudaEngine(*a, arith=1), resync=True,
                            shadow=lambda *a: replay.OracleEngine(*a, scale_mul=True), golden_dir=replay.GOLDEN_GPU_DIR)
    finally:
        ekv_lib.ekv_debug_set_chunk_variant(0)
    _check(rep, name)


@pytest.mark.parametrize("variant,cluster", [(2, -1), (1, -1), (0, 2), (6, 0), (5, 0), (5, 2)],
                         ids=["pingpong", "one_group", "cluster2", "round1_dispatch", "tcgen05", "tcgen05_pair"])
@pytest.mark.parametrize("name", [n for n in CASES if "decoding" in n or "auto" in n])
def test_reference_on_b200_replay_decode_kernels(ekv_lib, name, variant, cluster):
    """Decode steps of the B200-recorded runs forced through each decode kernel: the persistent ping-pong kernel
    (decode_kernel<T,G,2>, the headline instantiation), its one-group form, the cluster-split kernel, the round-1
    dispatch as a whole, and the tcgen05 GQA decode kernel (one CTA per unit and CTA pairs)."""
    import engines as E
    ekv_lib.ekv_debug_set_dispatch(variant, cluster)
    try:
        rep = replay.replay(name, lambda *a: E.CudaEngine(*a, arith=1), resync=True,
                            shadow=lambda *a: replay.OracleEngine(*a, scale_mul=True), golden_dir=replay.GOLDEN_GPU_DIR)
    finally:
        ekv_lib.ekv_debug_set_dispatch(0, 0)
    _check(rep, name)
