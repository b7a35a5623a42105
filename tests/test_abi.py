"""The C-ABI library builds, loads, and exports exactly what include/easykv_b200.h declares (no GPU needed)."""
import ctypes

import pytest
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "easykv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ekv_[a-z_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(ekv_lib):
    from easykv_b200 import _lib
    names = _declared()
    assert set(names) == set(_lib.EXPORTS)
    for n in names:
        assert getattr(ekv_lib, n) is not None


def test_abi_version_and_struct_layout(ekv_lib):
    from easykv_b200 import _lib
    assert ekv_lib.ekv_abi_version() == _lib.ABI_VERSION
    assert ctypes.sizeof(_lib.Step) == 18 * 4
    assert ctypes.sizeof(_lib.Shape) == 9 * 4
    assert ctypes.sizeof(_lib.LayerIO) == 18 * 8


def test_ctypes_structs_mirror_the_header_field_for_field():
    """Field names, order and C types of ekv_step / ekv_shape / ekv_layer_io in include/easykv_b200.h == the ctypes
    mirrors in easykv_b200/_lib.py (an ABI drift would otherwise only show up as wrong results on the GPU)."""
    from easykv_b200 import _lib
    src = open(os.path.join(ROOT, "include", "easykv_b200.h")).read()
    ctype = {"int32_t": ctypes.c_int32, "float": ctypes.c_float}

    def fields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(\w+(?:\s*,\s*\w+)*)$", decl)      # `int32_t B, H, Hkv, d`
            assert m, decl
            for fname in re.split(r"\s*,\s*", m.group(4)):
                out.append((fname, ctypes.c_void_p if m.group(3) else ctype[m.group(2)]))
        return out

    for cname, mirror in (("ekv_step", _lib.Step), ("ekv_shape", _lib.Shape), ("ekv_layer_io", _lib.LayerIO)):
        assert fields(cname) == list(mirror._fields_), cname


def test_argument_validation_without_gpu(ekv_lib):
    """Bad shapes are rejected before any CUDA call (reference: ValueError, llama_patch.py:204-228)."""
    from easykv_b200 import _lib
    sh = _lib.Shape(dtype=_lib.F16, B=1, H=6, Hkv=4, d=128, q_len=1, cap=16, n_before=0, n_phys=0)
    io = _lib.LayerIO()
    st = _lib.Step()
    assert ekv_lib.ekv_attend_evict(ctypes.byref(sh), ctypes.byref(io), ctypes.byref(st), 0, None) == _lib.ERR_INVALID
    assert b"multiple" in ekv_lib.ekv_last_error()
    sh.H = 4
    sh.n_phys = 32
    assert ekv_lib.ekv_select(ctypes.byref(sh), ctypes.byref(io), ctypes.byref(st), None) == _lib.ERR_INVALID
    assert ekv_lib.ekv_launch_count() == 0


def test_scratch_bytes_is_host_only_and_consistent(ekv_lib):
    """ekv_scratch_bytes needs no GPU: 0 for decode steps, > 0 (and growing with the cache) for 16-bit chunks,
    0 for fp32 chunks (general kernel), the head-mean staging on top for tova in encoding / ppl."""
    from easykv_b200 import _lib
    st = _lib.Step(policy=_lib.POLICY_ROCO, accumulate=1, evict=16)
    def need(dtype, q_len, n, step=st):
        sh = _lib.Shape(dtype=dtype, B=2, H=32, Hkv=8, d=128, q_len=q_len, cap=n + q_len, n_before=n, n_phys=n)
        return ekv_lib.ekv_scratch_bytes(ctypes.byref(sh), ctypes.byref(step))
    assert need(_lib.F16, 1, 8208) == 0
    assert need(_lib.F32, 16, 8208) == 0
    a, b = need(_lib.F16, 16, 1024), need(_lib.F16, 16, 8208)
    assert 0 < a < b < 1 << 30
    assert need(_lib.BF16, 16, 8208) == b
    tova = _lib.Step(policy=_lib.POLICY_TOVA, accumulate=1, evict=16, tova_head_mean=1)
    assert need(_lib.F16, 16, 8208, tova) == b + 2 * 8 * (8208 + 16) * 4
    assert need(_lib.F32, 16, 8208, tova) == 2 * 8 * (8208 + 16) * 4


def test_entry_limit_is_host_only_and_plausible(ekv_lib):
    """ekv_chunk_entry_limit needs no GPU: ~17.6 K entries for an evicting 16-bit chunk on the tensor-core path, ~11 K on
    the exact kernel (fp32, kernel = 1, head_dim 64), tens of thousands for decode steps, no ceiling without eviction."""
    from easykv_b200 import _lib
    def lim(dtype, H, Hkv, d, q_len, evict, kernel=0):
        sh = _lib.Shape(dtype=dtype, B=1, H=H, Hkv=Hkv, d=d, q_len=q_len, cap=0, n_before=0, n_phys=0)
        return ekv_lib.ekv_chunk_entry_limit(ctypes.byref(sh), evict, kernel)
    assert 17000 < lim(_lib.F16, 32, 8, 128, 16, 16) < 18500
    assert lim(_lib.BF16, 32, 32, 128, 64, 64) == lim(_lib.F16, 32, 32, 128, 64, 64)
    assert 10000 < lim(_lib.F32, 32, 8, 128, 16, 16) < 12000
    assert lim(_lib.F16, 32, 8, 128, 16, 16, kernel=1) == lim(_lib.F32, 32, 8, 128, 16, 16)
    assert 10000 < lim(_lib.F16, 32, 32, 64, 64, 64) < 12000
    assert lim(_lib.F16, 32, 32, 128, 1, 1) > lim(_lib.F16, 64, 8, 128, 1, 1) > 30000
    assert lim(_lib.F16, 32, 8, 128, 16, 0) == 2 ** 31 - 1


def test_check_schedule_fails_before_the_first_forward(ekv_lib):
    """A schedule whose evicting strided chunks would hold more entries per head than the chunk tail can select among is
    refused up front (the reference has no such ceiling; here it is EKV_ERR_UNSUPPORTED, never a mid-prompt surprise):
    Mistral layout, 32 K prompt, stride 16 — budget 0.5 (16 K retained) passes, 0.6 (19.7 K) is refused; kv_policy
    'full' (nothing evicted) always passes."""
    import torch
    from easykv_b200 import plan as P
    from easykv_b200.cache import BudgetedKVCache
    c = object.__new__(BudgetedKVCache)                       # no device memory: only the shape arithmetic is exercised
    c.lib, c.dtype, c.B, c.H, c.Hkv, c.d, c.cap, c.n, c.n_phys = ekv_lib, torch.float16, 1, 32, 8, 128, 0, [0], [0]
    def run(budget, policy):
        pl = P.resolve_plan("encoding", 32768, budget, 16, 0.1, 4)
        return c.check_schedule(pl.r_idx, list(P.schedule(pl, policy, 4, False)))
    assert run(0.5, "roco")
    assert run(0.6, "full")
    with pytest.raises(NotImplementedError, match="ekv_chunk_entry_limit"):
        run(0.6, "roco")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "easykv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("test oracle", ""), f


def test_integration_doc_quotes_the_current_abi():
    """INTEGRATION.md's stub and function count follow the header (they went stale once)."""
    import re
    from easykv_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    hdr = open(os.path.join(root, "include", "easykv_b200.h")).read()
    assert int(re.search(r"ekv_abi_version\(\) == (\d+)", doc).group(1)) == _lib.ABI_VERSION
    assert int(re.search(r"#define EKV_ABI_VERSION (\d+)", hdr).group(1)) == _lib.ABI_VERSION
    words = ["zero", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten", "eleven", "twelve",
             "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen", "nineteen", "twenty"]
    n_api = len(re.findall(r"^EKV_API ", hdr, flags=re.M))
    assert n_api == len(_lib.EXPORTS)
    assert f"{words[n_api]} functions" in doc
