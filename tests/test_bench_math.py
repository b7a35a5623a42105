"""The roofline numerators bench.py reports are SURVEY §8(d)'s per-unit figures (CPU test: no GPU, no library call)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_algorithmic_bytes_follow_survey_8d():
    import bench
    w = bench.WORKLOADS["c2"]
    # configs[1] in its evicting form: n = 1088 retained + 1, Hkv = 32, d = 128, fp16, roco (S, SQ, C read + written)
    kv = 2 * 32 * 1089 * 128 * 2
    new_rows = 2 * 32 * 128 * 2
    q_out = 2 * 32 * 128 * 2
    state = 6 * 32 * 1089 * 4
    victims = 32 * 4
    assert kv == 17_842_176 and new_rows == 16_384 and q_out == 16_384 and state == 836_352 and victims == 128
    assert bench.bytes_alg(w, 1) == kv + new_rows + q_out + state + victims == 18_711_424
    assert bench.bytes_alg(w, 64) == 1_197_531_136                          # one fused layer launch of the headline line
    assert bench.bytes_alg(w) == bench.bytes_alg(w, w["B"])


def test_every_workload_names_a_baseline_shape():
    import bench
    for name, w in bench.WORKLOADS.items():
        assert w["H"] % w["Hkv"] == 0 and w["H"] // w["Hkv"] in (1, 2, 4, 8), name
        assert w["kind"] in ("decode", "chunk") and w["policy"] in bench.A_POL, name
        assert bench.bytes_alg(w) > 0 and bench.flops_alg(w) > 0, name
        if w["kind"] == "chunk":
            assert w.get("stride", 1) > 1, name
    # the tensor-bound entry: 70B stride 64 (arithmetic intensity g * stride = 512 flop per K / V byte)
    c5c = bench.WORKLOADS["c5_chunk"]
    assert bench.flops_alg(c5c) / bench.bytes_alg(c5c) > 400
