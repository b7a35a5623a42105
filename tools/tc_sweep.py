import sys; sys.path.insert(0, "tools"); import sweep
for name, B, H, Hkv, n in [("70B n8256 b8", 8, 64, 8, 8256), ("70B n8256 b32", 32, 64, 8, 8256), ("70B n1088 b32", 32, 64, 8, 1088),
                           ("mistral n8208 b4", 4, 32, 8, 8208), ("mistral n8208 b16", 16, 32, 8, 8208), ("mistral n1088 b32", 32, 32, 8, 1088)]:
    for c, v in ((0, 0), (0, 4), (4, 4), (8, 4)):
        sweep.run_case(name, B, H, Hkv, n, 1, "roco", cluster=c, variant=v)
