import sys; sys.path.insert(0, "tools"); import sweep
for name, B, H, Hkv, n in [("mistral n1088 b32", 32, 32, 8, 1088), ("mistral n2056 b32", 32, 32, 8, 2056), ("mistral n4104 b32", 32, 32, 8, 4104),
                           ("mistral n8208 b16", 16, 32, 8, 8208), ("mistral n8208 b4", 4, 32, 8, 8208), ("g2 n2056 b32", 32, 32, 16, 2056)]:
    for c in (-1, 0):
        sweep.run_case(name, B, H, Hkv, n, 1, "roco", cluster=c)
sweep.run_case("g2 n2056 b32", 32, 32, 16, 2056, 1, "roco", cluster=1)
