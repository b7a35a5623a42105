import sys; sys.path.insert(0, "tools"); import sweep
for B in (8, 16, 32, 64):
    for c in (-1, 1):
        sweep.run_case(f"7B b{B}", B, 32, 32, 1088, 1, "roco", cluster=c)
for c in (-1, 1, 2):
    sweep.run_case("13B n2112 b32", 32, 40, 40, 2112, 1, "roco", cluster=c)
    sweep.run_case("7B n4352 b16", 16, 32, 32, 4352, 1, "roco", cluster=c)
