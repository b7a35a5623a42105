"""One chunk case for ncu (development aid): python tools/chunk_profile.py B H Hkv n stride policy"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import sweep
B, H, Hkv, n, q = (int(x) for x in sys.argv[1:6])
sweep.run_case("profile", B, H, Hkv, n, q, sys.argv[6] if len(sys.argv) > 6 else "roco", L=1, steps=2)
