"""One decode case for ncu (development aid): python tools/decode_profile.py B H Hkv n [cluster] [policy]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import sweep
B, H, Hkv, n = (int(x) for x in sys.argv[1:5])
cluster = int(sys.argv[5]) if len(sys.argv) > 5 else 0
sweep.run_case("profile", B, H, Hkv, n, 1, sys.argv[6] if len(sys.argv) > 6 else "roco", L=1, steps=2, cluster=cluster)
