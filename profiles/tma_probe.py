"""Probe: which patch origins does the TMA gather accept? (each case in its own process: a fault kills the context)"""
import subprocess, sys
CASE = r'''
import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from deephisto_b200 import ops
y, x, ps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ops.set_gather_variant("tma")
s = ops.DeviceSlide.synthetic(2048, 2048, 0)
c = torch.tensor([[y, x]] * 4, dtype=torch.int32, device="cuda")
out = ops.gather_normalize(s, c, ps)
torch.cuda.synchronize()
ops.set_gather_variant("direct")
ref = ops.gather_normalize(s, c, ps)
print("OK", bool(torch.equal(out, ref)))
'''
for y, x, ps in [(0, 0, 224), (0, 16, 224), (0, 224, 224), (0, 1, 224), (0, 4, 224), (0, 5, 224), (3, 0, 224), (0, 100, 224), (0, 1, 64), (0, 16, 64)]:
    r = subprocess.run([sys.executable, "-c", CASE, str(y), str(x), str(ps)], capture_output=True, text=True)
    tail = (r.stdout.strip().splitlines() or ["-"])[-1]
    err = [l for l in r.stderr.splitlines() if "Error" in l or "error" in l][-1:] if r.returncode else []
    print((y, x, ps), "rc", r.returncode, tail, err)
