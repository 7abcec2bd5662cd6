import sys
sys.path.insert(0, ".")
import torch
from deephisto_b200 import ops
st = ops.CoverState(40000, 40000, 224, 16, 2, 64, seed=0)
for _ in range(6):
    c, n = st.next_group(16)
torch.cuda.synchronize()
print("ok", n.tolist())
