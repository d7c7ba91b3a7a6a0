"""Small driver for ncu: a few SDF fwd+grad / bwd passes on 65 536 points in the chosen precision."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import factored_neus_b200 as fn
from factored_neus_b200 import ops
syn = fn.synthetic
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
ops.set_precision(prec)
st = syn.scene_states(seed=4)
sdf = fn.SDFNetwork(**syn.SDF_CONF); sdf.load_state_dict(st["sdf"]); sdf = sdf.cuda()
x = (torch.rand(N, 3, device="cuda") * 2 - 1)
for it in range(3):
    s, f, n = sdf.value_feature_normal(x)
    (s.sum() + f.sum() * 0.01 + n.sum()).backward()
torch.cuda.synchronize()
print("done")
