"""One pileup + one FA forward (used under compute-sanitizer)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clair3_b200 import synth
from clair3_b200.model import Clair3_P, Clair3_F
which = sys.argv[1] if len(sys.argv) > 1 else "p"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
if which == "p":
    sd = synth.pileup_state_dict(False, seed=0); x = synth.pileup_inputs(B, seed=1)
    m = Clair3_P(False, True, 18)
else:
    sd = synth.fa_state_dict(True, channels=8, seed=0); x = synth.fa_inputs(B, seed=1)
    m = Clair3_F(True, True, 8)
reps = 1
for a in sys.argv[3:]:
    k, v = a.split("=")
    if k == "reps":
        reps = int(v)          # extra forwards (ncu captures skip the first one)
    else:
        m.set_option(k, int(v))
m.to(torch.device("cuda")); m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
for _ in range(reps):
    y = m(torch.from_numpy(x).cuda()).cpu().numpy()
print("ok", y.shape, float(y.sum()))
