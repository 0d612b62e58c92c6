"""Version-agnostic multi-stream consistency check (used to bisect): python tools/stress_min.py <repo_root> [tile]"""
import sys, os
root = os.path.abspath(sys.argv[1]); sys.path.insert(0, root)
import numpy as np, torch
from clair3_b200 import synth
from clair3_b200.model import Clair3_P
sd = synth.pileup_state_dict(False, seed=0)
xs = [synth.pileup_inputs(1024, seed=100 + i) for i in range(8)]
m = Clair3_P(add_indel_length=False, predict=True, input_channels=18)
if len(sys.argv) > 2: m.set_option("lstm_tile", int(sys.argv[2]))
m.to(torch.device("cuda")); m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
xd = [torch.from_numpy(x).cuda() for x in xs]
ref = [m(x).cpu().numpy() for x in xd]
streams = [torch.cuda.Stream() for _ in range(8)]
worst, nbad = 0.0, 0
for rep in range(6):
    outs = [None] * 8
    for i in range(8):
        with torch.cuda.stream(streams[i]):
            outs[i] = m(xd[i])
    torch.cuda.synchronize()
    for i in range(8):
        d = float(np.abs(outs[i].cpu().numpy() - ref[i]).max()); worst = max(worst, d); nbad += d > 1e-4
print(os.path.basename(root), "worst diff", worst, "bad forwards", nbad, "of 48", flush=True)
