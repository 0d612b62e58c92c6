"""Clock stamps of epilogue thread 0 (CTA 0) of the CTA-pair LSTM kernels: steps 8..13, per phase
[before acc_full wait, after wait, after the TMEM reads / stage release, after the h store].  python tools/pair_trace.py [batch]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clair3_b200 import synth
from clair3_b200._ffi import check, ffi, lib
from clair3_b200.model import Clair3_P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sd = synth.pileup_state_dict(False, seed=0)
x = torch.from_numpy(synth.pileup_inputs(B, seed=1)).cuda()
m = Clair3_P(False, True, 18)
for kv in sys.argv[2:]:
    k, v = kv.split("="); m.set_option(k, int(v))
m.set_option("lstm1_impl", 1); m.set_option("lstm2_impl", 1)
m.to(torch.device("cuda")); m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
for _ in range(3):
    m(x)
m.set_option("lstm_trace", 1)
m(x); torch.cuda.synchronize()
out = np.zeros(264, dtype=np.int64)
check(lib().c3b_debug_lstm_trace(m._handle, ffi.cast("int64_t *", out.ctypes.data)))
for layer, phases in ((0, 4), (1, 5)):
    t = out[layer * 132:(layer + 1) * 132][:6 * phases * 4].reshape(6, phases, 4)
    base = t[0, 0, 0]
    print("layer %d (LSTM%d): per phase [wait, tmem+act(i,g,f)+release, act(o)+cell+store] cycles; step period" % (layer, layer + 1))
    for s in range(6):
        row = ["%5d/%5d/%5d" % (t[s, p, 1] - t[s, p, 0], t[s, p, 2] - t[s, p, 1], t[s, p, 3] - t[s, p, 2]) for p in range(phases)]
        per = (t[s + 1, 0, 0] - t[s, 0, 0]) if s + 1 < 6 else 0
        print("  step %2d: %s   period %6d" % (8 + s, "  ".join(row), per))
