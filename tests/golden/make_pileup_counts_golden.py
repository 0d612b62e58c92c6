"""Mints tests/golden/pileup_counts.npz: the pileup feature counter's oracle (oracle/pileup_oracle.c) on two small seeded record
sets, INPUT RECORDS INCLUDED, so that the committed vectors pin both the oracle and (on a GPU) the CUDA counter against drift.
These are NOT reference-minted (libclair3 / htslib cannot be built in this image - see the oracle's header): they freeze the
hand-verified restatement.   python tests/golden/make_pileup_counts_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from clair3_b200 import synth_reads as sr  # noqa: E402
from oracle import pileup_oracle as po  # noqa: E402

out = {}
for tag, kw, cnt in (("a", dict(region_len=600, depth=25, read_len=300, seed=71, wild=True, indel_rate=0.08, n_rate=0.01), dict(gvcf=True)),
                     ("b", dict(region_len=900, depth=12, read_len=150, seed=72, gaps=[(1400, 1470)], origin=1000), dict(call_ht=True, max_indel_length=5))):
    rec, ref, rs = sr.random_alignment(**kw)
    start = kw.get("origin", 1000)
    end = start + kw["region_len"]
    r = po.clair3_pileup(rec, start, end, ref, rs, alt_info=True, **cnt)
    for k, v in rec.items():
        out["%s_rec_%s" % (tag, k)] = v
    out["%s_ref" % tag] = np.frombuffer(ref.encode(), dtype=np.uint8)
    out["%s_meta" % tag] = np.array([start, end, rs, int(cnt.get("gvcf", False)), int(cnt.get("call_ht", False)), cnt.get("max_indel_length", 50)], np.int64)
    for k in ("matrix", "major", "stats", "cand_cols", "cand_ok", "pos_ref_count", "pos_total_count"):
        out["%s_%s" % (tag, k)] = r[k]
    out["%s_alt_info" % tag] = np.array(r["alt_info"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pileup_counts.npz"), **out)
print({k: v.shape for k, v in out.items() if k.endswith(("matrix", "alt_info"))})
