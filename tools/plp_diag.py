"""GPU diagnosis / profiling driver of the pileup feature counter: python tools/plp_diag.py diag | prof [region]
diag: per-key mismatch report against the C oracle on the known-answer and a few random cases (which features, where).
prof: a few counts over a bench-shaped region (for ncu)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def report(tag, got, want):
    bad = False
    for k in ("major", "matrix", "stats", "cand_cols", "cand_ok"):
        g, w = got[k], want[k]
        if g.shape != w.shape:
            print("%s: %s SHAPE gpu %s oracle %s" % (tag, k, g.shape, w.shape))
            bad = True
            n = min(len(g), len(w))
            g, w = g[:n], w[:n]
        d = np.argwhere(g != w)
        if len(d):
            bad = True
            print("%s: %s %d mismatches; first %s" % (tag, k, len(d), [(x.tolist(), int(g[tuple(x)]), int(w[tuple(x)])) for x in d[:6]]))
            if g.ndim == 2:
                print("   per column of %s: %s" % (k, np.bincount(d[:, 1], minlength=g.shape[1]).tolist()))
    print("%s: %s (n_cols %d, candidates %d)" % (tag, "MISMATCH" if bad else "bit-exact", len(want["major"]), len(want["cand_cols"])))
    return not bad


def main():
    import torch
    from clair3_b200 import pileup_counts as pc, synth_reads as sr
    from oracle import pileup_oracle as po
    mode = sys.argv[1] if len(sys.argv) > 1 else "diag"
    ctr = pc.PileupCounter(0)
    if mode == "prof":
        region = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
        rec, ref, rs = sr.random_alignment(region, depth=40, read_len=8000, seed=5, indel_rate=0.04, origin=10000)
        d = pc.BamRecords.from_dict(rec).to_device(torch.device("cuda:0"), ref)
        reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
        ms = []
        for _ in range(reps):
            ctr.count(d, 10000, 10000 + region, None, rs)
            torch.cuda.synchronize()
            ms.append(ctr.last_ms()[0])
        nc, nk = ctr.sizes()
        print("kernel %s, region %d: count %s ms (median %.3f), %d columns, %d candidates"
              % ("shipped", region, ["%.3f" % m for m in ms[:8]], float(np.median(ms[1:] or ms)), nc, nk))
        return
    from test_pileup_oracle import case_indels, case_quirks
    ok = True
    rec, ref, *_ = case_indels()
    ok &= report("known_indels", ctr.count(rec, 0, 20, ref, 0, call_ht=True).fetch(), po.clair3_pileup(rec, 0, 20, ref, 0, call_ht=True))
    rec, ref, *_ = case_quirks()
    ok &= report("known_quirks", ctr.count(rec, 0, 12, ref, 0, call_ht=True, min_depth=1).fetch(),
                 po.clair3_pileup(rec, 0, 12, ref, 0, call_ht=True, min_depth=1))
    for seed, (width, depth, rl, wild) in enumerate([(700, 4, 60, True), (256, 12, 150, False), (1025, 35, 400, True), (5000, 30, 2000, False)]):
        rec, ref, rs = sr.random_alignment(width, depth=depth, read_len=rl, seed=200 + seed, wild=wild, indel_rate=0.08, n_rate=0.01)
        ok &= report("random_%d" % seed, ctr.count(rec, 1000, 1000 + width, ref, rs).fetch(), po.clair3_pileup(rec, 1000, 1000 + width, ref, rs))
    rec, ref, rs = sr.random_alignment(3000, depth=40, read_len=600, seed=61, indel_rate=0.06, n_rate=0.01)
    want = po.clair3_pileup(rec, 1000, 4000, ref, rs, alt_info=True)
    got = ctr.count(rec, 1000, 4000, ref, rs, alt_info=True).fetch()
    ok &= report("alt_info_counts", got, want)
    text = ctr.alt_info_strings(got)
    bad = [(a, b) for a, b in zip(text, want["alt_info"]) if a != b]
    print("alt_info text: %d strings, %d differ%s" % (len(text), len(bad) + abs(len(text) - len(want["alt_info"])), (" first: %r vs %r" % bad[0]) if bad else ""))
    ok &= not bad and len(text) == len(want["alt_info"])
    print("DIAG", "ALL BIT-EXACT" if ok else "HAS MISMATCHES")


if __name__ == "__main__":
    main()
