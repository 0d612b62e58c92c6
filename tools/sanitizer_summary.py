"""Summarise compute-sanitizer logs (gpurun_out/<R>_san_*.log) into a markdown table: python tools/sanitizer_summary.py r2a r2b > profiles/r2_sanitizer.md"""
import glob
import os
import re
import sys

print("# compute-sanitizer evidence\n")
print("One small forward of each network per run (`tools/sanitize_case.py`; `tools/sanitize_round.sh`, `tools/round2_f.sh`), B200, "
      "`compute-sanitizer --tool <tool>`.  `ok (B, out)` is the script's own success line (the forward completed and returned).\n")
print("| log | tool | case | result |")
print("|---|---|---|---|")
for r in sys.argv[1:]:
    for path in sorted(glob.glob("gpurun_out/%s_san_*.log" % r)):
        txt = open(path, errors="replace").read()
        name = os.path.basename(path)[:-4]
        tool = next((t for t in ("memcheck", "racecheck", "synccheck", "initcheck") if t in name), "synccheck" if "sync" in name else "?")
        summ = re.findall(r"(ERROR SUMMARY: \d+ errors?|RACECHECK SUMMARY: [^\n]*)", txt)
        ok = re.findall(r"ok \(\d+, \d+\)", txt)
        print("| `%s` | %s | %s | %s %s |" % (name, tool, name.replace(r + "_san_", ""), "; ".join(dict.fromkeys(summ)) or "no summary", ok[0] if ok else "(forward did not finish)"))
