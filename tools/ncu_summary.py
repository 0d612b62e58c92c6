"""Summarise an `ncu --page raw --csv` export into the few metrics DESIGN.md / bench.py cite.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > X_raw.csv ; python tools/ncu_summary.py X_raw.csv > profiles/X.md
"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (math) active % of active cycles"),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tensor-core unit busy % (incl. operand fetch)"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem->TC operand wavefronts % of peak"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe % of active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(k[1] for k in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("<unnamed>::", "").replace("void ", "")
        cells = []
        for key, _ in KEYS:
            if key in idx:
                cells.append("%s %s" % (r[idx[key]], units[idx[key]]))
            else:
                cells.append("n/a")
        print("| `%s` | " % name + " | ".join(cells) + " |")


def traffic(path, workload, batch, out_json):
    """Append DRAM bytes per launch (read + write) of each captured kernel to profiles/traffic.json (bench.py's roofline.traffic)."""
    import json
    import os
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ent = {}
    nconv = 0
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        byts = sum(float(r[idx[k]]) * scale[units[idx[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        key = None
        if "lstm_pair_kernel" in name:
            key = "lstm2" if "1>" in name.replace("(bool)", "") else "lstm1"
        elif "lstm_tc_kernel" in name:
            key = "lstm2" if ", 1, " in name.replace("(bool)", "") or ",1," in name else "lstm1"
        elif "lstm_pair_kernel" in name:
            key = "lstm2" if "1>" in name.replace("(bool)", "") else "lstm1"
        elif "proj2_kernel" in name:
            key = "proj2"
        elif "tail_kernel" in name:
            key = "tail"
        elif "ingest_pileup_tc_kernel" in name or "ingest_fa_tc_kernel" in name:
            key = "ingest"
        elif "igemm_kernel" in name:
            key = "proj2" if ("1, 1>" in name.replace("(bool)", "").replace("(IgemmEpilogue)", "")) else "l4"
        elif "pconv_kernel" in name:
            key = "conv%d" % nconv
            nconv += 1
        elif "heads_kernel" in name:
            key = "heads"
        elif "spp_tc_kernel" in name:
            key = "spp"
        if key and key not in ent:
            ent[key] = {"batch": batch, "dram_bytes_per_launch": byts, "kernel": name.replace("<unnamed>::", "")[:80],
                        "duration_us_under_ncu": float(r[idx["gpu__time_duration.sum"]])}
    cur = {}
    if os.path.exists(out_json):
        cur = json.load(open(out_json))
    cur[workload] = ent
    json.dump(cur, open(out_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if len(sys.argv) >= 5 and sys.argv[1] == "--traffic":
        traffic(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5] if len(sys.argv) > 5 else "profiles/traffic.json")
    else:
        main(sys.argv[1])
