"""cuobjdump -sass clair3_b200/libclair3b200.so | python tools/sass_evidence.py > profiles/r1_sass_evidence.md

Static per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMA path (B200_PROFILING.md): UTCHMMA = tcgen05.mma,
UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st, SYNCS = mbarrier, LDGSTS = cp.async."""
import collections
import re
import subprocess
import sys

MN = ["UTCHMMA", "UTCBAR", "UBLKCP", "LDTM", "STTM", "SYNCS", "ELECT", "MUFU.TANH", "LDGSTS", "R2UR"]


def main():
    cur, counts = None, collections.OrderedDict()
    for line in sys.stdin:
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in MN:
                if op.startswith(k):
                    counts[cur][k] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence (`cuobjdump -sass clair3_b200/libclair3b200.so`, sm_100a; tools/sass_evidence.py)\n")
    print("Static instruction counts per kernel: `UTCHMMA` = tcgen05.mma (`gdesc,gdesc` = both operands in shared memory; the LSTM2")
    print("kernel also has the `tmem,gdesc` A-in-TMEM form), `UTCBAR` = tcgen05.commit, `UBLKCP` = cp.async.bulk (TMA engine),")
    print("`LDTM`/`STTM` = tcgen05.ld/st, `SYNCS` = mbarrier ops, `LDGSTS` = cp.async (only the fp32 debug heads kernel's")
    print("weight ring uses it).  The CTA-pair LSTM kernel (`lstm_pair_kernel`) issues `UTCHMMA.2CTA`, commits with")
    print("`UTCBAR.2CTA.MULTICAST` and allocates TMEM with `UTCATOMSWS.2CTA` (cuobjdump -sass clair3_b200/csrc/_obj/lstm2x_tc.o |")
    print("grep -o 'UTC[A-Z0-9_.]*' | sort | uniq -c: 2 x `UTCHMMA.2CTA`, 3 x `UTCBAR.2CTA.MULTICAST`, 4 x `UTCATOMSWS.2CTA.FIND_AND_SET.ALIGN`).\n")
    print("| kernel | instrs | " + " | ".join(MN) + " |")
    print("|---|---|" + "---|" * len(MN))
    for (k, c), name in zip(counts.items(), names):
        if c["_total"] == 0 or not any(c[m] for m in MN if m != "R2UR"):
            continue
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name))
        print("| `%s` | %d | " % (name[:64], c["_total"]) + " | ".join(str(c[m]) for m in MN) + " |")


if __name__ == "__main__":
    main()
