"""Executed-instruction mix of one kernel of an ncu report (source page), weighted by execution count:
    python tools/ncu_opmix.py report.ncu-rep kernel_regex|id:N [n]   (id:N = N-th launch, 1-based)
Also prints the share of mbarrier try_wait spin loops (SYNCS ... TRYWAIT) in the executed instructions."""
import collections
import csv
import subprocess
import sys


def main():
    path, rx = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    sel = ["--kernel-id", ":::" + rx[3:]] if rx.startswith("id:") else ["--kernel-name", "regex:" + rx]
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    si = hdr.index("Warp Stall Sampling (All Samples)")
    ii = hdr.index("Instructions Executed")
    data = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
    addr0 = data[0][0]
    for k in range(1, len(data)):
        if data[k][0] == addr0:
            data = data[:k]
            break
    cnt, samp = collections.Counter(), collections.Counter()
    tot = 0
    trywait = 0
    for r in data:
        parts = r[1].split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        c = int(r[ii])
        cnt[op] += c
        samp[op] += int(r[si])
        tot += c
        if "TRYWAIT" in r[1]:
            trywait += c
    print(f"{len(data)} static instructions, {tot} executed (warp level), {sum(samp.values())} samples")
    print(f"mbarrier try_wait executions: {trywait} (x ~9 instructions per spin = {900.0 * trywait / tot:.1f}% of all)")
    for op, c in cnt.most_common(n):
        print(f"  {op:10s} {c:12d} {100.0 * c / tot:5.1f}%   samples {samp[op]}")


if __name__ == "__main__":
    main()
