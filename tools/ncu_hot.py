"""Top stall lines of one kernel of an ncu report (source page): python tools/ncu_hot.py report.ncu-rep kernel_regex [n]"""
import csv
import subprocess
import sys


def main():
    path, rx = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    sel = ["--kernel-id", ":::" + rx[3:]] if rx.startswith("id:") else ["--kernel-name", "regex:" + rx]
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    si = hdr.index("Warp Stall Sampling (All Samples)")
    ii = hdr.index("Instructions Executed")
    data = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
    # the page lists the kernel once per view; keep the first copy
    half = len(data)
    addr0 = data[0][0]
    for k in range(1, len(data)):
        if data[k][0] == addr0:
            half = k
            break
    data = data[:half]
    tot = sum(int(r[si]) for r in data)
    print(f"{len(data)} instructions, {tot} samples")
    for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][si]))[:n]:
        print(f"{idx:6d} {int(r[si]):7d} {100.0 * int(r[si]) / tot:5.1f}%  x{r[ii]:>9}  {r[1][:100]}")


if __name__ == "__main__":
    main()
