"""SASS opcode summary of libedsnet_b200.so (no GPU needed): per kernel, how many tensor-core (UTC*MMA), tensor-memory
(LDTM / STTM), TMA (UTMALDG / UTMASTG), FFMA and other instructions it contains.  Evidence for which contractions run on
tcgen05 and which on CUDA cores.

    python tools/sass_summary.py > profiles/rNN_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "edsnet-efficient-dsnet-for-video-summarization_b200", "csrc", "libedsnet_b200.so")
GROUPS = [("UTC*MMA (tcgen05.mma)", re.compile(r"^UTC[A-Z]*MMA")), ("UTCBAR (tcgen05.commit)", re.compile(r"^UTCBAR")),
          ("LDTM / STTM (tcgen05.ld / st)", re.compile(r"^(LDTM|STTM)")), ("UTMALDG / UTMASTG (TMA)", re.compile(r"^UTMA(LDG|STG)")),
          ("HMMA / IMMA (mma.sync)", re.compile(r"^(HMMA|IMMA|DMMA)")), ("FFMA", re.compile(r"^FFMA")),
          ("DFMA", re.compile(r"^DFMA")), ("SHFL", re.compile(r"^SHFL")), ("ATOM / RED", re.compile(r"^(ATOM|RED|ATOMS|ATOMG)"))]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    counts = collections.OrderedDict()
    cur, i = None, -1
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            i += 1
            cur = re.sub(r"\(.*", "", names[i]).replace("void ", "")
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            op = m.group(1)
            counts[cur]["total"] += 1
            for g, rx in GROUPS:
                if rx.match(op):
                    counts[cur][g] += 1
    print("# SASS opcode counts per kernel of libedsnet_b200.so (cuobjdump -sass, sm_100a)\n")
    print("| kernel | instr | " + " | ".join(g for g, _ in GROUPS) + " |")
    print("|---|---:|" + "---:|" * len(GROUPS))
    for k, c in sorted(counts.items(), key=lambda kv: (-kv[1][GROUPS[0][0]], kv[0])):
        print(f"| `{k[:100]}` | {c['total']} | " + " | ".join(str(c[g]) if c[g] else "" for g, _ in GROUPS) + " |")
    tc = [k for k, c in counts.items() if c[GROUPS[0][0]]]
    print(f"\n{len(tc)} of {len(counts)} kernels issue tcgen05.mma; kernels without tensor instructions are the CUDA-core "
          "(fp32 mode, integer / index, latency-bound 64 x 64) stages listed in DESIGN.md section 5.")


if __name__ == "__main__":
    main()
