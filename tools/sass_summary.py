#!/usr/bin/env python
"""Per-kernel count of the SASS mnemonics that prove (or disprove) Blackwell-native code in libsnb200.so:
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk: the TMA unit's 1-D bulk copy), UTMALDG (cp.async.bulk.tensor:
tiled TMA - not used), SYNCS (mbarrier), UTCBAR (tcgen05.commit).  usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hobot_stereonet_b200", "lib", "libsnb200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "UBLKCP", "UTMALDG", "SYNCS", "UTCBAR", "HMMA", "FFMA", "STG", "LDG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("snb::", "")
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in KEYS:
                per[cur][op] += 1
            per[cur]["_all"] += 1
    arch = re.findall(r"arch = (sm_\w+)", out)
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({', '.join(sorted(set(arch)))})")
    print(f"{'kernel':58s} " + " ".join(f"{k:>8s}" for k in KEYS) + f" {'instrs':>8s}")
    tot = collections.Counter()
    for k, c in per.items():
        print(f"{k[:58]:58s} " + " ".join(f"{c[x]:8d}" for x in KEYS) + f" {c['_all']:8d}")
        tot.update(c)
    print(f"{'TOTAL':58s} " + " ".join(f"{tot[x]:8d}" for x in KEYS) + f" {tot['_all']:8d}")


if __name__ == "__main__":
    main()
