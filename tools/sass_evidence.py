#!/usr/bin/env python
"""Counts the Blackwell-native SASS mnemonics per kernel of libglare_b200.so (cuobjdump -sass; runs without a GPU).

    python tools/sass_evidence.py > profiles/sass_evidence.txt

UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tiled load / store, UBLKCP = bulk copy, UTCBAR = tcgen05.commit,
SYNCS = mbarrier operations (/opt/skills/guides/B200_PROFILING.md "What proves a Blackwell-native kernel")."""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "glare_b200", "libglare_b200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|UTCBAR|SYNCS|HMMA|MUFU\.EX2|F2FP)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
        elif cur:
            for mm in PAT.finditer(line):
                cnt[cur][mm.group(1)] += 1
    names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonics per kernel of glare_b200/libglare_b200.so (sm_100a), static instruction counts; kernels without any of them omitted")
    agg = collections.OrderedDict()
    for raw, name in zip(cnt, names):
        if cnt[raw]:
            key = re.sub(r"\(.*", "", name).replace("void ", "")
            agg[key] = cnt[raw]
    for k, c in sorted(agg.items()):
        print("%-64s %s" % (k[:64], "  ".join("%s=%d" % kv for kv in sorted(c.items()))))


if __name__ == "__main__":
    main()
