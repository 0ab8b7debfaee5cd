#!/usr/bin/env python
"""SASS evidence per kernel of liblstep_b200.so (cuobjdump -sass): which Blackwell-specific instructions each hot kernel
contains. Regenerate with `python profiles/sass_evidence.py > profiles/r02_sass_evidence.txt` (no GPU needed).

  UTCHMMA / UTCQMMA   tcgen05.mma            LDTM / STTM        tcgen05.ld / st (TMEM)
  UTCBAR              tcgen05.commit         UTCATOMSWS         tcgen05.alloc / dealloc
  UBLKCP              cp.async.bulk (TMA bulk copy engine)      UTMALDG / UTMASTG  cp.async.bulk.tensor
  STAS                st.async (DSMEM store completing an mbarrier)   SYNCS   mbarrier ops
  FFMA2               packed fp32x2 FMA (sm_100)                HMMA    legacy mma.sync
  ACQBULK / CCTL ... not listed
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "l-step_b200", "liblstep_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "STAS", "SYNCS", "FFMA2",
        "FFMA", "HMMA", "ATOMG", "REDG", "RED", "LDG", "STG", "MUFU", "F2FP", "SHFL", "BAR", "ACQBULK", "UCGABAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("lstep::", "")
            name = re.sub(r"\(.*", "", name)
            cur = name
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
                    total[k] += 1
                    break
    arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    print("library:", os.path.relpath(LIB, ROOT), "| ELF:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))))
    print("totals:", ", ".join(f"{k} {v}" for k, v in total.items() if k in ("UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "STAS", "SYNCS", "FFMA2", "HMMA")))
    print()
    for name, c in counts.items():
        if c["_all"] < 50:
            continue
        keys = " ".join(f"{k}={c[k]}" for k in KEYS if c[k] and k not in ("FFMA", "LDG", "STG", "SHFL", "BAR"))
        print(f"{name[:70]:70s} {c['_all']:6d} instr | FFMA={c['FFMA']} LDG={c['LDG']} STG={c['STG']} | {keys}")


if __name__ == "__main__":
    main()
