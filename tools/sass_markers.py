"""Blackwell evidence: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, bulk/tensor TMA -> UBLKCP/UTMALDG/UTMASTG,
tcgen05.commit -> UTCBAR, mbarrier -> SYNCS), from `cuobjdump -sass` of the built library.
usage: python tools/sass_markers.py [path/to/libmcnerf.so] > profiles/r02_sass_markers.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mc_nerf_b200", "csrc", "libmcnerf.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
MARK = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "UTCATOM", "SYNCS", "HMMA", "HGMMA",
        "UCGABAR", "CCTL", "MEMBAR", "ATOMG", "REDG"]
counts, total, kern, arch = collections.OrderedDict(), collections.Counter(), None, set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        counts.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if kern is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[kern]["_insts"] += 1
        if op in MARK:
            counts[kern][op] += 1
            total[op] += 1
print(f"SASS markers of {os.path.relpath(so, ROOT)} (arch {', '.join(sorted(arch))}); columns = instruction counts per kernel")
cols = [c for c in MARK if total[c]]
print(f"{'kernel':58s} {'insts':>7s} " + " ".join(f"{c:>8s}" for c in cols))
for k, c in counts.items():
    if not any(c[x] for x in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "UTMALDG", "SYNCS")):
        continue
    print(f"{k[:58]:58s} {c['_insts']:7d} " + " ".join(f"{c[x]:8d}" for x in cols))
print(f"{'TOTAL (all ' + str(len(counts)) + ' kernels)':58s} {sum(c['_insts'] for c in counts.values()):7d} "
      + " ".join(f"{total[x]:8d}" for x in cols))
print("\nNotes: UTCHMMA = tcgen05.mma kind::f16 (bf16 operands, fp32 accumulate in TMEM); LDTM = tcgen05.ld; UTCBAR = tcgen05.commit;\n"
      "UBLKCP = cp.async.bulk (1-D bulk copies on the TMA engine: the packed weight / tile images are contiguous, so no tensor\n"
      "maps are needed - cp.async.bulk.tensor (UTMALDG) was measured and brought nothing, DESIGN.md section 4.2); no legacy HMMA\n"
      "(mma.sync) and no HGMMA (wgmma) anywhere.")
