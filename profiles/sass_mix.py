#!/usr/bin/env python
"""Dynamic instruction mix of the config 3 sweep kernels from an `ncu --set full --import-source on`
capture:   python profiles/sass_mix.py gpurun_out/r02_sweep.ncu-rep profiles/r02_sass_mix.json
Sums the per-instruction 'Instructions Executed' column of the source page by pipe class and
normalises to warp instructions per warp-level site group (32 threads x 4 replica words = 4096 flip
attempts; one colour phase of config 3 has 32768 of them).  The source page's total is scaled to
smsp__inst_executed.sum of the raw page (it counts every instruction once per collection pass)."""
import csv
import io
import json
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
groups = float(sys.argv[3]) if len(sys.argv) > 3 else 32768.0


def page(name):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(page("raw"))))
hdr = raw[0]
ix = {h: i for i, h in enumerate(hdr)}
executed = {}
for r in raw[2:]:
    executed.setdefault(r[ix["Kernel Name"]], float(r[ix["smsp__inst_executed.sum"]].replace(",", "")))


def klass(op):
    if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"):
        return "mul_wide_or_hi"
    if op.startswith("IMAD") or op.startswith("FFMA") or op.startswith("FMUL") or op.startswith("FADD"):
        return "imad"
    # (U-prefixed opcodes run on the uniform datapath and count as "other")
    if re.match(r"(LOP3|IADD3|IADD|SHF|SEL|ISETP|LEA|PLOP3|FLO|BREV|POPC|IABS|PRMT|MOV|SGXT|VIADD|LOP|SHL|SHR|VABSDIFF|IMNMX|VIMNMX)", op):
        return "alu"
    return "other"


mix = {}
cur, counts = None, None
sections = []
for r in csv.reader(io.StringIO(page("source"))):
    if r and r[0] == "Kernel Name":
        cur = r[1]
        counts = {"alu": 0.0, "imad": 0.0, "mul_wide_or_hi": 0.0, "other": 0.0}
        sections.append((cur, counts))
        col = None
        continue
    if r and r[0] == "Address":
        col = r.index("Instructions Executed")
        continue
    if cur is None or len(r) < 6 or not r[0].startswith("0x"):
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
    if not m:
        continue
    counts[klass(m.group(1))] += float(r[col])
seen = set()
for name, counts in sections:
    if name in seen:
        continue
    seen.add(name)
    acc = bool(re.search(r"\(int\)4, \(bool\)1", name))
    total = sum(counts.values())
    scale = executed.get(name, total) / total if total else 1.0
    key = "acc" if acc else "plain"
    mix[key] = {k: round(v * scale / groups, 1) for k, v in counts.items()}
    mix[key]["total"] = round(sum(v * scale / groups for v in counts.values()), 1)
mix["_note"] = ("warp instructions per warp-level site group (32 threads x 4 replica words = 4096 flip attempts), "
                "dynamic counts of one launch of k_sweep_rows<3,1,6,7,4,ACC,0> on BASELINE config 3 from the "
                "per-instruction 'Instructions Executed' column of the ncu source page (profiles/sass_mix.py); alu = "
                "LOP3/IADD3/SHF/SEL/ISETP/LEA..., mul_wide_or_hi = IMAD.WIDE / IMAD.HI, imad = other IMAD, other = "
                "loads, stores, LDC/LDCU, branches, barriers")
json.dump(mix, open(out, "w"), indent=1)
print(json.dumps(mix, indent=1))
