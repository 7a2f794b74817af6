#!/usr/bin/env python
"""Attribute ncu per-instruction counts (source page CSV) to CUDA source lines using the
line table of the same cubin (nvdisasm -g). Usage:
    ncu_lines.py <report.ncu-rep> <kernel-substring> <cubin> <mangled-substring> [top]
"""
import collections, csv, io, re, subprocess, sys

rep, kname, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and mangled in l][0]
cur, seq = None, []
for l in sass[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((m.group(2), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, name = [], None
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name":
        name = r[1]
    if r and r[0] == "Address":
        blocks.append((name, i))
for bi, (name, i) in enumerate(blocks):
    if kname not in (name or ""):
        continue
    j = blocks[bi + 1][1] - 1 if bi + 1 < len(blocks) else len(rows)
    hdr, body = rows[i], [r for r in rows[i + 1:j] if len(r) > 10]
    ei, sm, si = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    if len(body) != len(seq):
        print(f"{name}: instruction count mismatch {len(body)} vs {len(seq)}; skipping")
        continue
    byline, samples, ops, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0
    for k, r in enumerate(body):
        n, s = int(r[ei]), int(r[sm])
        byline[seq[k][1]] += n
        samples[seq[k][1]] += s
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[si])
        ops[m.group(2) if m else "?"] += n
        tot += n
    print(name, "total warp instr", tot, "static", len(body))
    byfile = collections.Counter()
    for k, n in byline.items():
        byfile[k[0] if k else None] += n
    print({k: f"{v / tot * 100:.1f}%" for k, v in byfile.most_common(8)})
    print(" ".join(f"{o}:{n / tot * 100:.1f}%" for o, n in ops.most_common(24)))
    for k, n in byline.most_common(top):
        print(f"  {k}  {n / tot * 100:.2f}%  samples {samples[k]}")
