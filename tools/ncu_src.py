#!/usr/bin/env python
"""Attribute per-instruction executed counts of an `ncu --page source --csv` export to CUDA source
lines via the line table of the matching object file (nvdisasm -g).
    ncu_src.py <source.csv[.gz]> <kernel-substring> <object> <mangled-substring> [npix] [top]"""
import collections, csv, gzip, io, re, subprocess, sys
src, kname, obj, mangled = sys.argv[1:5]
npix = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
sass = subprocess.run(["nvdisasm", "-g", "-c", obj], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and mangled in l][0]
cur, seq, inl = None, [], ""
for l in sass[start + 1:]:
    if l.startswith("//-----") or (l.startswith(".text.") and mangled not in l):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((m.group(2), cur))
op = gzip.open(src, "rt") if src.endswith(".gz") else open(src)
rows = list(csv.reader(op))
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
    ei, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
    if len(body) != len(seq):
        print(f"{name}: instruction count mismatch {len(body)} vs {len(seq)}; skipping")
        continue
    byline, samples, ops, tot, tots = collections.Counter(), collections.Counter(), collections.Counter(), 0, 0
    for r, (ins, loc) in zip(body, seq):
        n = int(r[ei] or 0); s = int(r[sm] or 0)
        byline[loc] += n; samples[loc] += s; tot += n; tots += s
        ops[ins.split()[0].split(".")[0] if not ins.startswith("@") else ins.split()[1].split(".")[0]] += n
    print(f"{name}: {tot} warp instructions = {tot * 32 / npix:.1f} per unit; {tots} samples")
    for loc, n in byline.most_common(top):
        print(f"  {n * 32 / npix:8.1f}  {100.0 * samples[loc] / max(tots, 1):5.1f}%smp  {loc}")
    print("  ops:", ", ".join(f"{k} {v * 32 / npix:.1f}" for k, v in ops.most_common(24)))
    break
