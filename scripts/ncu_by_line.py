#!/usr/bin/env python3
"""Join an ncu source-page CSV (SASS view) with nvdisasm line info and aggregate per CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <kernel-section-substring> [top_n]
Prints executed warp-instructions and stall samples per source line of m2s_kernels.cu.
"""
import csv, os, re, subprocess, sys, tempfile
rep, sect = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "mesh2splat_b200", "libm2s.so")], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, "m2s_kernels.sm_100a.cubin")
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# offsets -> line within the chosen section
off2line, cur, insec = {}, None, False
for l in sass:
    if l.startswith(".text.") or "\t.section\t.text." in l:
        insec = sect in l
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
kflt = "raster_kernel" if "raster" in sect else ("fragment_kernel" if "fragment" in sect else sect)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kflt], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(rows[2][ia], 16)
agg = {}
tot_ex = tot_s = 0
seen_hdr = 0
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break  # the same kernel listed again
    if len(r) <= iex or r[ia] == 'Address':
        continue
    off = int(r[ia], 16) - base
    line = off2line.get(off, ("?", -1))
    ex, sm = int(r[iex] or 0), int(r[isamp] or 0)
    a = agg.setdefault(line, [0, 0]); a[0] += ex; a[1] += sm
    tot_ex += ex; tot_s += sm
src = open(os.path.join(root, "mesh2splat_b200", "csrc", "m2s_kernels.cu")).read().splitlines()
shift = int(os.environ.get('M2S_LINE_SHIFT', '0'))  # the library was built from an older source: lines moved by this much
src = [''] * max(0, -shift) + src[max(0, shift):] if shift else src
print(f"total warp-instr {tot_ex}  samples {tot_s}")
for (f, ln), (ex, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = src[ln - 1].strip()[:90] if f == "m2s_kernels.cu" and 0 < ln <= len(src) else f
    print(f"{f}:{ln:4d} ex {ex:9d} ({100*ex/tot_ex:4.1f}%) samp {sm:7d} ({100*sm/max(tot_s,1):4.1f}%)  {text}")

# coarse regions (line ranges of m2s_kernels.cu given as extra args "name:lo-hi")
regs = [a for a in sys.argv[4:] if ":" in a]
if regs:
    print("--- regions")
    for r in regs:
        name, rng = r.split(":"); lo, hi = map(int, rng.split("-"))
        ex = sum(v[0] for (f, ln), v in agg.items() if f == "m2s_kernels.cu" and lo <= ln <= hi)
        sm = sum(v[1] for (f, ln), v in agg.items() if f == "m2s_kernels.cu" and lo <= ln <= hi)
        print(f"{name:14s} ex {ex:9d} ({100*ex/tot_ex:4.1f}%)  samples {sm:6d} ({100*sm/max(tot_s,1):4.1f}%)")
    ex = sum(v[0] for (f, ln), v in agg.items() if f != "m2s_kernels.cu"); sm = sum(v[1] for (f, ln), v in agg.items() if f != "m2s_kernels.cu")
    print(f"{'other files':14s} ex {ex:9d} ({100*ex/tot_ex:4.1f}%)  samples {sm:6d} ({100*sm/max(tot_s,1):4.1f}%)")
