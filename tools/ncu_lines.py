"""Aggregate ncu warp-stall samples of one kernel by CUDA source line.

usage: python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]
Joins `ncu --page source --csv` (per-SASS-instruction samples) with `nvdisasm -g` line annotations.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr2line, cur, infunc = {}, None, False
for ln in dis:
    if ln.startswith(".text."):
        infunc = kname in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if "inlined at" not in ln:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and infunc:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isamp, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = rows[hi + 1:]
base = min(int(r[ia], 16) for r in body if r and r[ia])
by_line, by_line_stall, tot = defaultdict(int), defaultdict(lambda: defaultdict(int)), 0
for r in body:
    if not r or not r[ia]:
        continue
    off = int(r[ia], 16) - base
    n = int(r[isamp] or 0)
    tot += n
    key = addr2line.get(off, (None, ""))[0]
    by_line[key] += n
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            by_line_stall[key][hdr[i]] += v
src_cache = {}
def src(key):
    if not key:
        return "?"
    f, l = key
    if f not in src_cache:
        try:
            path = next(p for p in ["p4-fr-sorry-math-but-love-you_b200/csrc/" + f, f] if __import__("os").path.exists(p))
            src_cache[f] = open(path).read().splitlines()
        except StopIteration:
            src_cache[f] = []
    L = src_cache[f]
    return L[l - 1].strip()[:90] if 0 < l <= len(L) else ""
print("total samples", tot)
for key, n in sorted(by_line.items(), key=lambda kv: -kv[1])[:top]:
    st = sorted(by_line_stall[key].items(), key=lambda kv: -kv[1])[:2]
    print("%5.1f%%  %-28s %-40s | %s" % (100.0 * n / tot, "%s:%s" % key if key else "?", ",".join("%s=%d" % (k[6:], v) for k, v in st), src(key)))
