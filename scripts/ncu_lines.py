"""Attribute ncu warp-stall samples to CUDA source lines / functions.

usage: python scripts/ncu_lines.py <report.ncu-rep> [--top 25]
Joins `ncu --page source --csv` (SASS addresses + samples) with `nvdisasm -g` line info of the
kernel's cubin extracted from racon_gpu_b200/libb200poa.so (built with -lineinfo).
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
so = os.environ.get("B200POA_SO", os.path.join(ROOT, "racon_gpu_b200", "libb200poa.so"))

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("b200poa.") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
addr2line, cur = {}, None
for line in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", line)
    if m and cur:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())

# function ranges from the sources
funcs = {}
for fn in ("poa_core.cuh", "poa_fill.cuh", "b200poa.cu", "poa_simt.cuh"):
    path = os.path.join(ROOT, "racon_gpu_b200", "csrc", fn)
    marks = []
    for i, l in enumerate(open(path), 1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:POA_FN|__device__|__global__|static|POA_FN_NOINLINE).*?\b([A-Za-z_0-9]+)\s*\(", l)
        if m and not l.strip().startswith(("//", "*", "/*")) and m.group(1) not in ("if", "for", "while", "defined", "__launch_bounds__"):
            marks.append((i, m.group(1)))
        m2 = re.match(r"\s*__device__\s+(?:__forceinline__\s+)?\S+\s+(operator\(\)|[A-Za-z_0-9]+)\s*\(", l)
        if m2 and not marks[-1:] == [(i, m2.group(1))]:
            marks.append((i, m2.group(1)))
    funcs[fn] = marks

def func_of(fn, line):
    name = "?"
    for (i, n) in funcs.get(fn, []):
        if i <= line:
            name = n
    return name

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {h: i for i, h in enumerate(hdr)}
by_line, by_func, by_func_inst = collections.Counter(), collections.Counter(), collections.Counter()
total = 0
total_inst = 0
base = None
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        addr = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
    except ValueError:
        continue
    if base is None:
        base = addr
    s = int(float(r[ci["# Samples"]] or 0))
    inst = int(float(r[ci["Instructions Executed"]] or 0))
    key = addr2line.get(addr - base)
    if key is None:
        key = (("?", 0), r[ci["Source"]])
    (fn, ln), sass = key
    by_line[(fn, ln)] += s
    by_func[(fn, func_of(fn, ln))] += s
    by_func_inst[(fn, func_of(fn, ln))] += inst
    total += s
    total_inst += inst
print(f"total samples {total}, warp instructions {total_inst}")
print("--- by function (samples %, instructions %)")
for (fn, f), s in by_func.most_common(20):
    print(f"{100*s/total:6.2f}%  inst {100*by_func_inst[(fn,f)]/max(total_inst,1):6.2f}%  {fn}:{f}")
print("--- by line")
src_cache = {}
for (fn, ln), s in by_line.most_common(top):
    path = os.path.join(ROOT, "racon_gpu_b200", "csrc", fn)
    if fn not in src_cache and os.path.exists(path):
        src_cache[fn] = open(path).read().splitlines()
    text = src_cache.get(fn, [""] * (ln + 1))[ln - 1].strip() if ln and fn in src_cache else ""
    print(f"{100*s/total:6.2f}%  {fn}:{ln:<5d} {text[:100]}")

if "--inst" in sys.argv:
    by_line_inst = collections.Counter()
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            addr = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
        except ValueError:
            continue
        inst = int(float(r[ci["Instructions Executed"]] or 0))
        key = addr2line.get(addr - base)
        if key is None:
            continue
        (fn, ln), sass = key
        by_line_inst[(fn, ln)] += inst
    scale = float(sys.argv[sys.argv.index("--inst") + 1])  # divide counts by this (e.g. total rows)
    print("--- instructions per unit by line (top 60)")
    for (fn, ln), c in by_line_inst.most_common(60):
        text = src_cache.get(fn, [""] * (ln + 1))[ln - 1].strip() if ln and fn in src_cache else ""
        print(f"{c/scale:8.2f}  {fn}:{ln:<5d} {text[:90]}")
