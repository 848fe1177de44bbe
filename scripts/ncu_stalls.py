"""Per-function / per-line breakdown of ONE stall reason from an ncu report (default: long scoreboard).
usage: B200POA_SO=<matching .so> python scripts/ncu_stalls.py <report.ncu-rep> [stall_long_sb] [--top 30]"""
import csv, io, os, re, subprocess, sys, tempfile, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
reason = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "stall_long_sb"
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
so = os.path.abspath(os.environ.get("B200POA_SO", os.path.join(ROOT, "racon_gpu_b200", "libb200poa.so")))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("b200poa.") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
addr2line, cur, fn_of_addr, curfn = {}, None, {}, None
for line in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line) or re.match(r"^(\S+):\s*$", line)
    m2 = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m2:
        cur = (os.path.basename(m2.group(1)), int(m2.group(2)))
        continue
    m3 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", line)
    if m3 and cur:
        addr2line[int(m3.group(1), 16)] = (cur, m3.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {n: i for i, n in enumerate(hdr)}
by_line, by_sass = collections.Counter(), []
tot_all = tot = 0
base = None
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        addr = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
    except ValueError:
        continue
    if base is None:
        base = addr
    v = int(float(r[ci[reason]] or 0))
    tot_all += int(float(r[ci["# Samples"]] or 0))
    tot += v
    key = addr2line.get(addr - base, (("?", 0), ""))
    by_line[key[0]] += v
    by_sass.append((v, key[0], r[ci["Source"]]))
print(f"{reason}: {tot} of {tot_all} samples ({100*tot/max(tot_all,1):.1f}%)")
src = {}
for (fn, ln), v in by_line.most_common(top):
    path = os.path.join(ROOT, "racon_gpu_b200", "csrc", fn)
    if fn not in src and os.path.exists(path):
        src[fn] = open(path).read().splitlines()
    text = src[fn][ln - 1].strip() if fn in src and 0 < ln <= len(src[fn]) else ""
    print(f"{100*v/max(tot_all,1):6.2f}%  {fn}:{ln:<5d} {text[:110]}")
if "--sass" in sys.argv:
    for v, (fn, ln), sass in sorted(by_sass, reverse=True)[:top]:
        print(f"{100*v/max(tot_all,1):6.2f}%  {fn}:{ln:<5d} {sass[:100]}")
