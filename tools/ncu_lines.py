"""Aggregate an ncu SASS-level source page by CUDA source line, using nvdisasm -g line info.
usage: ncu_lines.py <sass_csv from `ncu --page source --csv`> <nvdisasm -g -c output> <mangled kernel substring>"""
import csv, re, sys, collections
sass_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1. offset -> (file, line) from nvdisasm
lines = open(dis, errors="replace").read().split("\n")
start = next(i for i, l in enumerate(lines) if ".text." in l and kname in l and l.startswith("//-----"))
cur = None; off2line = {}
for l in lines[start + 1:]:
    if l.startswith("//-----") and ".text." in l: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if "inlined at" not in l: cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m and cur: off2line[int(m.group(1), 16)] = cur
# 2. sass csv
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
A, S, NS, IE, TE = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
base = None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
tot = [0.0, 0.0, 0.0]
opc = collections.defaultdict(lambda: [0.0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= TE or not r[A] or r[A] == "Address": continue
    a = int(r[A], 16) if not r[A].isdigit() else int(r[A])
    if base is None: base = a
    key = off2line.get(a - base, ("?", 0))
    v = [float(r[NS] or 0), float(r[IE] or 0), float(r[TE] or 0)]
    for i in range(3): agg[key][i] += v[i]; tot[i] += v[i]
    op = r[S].split()[0] if r[S].split() else "?"
    if op.startswith("@"): op = r[S].split()[1]
    opc[op.split(".")[0]][0] += v[0]; opc[op.split(".")[0]][1] += v[1]
print("total samples %d, warp inst %.3g, avg active threads %.2f" % (tot[0], tot[1], tot[2] / max(1, tot[1])))
src = {}
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    if f not in src:
        try: src[f] = open("/root/repo/gym_rem2d_b200/csrc/" + f).read().split("\n")
        except Exception: src[f] = []
    text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
    print("%-18s %5d  samp %5.1f%%  inst %5.1f%%  thr %4.1f | %s" % (f, ln, 100 * v[0] / tot[0], 100 * v[1] / tot[1], v[2] / max(1, v[1]), text))
print("--- opcodes by samples")
for op, v in sorted(opc.items(), key=lambda kv: -kv[1][0])[:18]:
    print("%-10s samp %5.1f%%  inst %5.1f%%" % (op, 100 * v[0] / tot[0], 100 * v[1] / tot[1]))
