"""Print SASS instructions (with stall samples) that map to a range of source lines."""
import csv, re, sys
sass_csv, dis, kname, f, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
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
rows = list(csv.reader(open(sass_csv)))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi_]
A, S, NS, IE, TE = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
base = None; tot = 0
out = []
for r in rows[hi_ + 1:]:
    if len(r) <= TE or not r[A] or r[A] == "Address": continue
    a = int(r[A], 16) if not r[A].isdigit() else int(r[A])
    if base is None: base = a
    tot += float(r[NS] or 0)
    key = off2line.get(a - base, ("?", 0))
    out.append((a - base, key, r[S], float(r[NS] or 0), float(r[IE] or 0), float(r[TE] or 0)))
first = next(i for i, o in enumerate(out) if o[1][0] == f and lo <= o[1][1] <= hi)
last = max(i for i, o in enumerate(out) if o[1][0] == f and lo <= o[1][1] <= hi)
for o in out[first:last + 1]:
    print("%06x L%-4d samp %5.2f%% exec %9.3g thr %4.1f | %s" % (o[0], o[1][1], 100 * o[3] / tot, o[4], o[5] / max(1, o[4]), o[2][:90]))
