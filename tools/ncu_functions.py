"""Per-function / per-line breakdown of an ncu SASS source CSV using nvdisasm -g line info.
usage: ncu_functions.py <source.csv> <nvdisasm output> <kernel substring> [header file]"""
import csv, re, collections, sys
srccsv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
hdr = sys.argv[4] if len(sys.argv) > 4 else 'gym_rem2d_b200/csrc/rem2d_device.cuh'
lines=open(dis,errors='replace').read().split('\n')
start=next(i for i,l in enumerate(lines) if '.text.' in l and kname in l and l.startswith('//-----'))
cur=None; off2line={}
for l in lines[start+1:]:
    if l.startswith('//-----') and '.text.' in l: break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m:
        if 'inlined at' not in l: cur=(m.group(1).split('/')[-1],int(m.group(2)))
        continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/',l)
    if m and cur: off2line[int(m.group(1),16)]=cur
rows=list(csv.reader(open(srccsv)))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='Address'); h=rows[hi]
col={n:i for i,n in enumerate(h)}
A,TE=col['Address'],col['Thread Instructions Executed']
stalls=[n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
base=None
agg=collections.defaultdict(lambda: collections.defaultdict(float)); tot=collections.defaultdict(float)
for r in rows[hi+1:]:
    if len(r)<=TE or not r[A] or r[A]=='Address': continue
    a=int(r[A],16) if not r[A].isdigit() else int(r[A])
    if base is None: base=a
    key=off2line.get(a-base,('?',0))
    for n in ['# Samples','Instructions Executed','Thread Instructions Executed']+stalls:
        v=float(r[col[n]] or 0); agg[key][n]+=v; tot[n]+=v
print('total samples %d, warp inst %.4g, avg active threads %.2f' % (tot['# Samples'],tot['Instructions Executed'],tot['Thread Instructions Executed']/tot['Instructions Executed']))
print('stall mix (% of samples):', {n[6:]:round(100*tot[n]/tot['# Samples'],1) for n in stalls if tot[n]/tot['# Samples']>0.01})
src=open(hdr).read().split('\n')
def fn_of(line):
    for i in range(min(line,len(src)),0,-1):
        m=re.match(r'\s*(template\s*<[^>]*>\s*)?__device__[\w\s\*&]*?\b(\w+)\s*\(',src[i-1])
        if m: return m.group(2)
    return '?'
fagg=collections.defaultdict(lambda:[0.0,0.0,0.0,0.0])
for (f,ln),v in agg.items():
    name = fn_of(ln) if f==hdr.split('/')[-1] else f
    fagg[name][0]+=v['# Samples']; fagg[name][1]+=v['Instructions Executed']; fagg[name][2]+=v['stall_long_sb']; fagg[name][3]+=v['Thread Instructions Executed']
print('--- by function:            samples  inst   long_sb(of fn samples)  active thr')
for name,v in sorted(fagg.items(), key=lambda kv:-kv[1][0])[:30]:
    print('%-28s %5.1f%% %5.1f%%   %5.1f%%   %4.1f' % (name,100*v[0]/tot['# Samples'],100*v[1]/tot['Instructions Executed'],100*v[2]/max(1,v[0]), v[3]/max(1,v[1])))
print('--- top lines by long_sb stall')
for (f,ln),v in sorted(agg.items(), key=lambda kv:-kv[1]['stall_long_sb'])[:12]:
    t=src[ln-1].strip()[:100] if f==hdr.split('/')[-1] and 0<ln<=len(src) else ''
    print('%-16s %5d  %5.1f%% of all samples | %s' % (f,ln,100*v['stall_long_sb']/tot['# Samples'],t))
