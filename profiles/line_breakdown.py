import csv,re,sys
sys.path.insert(0,'/root/repo/profiles')
from line_hotspots import sass_lines
from collections import defaultdict
src_csv, kern, div = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows=list(csv.reader(open(src_csv)))
hdr=rows[1]; ii=hdr.index("Instructions Executed"); isrc=hdr.index("Source"); isamp=hdr.index("# Samples")
inst=[(int(r[ii]), r[isrc], int(r[isamp])) for r in rows[2:] if len(r)>ii]
lines=sass_lines('/tmp/f32_disasm.txt',kern)
agg=defaultdict(int); ops=defaultdict(lambda: defaultdict(int)); samp=defaultdict(int)
for (n,src,sm),ln in zip(inst,lines):
    agg[ln]+=n; samp[ln]+=sm
    t=src.split()
    op=t[1] if t[0].startswith('@') else t[0]
    ops[ln][op.split('.')[0]]+=n
tot=0; ts=sum(samp.values())
for ln in sorted(agg, key=lambda x:(x[0],x[1]) if x else ('',0)):
    v=agg[ln]/div
    tot+=v
    if v>=1.0:
        top=sorted(ops[ln].items(), key=lambda kv:-kv[1])[:6]
        print(ln, f"{v:6.1f} s{100*samp[ln]/ts:4.1f}%", " ".join(f"{k}:{c/div:.1f}" for k,c in top))
print("sum", tot)
