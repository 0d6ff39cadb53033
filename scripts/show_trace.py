import sys
rows=[l.split() for l in open('gpurun_out/mega_trace.log') if l.startswith('[trace]')]
idx=max(i for i,r in enumerate(rows) if r[1]=='1')
rows=rows[max(0,idx-12):]
rows=sorted(rows,key=lambda r:int(r[2]))
prev=None
for r in rows:
    t=int(r[2]); tag=int(r[1])
    print(f"tag {tag:4d} t={t/1960:7.2f} dt={(t-prev)/1960 if prev is not None else 0:6.2f}")
    prev=t
