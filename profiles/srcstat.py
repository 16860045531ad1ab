"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the hottest SASS lines."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]
si=hdr.index('Source'); ns=hdr.index('# Samples'); ie=hdr.index('Instructions Executed')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data=[r for r in rows[hi+1:] if len(r)==len(hdr)]
def I(x):
    try: return int(x)
    except: return 0
tot=sum(I(r[ns]) for r in data)
print('kernel', rows[0][1][:100]); print('total samples',tot,'warp-instr',sum(I(r[ie]) for r in data))
agg={hdr[i]:0 for i in stall_cols}
for r in data:
    for i in stall_cols: agg[hdr[i]]+=I(r[i])
print([(k,v) for k,v in sorted(agg.items(), key=lambda x:-x[1])[:8]])
for r in sorted(data,key=lambda r:-I(r[ns]))[:topn]:
    st=sorted(((hdr[i],I(r[i])) for i in stall_cols if I(r[i])>0), key=lambda x:-x[1])[:3]
    print(r[ns], r[ie], r[si][:80], st)
