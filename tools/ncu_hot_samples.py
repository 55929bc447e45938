"""Top source lines of an ncu report by stall SAMPLES (where warps wait), with the dominant stall reason."""
import csv, sys, subprocess
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; agg={}; hdr=None
for r in rows:
    if r and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': hdr=r; continue
    if r and r[0]=='Function Name': continue
    if hdr and r and r[0].isdigit():
        ie=hdr.index('Instructions Executed'); ns=hdr.index('# Samples')
        try: inst=int(r[ie]); samp=int(r[ns])
        except: continue
        a=agg.setdefault((cur,int(r[0])),[0,0,r[1].strip()[:90],{}])
        a[0]+=inst; a[1]+=samp
        for k,v in zip(hdr,r):
            if k.startswith('stall_') and not k.endswith('(Not Issued)'):
                try: a[3][k]=a[3].get(k,0)+int(v)
                except: pass
tot=sum(a[0] for a in agg.values()); ts=sum(a[1] for a in agg.values())
print('total inst',tot,'samples',ts)
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:topn]:
    top=sorted(a[3].items(), key=lambda kv:-kv[1])[:2]
    print(f'{k[0]}:{k[1]:4d} samp {100*a[1]/ts:5.1f}% inst {100*a[0]/tot:5.1f}% {",".join(f"{n[6:]}={v}" for n,v in top if v)}  {a[2]}')
