"""Stall samples and executed instructions of an ncu report summed over source-line ranges of snk_kernels.cu."""
import csv, sys, subprocess
rep=sys.argv[1]
RANGES=[('place_fruits',82,150),('reset_env',153,245),('tma/mbar helpers',291,328),('step_group',339,466),('tile_rules',469,552),
        ('viewer_origin/mask',554,568),('encode_fs1 legacy',575,658),('encode reg/direct',660,767),('encode stacked',769,900),
        ('kernel prologue',908,999),('kernel encode loop/epilogue',1000,1082)]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; hdr=None; agg={}; other=[0,0]
for r in rows:
    if r and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': hdr=r; continue
    if hdr and r and r[0].isdigit():
        try: inst=int(r[hdr.index('Instructions Executed')]); samp=int(r[hdr.index('# Samples')])
        except: continue
        ln=int(r[0]); name=None
        if cur=='snk_kernels.cu':
            for n,a,b in RANGES:
                if a<=ln<=b: name=n
        name=name or ('['+str(cur)+']')
        x=agg.setdefault(name,[0,0]); x[0]+=inst; x[1]+=samp
ti=sum(v[0] for v in agg.values()); ts=sum(v[1] for v in agg.values())
for n,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f'{n:32s} samp {100*v[1]/ts:5.1f}%  inst {100*v[0]/ti:5.1f}%')
