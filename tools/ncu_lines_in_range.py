"""Per-line instruction and sample counts for a line range of one source file in an ncu report."""
import csv, sys, subprocess
rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; hdr=None; agg={}; grid=None
for r in rows:
    if r and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': hdr=r; continue
    if hdr and r and r[0].isdigit() and cur==fname:
        try: inst=int(r[hdr.index('Instructions Executed')]); samp=int(r[hdr.index('# Samples')])
        except: continue
        ln=int(r[0])
        if lo<=ln<=hi:
            a=agg.setdefault(ln,[0,0,r[1].strip()[:110]]); a[0]+=inst; a[1]+=samp
div=float(sys.argv[5]) if len(sys.argv)>5 else 1.0
for ln in sorted(agg): print(f'{ln:5d} inst/tile {agg[ln][0]/div:8.1f} samp {agg[ln][1]:5d}  {agg[ln][2]}')
print('sum inst/tile', sum(a[0] for a in agg.values())/div)
