import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=None; tot={}
for r in rows:
    if r and r[0]=='Address': hdr=r; continue
    if hdr and len(r)==len(hdr):
        for k,v in zip(hdr,r):
            if k.startswith('stall_') and not k.endswith('(Not Issued)'):
                try: tot[k]=tot.get(k,0)+int(v)
                except: pass
s=sum(tot.values())
for k,v in sorted(tot.items(), key=lambda kv:-kv[1])[:12]: print(f'{k:28s} {100*v/s:5.1f}%')
