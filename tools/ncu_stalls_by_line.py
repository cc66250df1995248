import csv, subprocess, sys, collections
path=sys.argv[1]
out = subprocess.run(['ncu','-i',path,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
# find header row
hi=[i for i,r in enumerate(rows) if r and r[0]=='Line No'][0]
hdr=rows[hi]; data=rows[hi+1:]
iline=0; isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
stall=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg=collections.defaultdict(lambda:[0,0,collections.Counter(),''])
cur=None
tot=0
for r in data:
    if len(r)<len(hdr): continue
    ln=r[0]
    try: s=int(r[isamp]); e=int(r[iex])
    except: continue
    a=agg[ln]; a[0]+=s; a[1]+=e; tot+=s
    if not a[3]: a[3]=r[1]
    for i,h in stall:
        try: a[2][h[6:]]+=int(r[i])
        except: pass
print('total',tot)
for ln,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:45]:
    print(f"{ln:>5} {a[0]:>6} {100*a[0]/tot:5.1f}% ex={a[1]:>9} {a[3].strip()[:70]:70s} {dict(a[2].most_common(3))}")
