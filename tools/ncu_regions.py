"""Stall samples of one kernel per code region: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_regions.py src.csv
(regions = runs of SASS instructions with the same execution count; lists the instructions that collect long-scoreboard samples)."""
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
tot=0; data=[]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    a=int(r[ix['Address']],16) if r[ix['Address']].startswith('0x') else int(r[ix['Address']])
    s=int(r[ix['# Samples']] or 0); n=int(r[ix['Instructions Executed']] or 0)
    data.append((a,s,n,r[ix['Source']], int(r[ix['stall_long_sb']] or 0), int(r[ix['stall_wait']] or 0), int(r[ix['stall_short_sb']] or 0)))
    tot+=s
base=data[0][0]
seg=[]; cur=None
for a,s,n,src,l,w,sh in data:
    if cur is None or abs(n-cur['n'])>0.02*max(n,cur['n'],1):
        cur={'a0':a-base,'n':n,'s':0,'cnt':0,'l':0,'w':0,'sh':0}; seg.append(cur)
    cur['s']+=s; cur['cnt']+=1; cur['a1']=a-base; cur['l']+=l; cur['w']+=w; cur['sh']+=sh
print('total samples',tot)
for g in seg:
    if g['s']>tot*0.006:
        print(f"{g['a0']:06x}-{g['a1']:06x} instr={g['cnt']:4d} exec={g['n']:>9d} samples={g['s']:>7d} {100*g['s']/tot:5.1f}%  long_sb={g['l']} wait={g['w']} short={g['sh']}")
print('--- instructions with long_sb > 0.25% of samples')
for a,s,n,src,l,w,sh in data:
    if l>tot*0.0025: print(f"{a-base:06x} samples={s:6d} long_sb={l:6d} exec={n:9d} {src[:70]}")
