"""Per-opcode and per-source-line instruction / stall-sample digest of one kernel from an ncu report with source
(ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv).  SASS rows are de-duplicated by address (an inlined
instruction is listed under every file of its inline stack); the per-line table uses the rows of one file (call-site level) and
may count an instruction under more than one line of that file.
usage: python tools/ncu_source_digest.py src.csv <kernel name substring> <warps x steps to normalise by> <file for the per-line table>"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
kern=sys.argv[2]; per=float(sys.argv[3]); topfile=sys.argv[4]
secs=[i for i,r in enumerate(rows) if r and r[0]=="File Path"]; secs.append(len(rows))
num=lambda s:int(s) if s not in ("","-") else 0
seen={}; agg=collections.defaultdict(lambda:[0,0,0,collections.Counter()])
for si in range(len(secs)-1):
    a,b=secs[si],secs[si+1]
    if kern not in rows[a+1][1]: continue
    fn=rows[a][1].split('/')[-1]; hdr=rows[a+2]
    iS=[j for j,h in enumerate(hdr) if h=="Source"]
    iI=hdr.index("Instructions Executed"); iSm=hdr.index("# Samples"); iW=hdr.index("L1 Wavefronts Shared"); iA=hdr.index("Address")
    cur=None
    for r in rows[a+3:b]:
        if len(r)<len(hdr): continue
        if r[0]!="": cur=(int(r[0]), r[1].strip()[:90]); continue
        if r[iA] in ("","..."): continue
        t=r[iS[1]].split(); op=t[1] if t[0].startswith('@') else t[0]
        seen[r[iA]]=(num(r[iI]),num(r[iSm]),num(r[iW]),op)
        if fn==topfile:
            g=agg[cur]; g[0]+=num(r[iI]); g[1]+=num(r[iSm]); g[2]+=num(r[iW]); g[3][op]+=num(r[iI])
ti=sum(v[0] for v in seen.values()); ts=sum(v[1] for v in seen.values()); tw=sum(v[2] for v in seen.values())
print("unique: inst/warp-step %.1f  shared wavefronts/warp-step %.1f samples %d"%(ti/per, tw/per, ts))
opc=collections.Counter(); ops=collections.Counter()
for v in seen.values(): opc[v[3]]+=v[0]; ops[v[3]]+=v[1]
for k,v in opc.most_common(30): print(f"  {k:24s} {v/per:7.1f} {100*ops[k]/ts:5.1f}%")
print("by line of",topfile,"(inst/warp-step, samples%, shared wavefronts, top opcodes)")
ai=sum(v[0] for v in agg.values())
print("  covered inst/warp-step %.1f"%(ai/per))
for k,v in sorted(agg.items()):
    if v[0]/per<8: continue
    print(f"  {k[0]:5d} {v[0]/per:7.1f} {100*v[1]/ts:5.1f}% {v[2]/per:6.1f}  {k[1][:70]}  | "+" ".join(f"{o}:{c/per:.0f}" for o,c in v[3].most_common(5)))
