import csv, sys, subprocess, collections
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
for k,r in enumerate(rows[:6]):
    if '# Samples' in r: h=r; start=k+1; break
iS=h.index('# Samples'); iI=h.index('Instructions Executed'); iL=0; iSrc=1
agg=collections.OrderedDict(); cur=None
for r in rows[start:]:
    if len(r)<len(h): continue
    if r[iL].strip(): cur=(r[iL], r[iSrc])
    try: s=int(r[iS] or 0); ie=int(r[iI] or 0)
    except: continue
    a=agg.setdefault(cur,[0,0]); a[0]+=ie; a[1]+=s
toti=sum(a[0] for a in agg.values()); tots=sum(a[1] for a in agg.values())
print("tot warp inst", toti, "tot samples", tots)
print("--- top by instructions executed")
for (ln,src),a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:topn]:
    print(f"{a[0]:11d} {100*a[0]/toti:5.1f}%  samp {100*a[1]/tots:5.1f}%  L{ln:>4}: {src.strip()[:105]}")
print("--- top by samples")
for (ln,src),a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:topn]:
    print(f"{a[0]:11d} {100*a[0]/toti:5.1f}%  samp {100*a[1]/tots:5.1f}%  L{ln:>4}: {src.strip()[:105]}")
