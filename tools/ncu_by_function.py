import csv, sys, subprocess, collections, re
rep=sys.argv[1]; src_path=sys.argv[2]
# map line -> enclosing function by scanning source for "__device__|__global__" definitions
lines=open(src_path).read().split('\n')
func_at=[None]*(len(lines)+2); cur=None
for i,l in enumerate(lines, start=1):
    m=re.match(r'^(?:template.*\n)?\s*(?:__device__|__global__).*?(\w+)\(', l)
    if l.startswith('__device__') or l.startswith('__global__'):
        m=re.search(r'(\w+)\(', l)
        if m: cur=m.group(1)
    func_at[i]=cur
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
for k,r in enumerate(rows[:6]):
    if '# Samples' in r: h=r; start=k+1; break
iS=h.index('# Samples'); iI=h.index('Instructions Executed')
agg=collections.Counter(); aggs=collections.Counter(); cur=None; seen=set()
for r in rows[start:]:
    if len(r)<len(h): continue
    if r[0].strip():
        try: cur=int(r[0])
        except: cur=None
        continue
    try: s=int(r[iS] or 0); ie=int(r[iI] or 0)
    except: continue
    f=func_at[cur] if cur and cur < len(func_at) else 'other'
    agg[f]+=ie; aggs[f]+=s
ti=sum(agg.values()); ts=sum(aggs.values())
print("total warp inst", ti, "samples", ts)
for f,v in agg.most_common(): print(f"{str(f):28s} inst {100*v/ti:5.1f}%  samples {100*aggs[f]/ts:5.1f}%")
