"""Executed warp instructions and stall samples of one launch in an .ncu-rep, summed per source FUNCTION (the line ->
function map is taken from the current sources by matching the line text the report embeds).

    python profiles/hotspots_by_function.py REPORT.ncu-rep LAUNCH_INDEX WARP_FRAMES

WARP_FRAMES = warps x frames of the launch (the divisor of the per-warp-frame column). Needs -lineinfo, --import-source on."""
import csv,io,subprocess,sys,re,collections,bisect
rep,k=sys.argv[1],int(sys.argv[2]); nwf=float(sys.argv[3])
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--launch-skip",str(k),"--launch-count","1"],capture_output=True,text=True).stdout
cur,hdr,rows=None,None,[]
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0]=="File Path": cur,hdr=r[1].split("/")[-1],None
    elif r[0]=="Line No": hdr=r
    elif hdr and r[0]!="" and len(r)>=len(hdr)-2:
        d=dict(zip(hdr,r))
        try: rows.append((cur,int(r[0]),r[1].strip(),float(d["Instructions Executed"]),float(d["# Samples"])))
        except ValueError: pass
base='/root/repo/pikazoo_b200/csrc/'
def funcs(path):
    o=[]
    for i,l in enumerate(open(path),1):
        m=re.match(r'\s*(?:static )?(?:__host__ )?(?:__device__|__global__).*?(\w+)\(',l)
        if m: o.append((i,m.group(1)))
    return o
cache={}
byfn=collections.Counter(); smp=collections.Counter()
for f,ln,text,e,s in rows:
    if e==0 and s==0: continue
    try:
        if f not in cache:
            L=open(base+f).read().split('\n'); cache[f]=(L,funcs(base+f))
        L,fs=cache[f]
        # find the current line number by text near ln
        cand=[i+1 for i,t in enumerate(L) if t.strip()==text and text]
        cl=min(cand,key=lambda c:abs(c-ln)) if cand else ln
        i=bisect.bisect_right([x[0] for x in fs],cl)-1
        fn=fs[i][1] if i>=0 else '?'
    except FileNotFoundError:
        fn=f
    byfn[fn]+=e; smp[fn]+=s
tot=sum(byfn.values())
print('total',tot,'per warp-frame',tot/nwf)
for fn,v in byfn.most_common(40): print(f'{fn:32s} {v/nwf:7.1f} {100*v/tot:5.1f}%  smp {100*smp[fn]/sum(smp.values()):5.1f}%')
