import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=next(i for i,r in enumerate(rows) if r and r[0]=="Address")
names=rows[hdr]; si=names.index("Warp Stall Sampling (All Samples)"); src=names.index("Source"); ie=names.index("Instructions Executed")
data=[]
for r in rows[hdr+1:]:
    if len(r)!=len(names) or not r[si].strip().isdigit():
        if data: break
        continue
    data.append(r)
tot=sum(int(r[si]) for r in data); toti=sum(int(r[ie]) for r in data)
print("instructions", len(data), "samples", tot, "warp-instr executed", toti)
MARK=("BAR.SYNC","SYNCS","LDTM","UTCHMMA","UTMALDG","UTCBAR","STS","MUFU.EX2","ST.E","STG","LDG","LD.E","ATOM","RED","FENCE")
B=int(sys.argv[3]) if len(sys.argv)>3 else 100
for b0 in range(0,len(data),B):
    chunk=data[b0:b0+B]
    s=sum(int(r[si]) for r in chunk); ins=sum(int(r[ie]) for r in chunk)
    marks={}
    for r in chunk:
        for m in MARK:
            if m in r[src]: marks[m]=marks.get(m,0)+1
    stalls={}
    for j,n in enumerate(names):
        if n.startswith("stall_") and "Not Issued" not in n:
            v=sum(int(r[j] or 0) for r in chunk)
            if v: stalls[n[6:]]=v
    top=sorted(stalls.items(), key=lambda kv:-kv[1])[:3]
    print(f"{b0:5d}-{b0+B:5d} samples {100*s/tot:5.1f}%  instr {100*ins/toti:5.1f}%  {dict(top)}  {marks}")
