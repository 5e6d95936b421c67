import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
secs=[i for i,r in enumerate(rows) if r and r[0]=="File Path"]+[len(rows)]
allfiles={}
for a,b in zip(secs[:-1],secs[1:]):
    path=rows[a][1]; hdr=rows[a+2]
    iexec=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples")
    per=collections.defaultdict(lambda:[0,0])
    for r in rows[a+3:b]:
        try: n=int(r[iexec]); s=int(r[isamp])
        except: continue
        if r[0].isdigit(): per[int(r[0])][0]+=n; per[int(r[0])][1]+=s
    if path not in allfiles: allfiles[path]=per
tot=sum(v[0] for per in allfiles.values() for v in per.values()); tots=sum(v[1] for per in allfiles.values() for v in per.values())
print("total inst",tot,"samples",tots)
for path,per in allfiles.items():
    try: src=open(path).read().split('\n')
    except: src=None
    ft=sum(v[0] for v in per.values()); fs=sum(v[1] for v in per.values())
    print(f"== {path}: {ft/tot*100:.1f}% inst {fs/max(tots,1)*100:.1f}% samples")
    if 'maexp' not in path: continue
    for ln,(n,s) in sorted(per.items(), key=lambda kv:-kv[1][1])[:topn]:
        t=src[ln-1].strip()[:95] if src and ln<=len(src) else ""
        print(f"  {ln:>4} {n/tot*100:5.1f}%i {s/tots*100:5.1f}%s  {t}")
