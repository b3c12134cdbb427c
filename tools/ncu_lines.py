import sys,re,subprocess,csv,io,collections
rep,cubin,name,kre=sys.argv[1:5]
top=int(sys.argv[5]) if len(sys.argv)>5 else 30
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--launch-count","1","--kernel-name","regex:"+kre],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=next(r for r in rows if r and r[0]=="Address")
ie,si=hdr.index("Instructions Executed"),hdr.index("Source")
counts=[];seen=set()
for r in rows[rows.index(hdr)+1:]:
    if len(r)>ie and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0]); counts.append((r[si].strip(),int(r[ie])))
dis=subprocess.run(["nvdisasm","-c","-g",cubin],capture_output=True,text=True).stdout.splitlines()
start=next(i for i,l in enumerate(dis) if l.startswith("\t.section\t.text.") and name in l)
lines=[];cur=None
for l in dis[start+1:]:
    if l.startswith("\t.section"): break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split("/")[-1],int(m.group(2))); continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/",l): lines.append(cur)
print("sass: report %d disasm %d"%(len(counts),len(lines)))
by=collections.Counter()
for (s,n),ln in zip(counts,lines): by[ln]+=n
tot=sum(by.values()); srcs={}
for (f,n),c in by.most_common(top):
    if f not in srcs:
        try: srcs[f]=open("/root/repo/exvision_featurecraft_b200/csrc/"+f).read().splitlines()
        except OSError: srcs[f]=[]
    t=srcs[f][n-1].strip()[:110] if 0<n<=len(srcs[f]) else ""
    print("%5.1f%% %s:%d %s"%(100.0*c/tot,f,n,t))
