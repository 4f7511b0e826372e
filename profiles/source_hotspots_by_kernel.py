import csv,subprocess,sys
rep,kern=sys.argv[1],sys.argv[2]
top=int(sys.argv[3]) if len(sys.argv)>3 else 25
txt=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-name","regex:"+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
per={}; ii=None; fname=None; stall={}
for r in rows:
    if not r: continue
    if r[0]=="File Path": fname=r[1].split('/')[-1]
    elif r[0]=="Line No":
        hdr=r
        if "Instructions Executed" in hdr:
            ii,isamp=hdr.index("Instructions Executed"),hdr.index("# Samples")
            sidx={h:i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    elif r[0].isdigit() and ii is not None and len(r)>ii:
        try: n,s=int(r[ii]),int(r[isamp])
        except ValueError: continue
        key=(fname,int(r[0])); d=per.setdefault(key,[0,0,r[1].strip()[:100],{}]); d[0]+=n; d[1]+=s
        for h,i in sidx.items():
            try: v=int(r[i])
            except ValueError: v=0
            if v: d[3][h]=d[3].get(h,0)+v
tot=sum(v[0] for v in per.values()) or 1; tots=sum(v[1] for v in per.values()) or 1
print(kern,"inst",tot,"samples",tots)
for k,(n,s,src,st) in sorted(per.items(), key=lambda kv:-kv[1][1])[:top]:
    top2=sorted(st.items(),key=lambda kv:-kv[1])[:2]
    print("%5.1f%% inst %5.1f%% samp %s:%d %s   [%s]"%(100*n/tot,100*s/tots,k[0],k[1],src," ".join("%s=%d"%(a.replace('stall_',''),b) for a,b in top2)))
