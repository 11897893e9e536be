import csv, collections, sys, subprocess
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv)>3 else 30
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-name","regex:"+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; per=collections.Counter(); perinst=collections.Counter(); reasons=collections.Counter()
fileof=None; hdr=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": fileof=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No":
        hdr=r; isamp=hdr.index("# Samples"); iinst=hdr.index("Instructions Executed")
        st={n:hdr.index(n) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
        continue
    if hdr is None: continue
    if r[0].strip()!="":
        try: cur=(fileof,int(r[0]),r[1].strip()[:100])
        except: pass
    if len(r)>2 and r[2].startswith("0x"):
        try:
            n=int(r[isamp] or 0); ni=int(r[iinst] or 0)
        except Exception: continue
        per[cur]+=n; perinst[cur]+=ni
        for k,i in st.items():
            try: reasons[k]+=int(r[i] or 0)
            except: pass
tot=sum(per.values()); toti=sum(perinst.values())
print("total samples",tot,"warp insts",toti)
print(", ".join("%s %.1f%%"%(k[6:],100*v/tot) for k,v in reasons.most_common(9)))
for k,v in sorted(per.items(), key=lambda x:-x[1])[:top]:
    print("%5.1f%% inst %5.1f%%  %s:%d  %s"%(100*v/tot,100*perinst[k]/toti,k[0],k[1],k[2]))
