import csv,sys,re,subprocess
rep=sys.argv[1]; pat=sys.argv[2]
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","-k","regex:"+pat],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=None; data=[]; kname=None; done=False
def flush():
    if not data: return
    tot=sum(d[0] for d in data); print('==',kname[:70],'samples',tot,'instrs',len(data))
    start=0; acc=0
    for i,(n,s) in enumerate(data):
        acc+=n
        if re.search(r'BAR.SYNC|EXIT', s) or i==len(data)-1:
            ops={}
            for n2,s2 in data[start:i+1]:
                op=s2.split()[0] if not s2.startswith('@') else s2.split()[1]
                op=op.split('.')[0]; ops[op]=ops.get(op,0)+n2
            top=sorted(ops.items(),key=lambda x:-x[1])[:7]
            if acc*100/tot>0.5: print('instr %4d-%4d  %5.1f%%  %s'%(start,i,100*acc/tot,' '.join('%s:%d'%(k,v) for k,v in top)))
            start=i+1; acc=0
    # hottest single instructions
    for n,i,s in sorted([(n,i,s) for i,(n,s) in enumerate(data)],reverse=True)[:12]: print('   hot #%4d %5.1f%% %s'%(i,100*n/tot,s[:80]))
for r in rows:
    if r and r[0]=='Kernel Name':
        if data: break
        kname=r[1]; hdr=None; continue
    if r and r[0]=='Address': hdr=r; si=hdr.index('# Samples'); src=hdr.index('Source'); continue
    if hdr and len(r)>si and r[si].isdigit(): data.append((int(r[si]), r[src].strip()))
flush()
