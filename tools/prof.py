"""Helpers to read ncu reports brought back in gpurun_out/ (used while building profiles/ summaries)."""
import csv, re, subprocess, sys
from collections import defaultdict

def launches(path):
    rows=[r for r in csv.reader(open(path)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    agg=defaultdict(list)
    for r in rows[1:]:
        name=re.sub(r'\(.*','',r[ki]).replace('void ','').replace('<unnamed>::','')[:48]
        try: agg[name].append(float(r[vi].replace(',','')))
        except: pass
    tot=sum(sum(v) for v in agg.values())
    out=[]
    for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
        out.append((k,len(v),sum(v)/1e6,100*sum(v)/tot,[round(x/1e6,3) for x in v[:8]]))
    return out

def raw(rep, names):
    txt=subprocess.run(f"ncu -i {rep} --page raw --csv",shell=True,capture_output=True,text=True).stdout
    rr=list(csv.reader(txt.splitlines())); h=rr[0]; u=rr[1]; v=rr[2]
    return {n:(v[h.index(n)],u[h.index(n)]) for n in names if n in h}

def sass_blocks(rep, per, blk=80):
    txt=subprocess.run(f"ncu -i {rep} --page source --csv",shell=True,capture_output=True,text=True).stdout
    rows=list(csv.reader(txt.splitlines())); hdr=rows[1]; data=rows[2:]
    ia=hdr.index('Instructions Executed'); isrc=hdr.index('Source'); isamp=hdr.index('# Samples')
    tot=sum(int(r[ia]) for r in data if r[ia].isdigit())
    print("total warp instr", tot, "per unit", tot/per)
    for b in range(0,len(data),blk):
        chunk=data[b:b+blk]
        n=sum(int(r[ia]) for r in chunk if r[ia].isdigit())
        if n==0: continue
        s=sum(int(r[isamp]) for r in chunk if r[isamp].isdigit())
        ops={}
        for r in chunk:
            m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc])
            if m and r[ia].isdigit():
                op=m.group(2).split('.')[0]; ops[op]=ops.get(op,0)+int(r[ia])
        top=sorted(ops.items(), key=lambda kv:-kv[1])[:7]
        print(f"{b:5d} instr={n/per:9.1f} ({100*n/tot:5.1f}%) samples={s:6d}  {[(k,round(v/per)) for k,v in top]}")

if __name__=="__main__":
    if sys.argv[1]=="launches":
        for r in launches(sys.argv[2]): print(f"{r[0]:50s} n={r[1]:3d} total={r[2]:9.3f} ms share={r[3]:5.1f}% each={r[4]}")
    elif sys.argv[1]=="sass":
        sass_blocks(sys.argv[2], float(sys.argv[3]))
    elif sys.argv[1]=="raw":
        for k,v in raw(sys.argv[2], sys.argv[3:]).items(): print(k, v)
