"""Per-instruction shared-memory wavefronts of one kernel from an .ncu-rep: python scripts/ncu_smem.py rep kernel_regex [top_n]"""
import csv,io,subprocess,sys
rep,rx=sys.argv[1],sys.argv[2]; topn=int(sys.argv[3]) if len(sys.argv)>3 else 30
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+rx],capture_output=True,text=True).stdout
lines=raw.splitlines()
rows=list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
def num(x):
    try: return int(x)
    except: return 0
tot=0; items=[]; totex=0
for r in rows[1:]:
    if len(r)<len(hdr): continue
    w=num(r[ix["L1 Wavefronts Shared"]]); ex=num(r[ix["Instructions Executed"]]); ideal=num(r[ix["L1 Wavefronts Shared Ideal"]])
    tot+=w; totex+=ex
    if w: items.append((w,ex,ideal,r[ix["Source"]]))
print("total smem wavefronts",tot,"warp-instructions",totex)
for w,ex,ideal,src in sorted(items,reverse=True)[:topn]:
    print(f"{100*w/tot:5.1f}% wf={w:>10} ex={ex:>9} wf/inst={w/max(ex,1):.2f} ideal/inst={ideal/max(ex,1):.2f}  {src[:60]}")
