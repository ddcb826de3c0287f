"""tools/ncu_regions.py — developer aid: split a kernel's `ncu --page source --csv --print-source sass` export into the
regions between its barriers and print, per region, the share of executed warp-instructions, of fp64-pipe and XU
instructions, the shared-memory wavefronts, the share of stall samples and the largest stall classes.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv; python tools/ncu_regions.py x.csv KERNEL_INDEX [BREAK_REGEX]"""
import csv,re,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]
which=int(sys.argv[2])
s=starts[which]; e=starts[which+1] if which+1<len(starts) else len(rows)
print(rows[s][1][:80])
hdr=rows[s+1]; data=rows[s+2:e]
ix={h:i for i,h in enumerate(hdr)}
def f(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
tot=sum(f(r,"Instructions Executed") for r in data); tots=sum(f(r,"# Samples") for r in data)
print("total warp-inst",tot,"samples",tots)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def new(i): return dict(start=i,inst=0,samp=0,shw=0,fp64=0,xu=0,st=collections.Counter())
seg=[];cur=new(0)
brk=re.compile(sys.argv[3]) if len(sys.argv)>3 else re.compile(r'\b(BAR|WARPSYNC)\b')
for i,r in enumerate(data):
    s_=r[ix["Source"]]; n=f(r,"Instructions Executed")
    cur['inst']+=n; cur['samp']+=f(r,"# Samples"); cur['shw']+=f(r,"L1 Wavefronts Shared")
    if re.search(r'\b(DADD|DMUL|DFMA|DSETP)\b',s_): cur['fp64']+=n
    if re.search(r'\b(MUFU|I2F|F2F|F2I)',s_): cur['xu']+=n
    for h in stalls: cur['st'][h]+=f(r,h)
    if brk.search(s_) and n>0:
        cur['end']=i; cur['bar']=s_.strip()[:44]; seg.append(cur); cur=new(i+1)
cur['end']=len(data); cur['bar']='END'; seg.append(cur)
for s_ in seg:
    if s_['inst']<1: continue
    top=', '.join(f"{k[6:]}:{v/max(s_['samp'],1)*100:.0f}%" for k,v in s_['st'].most_common(5))
    print(f"[{s_['start']:4d}-{s_['end']:4d}] inst {s_['inst']/tot*100:5.1f}% fp64 {s_['fp64']/tot*100:5.1f}% xu {s_['xu']/tot*100:4.1f}% shw {s_['shw']/1e3:7.0f}k samp {s_['samp']/tots*100:5.1f}%  {top} | {s_['bar']}")
print("fp64 total %.1f%% of inst = %.2fM ; xu %.2fM"%(sum(s_['fp64'] for s_ in seg)/tot*100,sum(s_['fp64'] for s_ in seg)/1e6,sum(s_['xu'] for s_ in seg)/1e6))
