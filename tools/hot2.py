# aggregate ncu source-page instruction counts by function (line ranges) for k_pileup_fast
import re,csv,collections,sys,os
rep=sys.argv[1]; kern=sys.argv[2] if len(sys.argv)>2 else '_Z13k_pileup_fast9PlpParams'
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.system("cd /tmp && rm -rf cub && mkdir cub && cd cub && cuobjdump -xelf all "+ROOT+"/raer_b200/libraer_b200.so >/dev/null 2>&1 && "
          "for c in *.cubin; do nvdisasm -g $c > $c.sass 2>/dev/null; done; for c in *.sass; do if grep -q '^.text."+kern+":' $c; then cp $c /tmp/all.sass; fi; done")
os.system(f"ncu -i {rep} --page source --csv 2>/dev/null > /tmp/sass_m.csv")
lines=open('/tmp/all.sass').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.'+kern+':')][0]
cur=None; off2line={}
for l in lines[start+1:]:
    if (l.startswith('.text.') or l.startswith('//-----')) and off2line: break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/',l)
    if m: off2line[int(m.group(1),16)]=cur
# function ranges from sources
def ranges(path):
    out=[]; src=open(path).read().split('\n')
    for i,l in enumerate(src):
        m=re.match(r'^(?:template.*\n)?(?:static )?(?:__device__|__global__|extern "C").*?(\w+)\(',l)
        if m and not l.startswith(' '): out.append((i+1,m.group(1)))
    return out
R={f:ranges(ROOT+'/raer_b200/csrc/'+f) for f in ('raer_fast.cuh','raer_kernels.cu','raer_walk.cuh','raer_lane.cuh','raer_ginflate.cuh','raer_gingest.cu','raer_sc.cu','raer_radix.cuh')}
def fn(ln):
    if not ln: return '?'
    f,n=ln
    if f not in R: return f
    name='?'
    for s,nm in R[f]:
        if s<=n: name=nm
        else: break
    return f.split('.')[0][5:]+':'+name
rows=list(csv.reader(open('/tmp/sass_m.csv')))
hdr=rows[1]; ix={n:i for i,n in enumerate(hdr)}; data=rows[2:]
base=int(data[0][0],16)
agg=collections.defaultdict(lambda:[0,0,0]); ti=ts=0
for r in data:
    ln=off2line.get(int(r[0],16)-base)
    inst=float(r[ix['Instructions Executed']] or 0); samp=float(r[ix['# Samples']] or 0); th=float(r[ix['Thread Instructions Executed']] or 0)
    a=agg[fn(ln)]; a[0]+=inst; a[1]+=samp; a[2]+=th; ti+=inst; ts+=samp
print("total warp inst %.3e samples %d"%(ti,ts))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print("%-40s inst %5.1f%%  (%.2e)  thr/inst %4.1f  samples %5.1f%%"%(k,100*a[0]/ti,a[0],a[2]/max(a[0],1),100*a[1]/ts))

# hottest source lines
lagg=collections.defaultdict(lambda:[0,0,0])
for r in data:
    ln=off2line.get(int(r[0],16)-base)
    inst=float(r[ix['Instructions Executed']] or 0); samp=float(r[ix['# Samples']] or 0); th=float(r[ix['Thread Instructions Executed']] or 0)
    a=lagg[ln]; a[0]+=inst; a[1]+=samp; a[2]+=th
if len(sys.argv)>3:
    print("\nhottest source lines:")
    for k,a in sorted(lagg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[3])]:
        print("%-28s inst %5.1f%% thr/inst %4.1f samples %5.1f%%"%(str(k),100*a[0]/ti,a[2]/max(a[0],1),100*a[1]/ts))
if len(sys.argv)>5:
    f=sys.argv[4]; lo,hi=[int(x) for x in sys.argv[5].split('-')]
    src=open(ROOT+'/raer_b200/csrc/'+f).read().split('\n')
    print("\nper line %s %d-%d (warp inst in 1e6, thr/inst, samples%%):"%(f,lo,hi))
    for n in range(lo,hi+1):
        a=lagg.get((f,n))
        if a: print("%4d %8.1f %5.1f %5.1f%% | %s"%(n,a[0]/1e6,a[2]/max(a[0],1),100*a[1]/ts,src[n-1][:110]))
        else: print("%4d %8s %5s %6s | %s"%(n,'','','',src[n-1][:110]))
