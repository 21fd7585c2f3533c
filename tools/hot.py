import re,csv,collections,subprocess,sys,os
rep=sys.argv[1]; kern=sys.argv[2] if len(sys.argv)>2 else '_Z14k_pileup_tiles9PlpParams'
os.system("cd /tmp && rm -rf cub && mkdir cub && cd cub && cuobjdump -xelf all /root/repo/raer_b200/libraer_b200.so >/dev/null 2>&1 && nvdisasm -g raer_kernels.sm_100a.cubin > /tmp/all.sass 2>/dev/null")
os.system(f"ncu -i {rep} --page source --csv 2>/dev/null > /tmp/sass_m.csv; ncu -i {rep} --page raw --csv 2>/dev/null > /tmp/raw_m.csv")
lines=open('/tmp/all.sass').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.'+kern+':')][0]
cur=None; off2line={}
for l in lines[start+1:]:
    if (l.startswith('.text.') or l.startswith('//-----')) and off2line: break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/',l)
    if m: off2line[int(m.group(1),16)]=cur
rows=list(csv.reader(open('/tmp/sass_m.csv')))
hdr=rows[1]; ix={n:i for i,n in enumerate(hdr)}; data=rows[2:]
base=int(data[0][0],16)
agg=collections.defaultdict(lambda:[0,0,0])
ti=ts=tt=0
for r in data:
    ln=off2line.get(int(r[0],16)-base)
    inst=float(r[ix['Instructions Executed']] or 0); samp=float(r[ix['# Samples']] or 0); th=float(r[ix['Thread Instructions Executed']] or 0)
    a=agg[ln]; a[0]+=inst; a[1]+=samp; a[2]+=th; ti+=inst; ts+=samp; tt+=th
src=open('/root/repo/raer_b200/csrc/'+os.environ.get('SRCF','raer_fast.cuh')).read().split('\n')
r=list(csv.reader(open('/tmp/raw_m.csv'))); h,u,v=r[0],r[1],r[2]
for i,n in enumerate(h):
    if n in('gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_shared_atom.sum','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active'): print(n,u[i],v[i])
print("total warp inst %.3e thread inst %.3e samples %d"%(ti,tt,ts))
N=int(sys.argv[3]) if len(sys.argv)>3 else 30
for ln,a in sorted(agg.items(), key=lambda kv:-kv[1][int(os.environ.get("SORTCOL","0"))])[:N]:
    txt = src[ln[1]-1].strip()[:95] if ln and ln[0]==os.environ.get('SRCF','raer_fast.cuh') else str(ln)
    print("%-5s inst %5.1f%% thr/inst %4.1f samp %5.1f%%  %s"%(ln[1] if ln else '?',100*a[0]/ti,a[2]/max(a[0],1),100*a[1]/ts,txt))
