"""Summarises an `ncu --page source --csv` export: stall mix overall and per window of SASS."""
import csv, sys, collections
src, raw = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
if raw:
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ['gpu__time_duration.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','launch__registers_per_thread ','dram__bytes_read.sum ','dram__bytes_write.sum ']
    for h,u,v in zip(hdr,units,vals):
        if any(h.startswith(w.strip()) for w in want): print(h,u,v)
rows = list(csv.reader(open(src)))
hdr = rows[1]; data = rows[2:]
ix = {h:i for i,h in enumerate(hdr)}
def f(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
tot=sum(f(r,'# Samples') for r in data)
print('instrs',len(data),'samples',tot)
keys=['stall_short_sb','stall_wait','stall_selected','stall_math','stall_barrier','stall_long_sb','stall_not_selected','stall_no_inst','stall_mio','stall_dispatch','stall_branch_resolving']
print(' '.join('%s:%.1f%%'%(k[6:],100*sum(f(r,k) for r in data)/tot) for k in keys))
W=int(sys.argv[3]) if len(sys.argv)>3 else 120
for s0 in range(0,len(data),W):
    seg=data[s0:s0+W]
    sm=sum(f(r,'# Samples') for r in seg)
    if sm/tot<0.008: continue
    ex=sum(f(r,'Instructions Executed') for r in seg)/len(seg)
    st={k:sum(f(r,k) for r in seg) for k in keys}
    ops=collections.Counter((r[ix['Source']].split()[1] if r[ix['Source']].startswith('@') else r[ix['Source']].split()[0]).split('.')[0] for r in seg if r[ix['Source']])
    top=' '.join('%s:%d'%(k[6:],100*v/sm) for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:4])
    print('%5d-%5d %5.2f%% exec %6.1fM | %s | %s'%(s0,s0+W,100*sm/tot,ex/1e6,top,' '.join('%s%d'%(k,v) for k,v in ops.most_common(5))))
