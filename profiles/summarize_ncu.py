"""Summarise an ncu report: python profiles/summarize_ncu.py raw.csv source.csv  (CSV pages from
`ncu -i X.ncu-rep --page raw --csv` and `--page source --csv`).  Prints per-kernel headline metrics, top stall
reasons and the SASS sections that account for the executed instructions / stall samples."""
import csv, sys
raw, sass = sys.argv[1], sys.argv[2]
rows=list(csv.reader(open(raw))); hdr=rows[0]; units=rows[1]; data=rows[2:]
want=['gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for d in data:
    print('-----', d[hdr.index('Kernel Name')][:50])
    for w in want:
        if w in hdr: i=hdr.index(w); print(f'  {w:66s} {d[i][:40]} {units[i]}')
    st=[(h,d[i]) for i,h in enumerate(hdr) if 'smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    print('  stalls:', ', '.join(f"{h.split('stalled_')[1].split('_per')[0]}={float(v):.2f}" for h,v in sorted(st,key=lambda x:-float(x[1].replace(',','') or 0))[:7]))
rows=list(csv.reader(open(sass)))
blocks=[]; cur=None
for r in rows:
    if len(r)>3 and r[0]=='Address': cur={'hdr':r,'rows':[]}; blocks.append(cur)
    elif cur is not None and len(r)==len(cur['hdr']): cur['rows'].append(r)
for bi,b in enumerate(blocks):
    h=b['hdr']; iS=h.index('Source'); iE=h.index('Instructions Executed'); iSamp=h.index('# Samples')
    tot=sum(int(r[iE]) for r in b['rows']); tots=sum(int(r[iSamp]) for r in b['rows'])
    print('=== kernel',bi,'total inst %.1fM'%(tot/1e6),'samples',tots)
    sec=[]
    for idx,r in enumerate(b['rows']):
        e=int(r[iE])
        if sec and abs(sec[-1][2]-e)<=0.02*max(e,1): sec[-1][1]=idx; sec[-1][3]+=e; sec[-1][4]+=int(r[iSamp])
        else: sec.append([idx,idx,e,e,int(r[iSamp])])
    for s in sec:
        if s[3]>0.004*tot or s[4]>0.01*tots: print(f'  lines {s[0]:4d}-{s[1]:4d} n={s[1]-s[0]+1:3d} exec/inst {s[2]/1e6:7.2f}M total {s[3]/1e6:7.1f}M {100*s[3]/tot:5.1f}%  samples {100*s[4]/tots:5.1f}%  | {b["rows"][s[0]][iS][:50]}')
