#!/bin/bash
# usage: tools/ncu_summary.sh <report.ncu-rep>  -- headline metrics + stall ratios of the first kernel in the report
rep=$1
ncu -i $rep --page details --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; ix={k:i for i,k in enumerate(h)}
keep=['Duration','Executed Ipc Active','No Eligible','Eligible Warps Per Scheduler','Active Warps Per Scheduler','Achieved Occupancy','Theoretical Occupancy','Registers Per Thread','Avg. Active Threads Per Warp','Warp Cycles Per Issued Instruction','Executed Instructions','Dynamic Shared Memory Per Block','Block Limit Registers','Block Limit Shared Mem','Local Load','Local Store']
print(rows[1][ix['Kernel Name']][:60])
for r in rows[1:]:
    n=r[ix['Metric Name']]
    if n in keep: print('  ', n.ljust(44), r[ix['Metric Value']], r[ix['Metric Unit']])
"
ncu -i $rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
out=[]
for i,k in enumerate(h):
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
        try: out.append((float(rows[2][i].replace(',','')),k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
        except: pass
print('   stall cycles per issue:', ', '.join(f'{k} {v:.2f}' for v,k in sorted(out,reverse=True)[:8]))
for name in ('dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','lts__t_bytes.sum'):
    if name in h: print('  ', name, rows[2][h.index(name)], rows[1][h.index(name)])
"
