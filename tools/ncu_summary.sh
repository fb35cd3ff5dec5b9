#!/bin/bash
# usage: tools/ncu_summary.sh <report.ncu-rep>  -- headline metrics, stall ratios, pipe utilisation, FP64 op counts, shared /
# local memory instruction counts and bank conflicts, instruction-cache hit rate of the first kernel in the report
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
def val(name):
    if name not in h: return None
    try: return float(rows[2][h.index(name)].replace(',',''))
    except ValueError: return None
out=[]
for i,k in enumerate(h):
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
        try: out.append((float(rows[2][i].replace(',','')),k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
        except: pass
print('   stall cycles per issue:', ', '.join(f'{k} {v:.2f}' for v,k in sorted(out,reverse=True)[:8]))
for name in ('dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','lts__t_bytes.sum'):
    if name in h: print('  ', name, rows[2][h.index(name)], rows[1][h.index(name)])
# pipe utilisation (% of peak, sustained while active)
pipes=[]
for p in ('fp64','lsu','alu','fma','xu','cbu','adu','uniform'):
    v=val(f'sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active')
    if v is not None: pipes.append(f'{p} {v:.1f}%')
print('   pipe utilisation (pct of peak sustained active):', ', '.join(pipes))
# executed FP64 work: thread-instructions per cycle (summed over the SMs) x elapsed cycles; a DFMA counts 2 flops
cyc=val('sm__cycles_elapsed.max') or val('sm__cycles_elapsed.avg')
ops={}
for op in ('dfma','dadd','dmul'):
    v=val(f'smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed')
    if v is not None and cyc: ops[op]=v*cyc
if ops:
    flop=2*ops.get('dfma',0)+ops.get('dadd',0)+ops.get('dmul',0)
    print('   executed FP64 thread-instructions per launch:', ', '.join(f'{k} {v:.3e}' for k,v in ops.items()), f'=> {flop:.3e} flop executed (compare with the algorithmic FLOPs per launch in the bench line)')
ti=val('smsp__thread_inst_executed.sum'); wi=val('smsp__inst_executed.sum')
if ti and ops: print(f'   FP64 share of executed thread-instructions: {100*sum(ops.values())/ti:.1f}%')
for name,label in (('smsp__sass_inst_executed_op_shared_ld.sum','shared loads'),('smsp__sass_inst_executed_op_shared_st.sum','shared stores'),('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','shared bank conflicts'),('smsp__sass_inst_executed_op_local_ld.sum','local (spill) loads'),('smsp__sass_inst_executed_op_local_st.sum','local (spill) stores')):
    v=val(name)
    if v is not None: print(f'   {label}: {v:.3e} warp-instructions' + (f' ({100*v/wi:.1f}% of all)' if wi and 'conflict' not in label else ''))
v=val('sm__icc_request_hit_rate.pct')
if v is not None: print(f'   instruction cache (L1.5) hit rate: {v:.1f}%')
"
