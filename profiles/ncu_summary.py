import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; vals=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sectors_srcunit_tex_op_red.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__inst_executed_pipe_lsu.sum','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','l1tex__t_requests_pipe_lsu_mem_global_op_red.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_op_atom.sum','lts__t_sectors_op_red.sum']
for i,h in enumerate(hdr):
    if h in want: print(f'{h:72s} {rows[1][i]:10s} {vals[i]}')
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines())); hdr=rows[1]; ci={h:i for i,h in enumerate(hdr)}; data=rows[2:]
def f(x):
    try: return float(x)
    except: return 0.0
tot=sum(f(r[ci['# Samples']]) for r in data)
print('--- stall samples total',tot)
for s_ in ['stall_long_sb','stall_lg','stall_wait','stall_not_selected','stall_selected','stall_math','stall_short_sb','stall_branch_resolving','stall_drain','stall_mio','stall_no_inst','stall_dispatch']:
    print(f'  {s_:24s} {sum(f(r[ci[s_]]) for r in data):8.0f}')
from collections import defaultdict
agg=defaultdict(lambda:[0,0,0,0,0])
for r in data:
    t=r[ci['Source']].split()
    if not t: continue
    op=t[1] if t[0].startswith('@') else t[0]
    op=op.split('.')[0]
    a=agg[op]; a[0]+=f(r[ci['# Samples']]); a[1]+=f(r[ci['Instructions Executed']]); a[2]+=f(r[ci['stall_lg']]); a[3]+=f(r[ci['stall_long_sb']]); a[4]+=f(r[ci['L1 Tag Requests Global']])
print('op        samples  warp_instr  stall_lg stall_long_sb L1tagreq')
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:12]:
    print(f'{k:9s} {v[0]:7.0f} {v[1]:11.0f} {v[2]:8.0f} {v[3]:8.0f} {v[4]:10.0f}')
print('total warp instr', sum(v[1] for v in agg.values()))
