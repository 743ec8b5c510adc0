#!/bin/bash
# usage: summarise_profile.sh gpurun_out/X.ncu-rep profiles/NAME   -> NAME.details.txt + NAME.metrics.csv
REP=$1; OUT=$2
ncu -i $REP --page details 2>/dev/null | grep -vE "^\s*$" > $OUT.details.txt
ncu -i $REP --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr,units,vals=rows[0],rows[1],rows[2]
keep=('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed.avg.per_cycle_active','Kernel Name','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')
w=csv.writer(sys.stdout); w.writerow(['metric','unit','value'])
for h,u,v in zip(hdr,units,vals):
    if h in keep: w.writerow([h,u,v])
" > $OUT.metrics.csv
