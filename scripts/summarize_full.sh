#!/bin/bash
# Key metrics of the `ncu --set full` captures (scripts/ncu_full.sh) -> profiles/r02_ncu_full_summary.txt
cd "$(dirname "$0")/.." || exit 1
METRICS='gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__pipe_tensor_op.*pct|sm__warps_active.avg.pct_of_peak|launch__registers_per_thread|launch__grid_size|launch__block_size|smsp__issue_active.avg.pct|sm__inst_executed_pipe_xu.avg.pct|sm__pipe_xu_cycles_active.avg.pct|lts__t_sector_hit_rate.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum |smsp__warp_issue_stalled.*_per_warp_active.pct|sm__throughput.avg.pct|lts__throughput.avg.pct|launch__shared_mem_per_block_dynamic|l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed|smsp__average_warps_issue_stalled_(long_scoreboard|wait|short_scoreboard|barrier|mio_throttle|math_pipe_throttle)_per_issue_active'
for k in conv lin geglu gn attn; do
  f=gpurun_out/r02_full_$k.ncu-rep
  [ -f "$f" ] || continue
  echo "== $k ($f)"
  ncu -i "$f" --page raw --csv 2>/dev/null | python3 -c "
import csv, sys, re
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    names, units, vals = rows[0], rows[1], rows[2]
    pat = re.compile(r'$METRICS')
    for n, u, v in zip(names, units, vals):
        if n == 'Kernel Name': print('kernel', v[:110])
        if pat.search(n): print(f'  {n:78s} {v:>16s} {u}')
"
done
