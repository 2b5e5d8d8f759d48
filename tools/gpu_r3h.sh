set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:t3_gemm_kernel -s 58 -c 3 -o gpurun_out/r3h_t3 -f python tools/image_prof.py 256 > gpurun_out/r3h_ncu.log 2>&1; tail -3 gpurun_out/r3h_ncu.log
ncu -i gpurun_out/r3h_t3.ncu-rep --page raw --csv > gpurun_out/r3h_t3_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r3h_t3_raw.csv')))
h=rows[0]
want=['Kernel Name','Grid Size','gpu__time_duration.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_membar_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_sleeping_per_warp_active.pct','smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','lts__t_bytes.sum','sm__inst_executed_pipe_tensor.sum','smsp__cycles_active.avg','smsp__issue_active.avg.pct_of_peak_sustained_active']
idx={n:i for i,n in enumerate(h)}
for r in rows[2:]:
    print('----')
    for w in want:
        if w in idx: print(w, '=', r[idx[w]])
stall=[n for n in h if 'warp_issue_stalled' in n and n.endswith('_per_warp_active.pct')]
for r in rows[2:3]:
    print(sorted([(float(r[idx[n]].replace(',','') or 0), n.replace('smsp__warp_issue_stalled_','').replace('_per_warp_active.pct','')) for n in stall], reverse=True)[:8])
PY
