set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_qnet_tc_gpu.py -m gpu -q --timeout 120 2>&1 | tail -30 > gpurun_out/r2g_tc.txt; tail -30 gpurun_out/r2g_tc.txt
timeout 300 python tools/qnet_tc_bench.py --out gpurun_out/r2g_qnet_tc_bench.json 2>&1 | tail -12 | cut -c1-400
# tensor-pipe utilisation of the tcgen05 kernel (ncu, one launch)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_bf16_tc_kernel -s 6 -c 1 -o gpurun_out/r2g_dense_tc -f python tools/qnet_tc_bench.py > gpurun_out/r2g_ncu.log 2>&1; tail -3 gpurun_out/r2g_ncu.log
ncu -i gpurun_out/r2g_dense_tc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active','sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','launch__grid_size']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print({k:d.get(k) for k in want})
    print({k:v for k,v in d.items() if 'tensor' in k.lower()})
" | cut -c1-3000
