mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k returns 2>&1 | tail -3
timeout 300 python tools/returns_bench.py --out gpurun_out/returns_bench.json 2>&1 | tail -1 | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'returns_scan_kernel' -s 8 -c 1 -o gpurun_out/r1_j_returns -f python tools/returns_bench.py > gpurun_out/ncu_returns.log 2>&1
python tools/ncu_summary.py gpurun_out/r1_j_returns.ncu-rep > gpurun_out/r1_j_returns_ncu_summary.json; cat gpurun_out/r1_j_returns_ncu_summary.json
ncu -i gpurun_out/r1_j_returns.ncu-rep --page details --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; mi=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
for r in rows[1:]:
    if r[mi] in ('Duration','DRAM Throughput','Memory Throughput','L2 Cache Throughput','Achieved Occupancy','Theoretical Occupancy','Issue Slots Busy','Max Bandwidth','Mem Busy','L1/TEX Hit Rate','L2 Hit Rate','Registers Per Thread','Block Limit Registers','Block Limit Shared Mem'): print(r[mi], r[vi], r[ui])
"
