mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "tree or Tree or memory or Memory or per or sumtree" 2>&1 | tail -6
timeout 120 python tools/tree_update_bench.py 2>&1 | tail -6
timeout 300 python tools/sumtree_speedtest.py --skip python --out gpurun_out/sumtree_speedtest.json 2>&1 | tail -3 | cut -c1-330
