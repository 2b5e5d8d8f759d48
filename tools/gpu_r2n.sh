set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cabi.py tests/test_srl_classes_gpu.py -m gpu -q --timeout 300 -k "memory or tree or Memory or cabi or library" 2>&1 | tail -5
timeout 300 python tools/sumtree_speedtest.py --skip python --out gpurun_out/r2n_sumtree_speedtest.json 2>&1 | tail -5 | cut -c1-400
timeout 300 python tools/seam_profile.py 2>&1 | tail -4
