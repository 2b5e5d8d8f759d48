set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 200 -k image 2>&1 | tail -15 | tee gpurun_out/r3k_image_dp_test.txt
timeout 200 python tools/image_dp_check.py 2>&1 | grep IMAGEDP | tee gpurun_out/r3k_image_dp.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/image_dp_check.py 2>&1 | grep IMAGEDP | tee -a gpurun_out/r3k_image_dp.txt
