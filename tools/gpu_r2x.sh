set -x
mkdir -p gpurun_out
timeout 900 python tools/r2d2_learning_curve.py --out gpurun_out/r2x_r2d2_learning_curve.json 2>&1 | tail -22 | cut -c1-300
