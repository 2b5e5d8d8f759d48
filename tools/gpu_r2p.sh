set -x
mkdir -p gpurun_out
timeout 600 python tools/learning_curve.py > gpurun_out/r2z_learning_curve.json 2> gpurun_out/r2z_lc.err; tail -c 300 gpurun_out/r2z_lc.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2z_learning_curve.json'))
c=d.get('curve', d)
print(len(c)); print(c[:3]); print(c[-3:])
PY
