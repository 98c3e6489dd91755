for f in experiments/variants/lib_*.so; do
  v=$(basename $f .so)
  TRL_LIB_PATH=$PWD/$f timeout 300 python -m pytest tests/test_gpu_pnet_hybrid.py -x -q -m gpu 2>&1 | tail -1
  TRL_LIB_PATH=$PWD/$f timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_$v.json 2> gpurun_out/r2q_$v.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2q_$v.json")); print("$v", round(d["value"]), round(d["stages"]["pnet"]["ms_per_step"],3), d["result"]["score"])
except Exception as e: print("$v", "FAILED", e)
P
done
