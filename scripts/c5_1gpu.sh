timeout 600 python bench.py --steps 2 --warmup 3 --side 913 --nodes 20000 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/c5_1gpu.log 2> gpurun_out/c5_1gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/c5_1gpu.log").read().strip().splitlines()[-1])
print(d["value"], {k:round(v["ms_per_step"]/10*1000,1) for k,v in d["families"].items()})
PY
