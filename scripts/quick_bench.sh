# usage: quick_bench.sh TAG  — resident C3 bench, prints value + per-family microseconds per GN iteration
timeout 200 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/qb_$1.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/qb_$1.log").read().strip().splitlines()[-1])
print("$1", round(d["value"],1), round(d["e2e"]["value"],1), d["config"]["tuple_runs"], {k:round(v["ms_per_step"]/10*1000,1) for k,v in d["families"].items()})
PY
