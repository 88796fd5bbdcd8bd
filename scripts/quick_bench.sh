# usage: quick_bench.sh TAG [extra bench args] — resident C3 bench, prints value + per-family microseconds per GN iteration
TAG=$1; shift
timeout 300 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e "$@" > gpurun_out/qb_$TAG.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/qb_$TAG.log").read().strip().splitlines()[-1])
print("$TAG", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "profiled", round(d["profiled_pass_ms_per_step"],3), d["graph"]["tuple_runs"], {k:round(v["ms_per_step"]/10*1000,1) for k,v in d["families"].items()})
PY
