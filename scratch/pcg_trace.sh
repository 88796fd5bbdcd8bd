for g in 1 2 3 4; do
MVS_PCG_GATHER=$g MVS_PCG_TRACE=1 timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary --no-e2e 2>&1 | grep "pcg trace" | sed -n '25,28p' | awk -v t=$g '{print "gather "t": "$0}'
done
