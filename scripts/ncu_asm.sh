set -x
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary"
ncu --set full --clock-control none --import-source on -k regex:"k_assemble_fit" -s 12 -c 1 -o gpurun_out/r1g_asm $CMD > gpurun_out/ncu_g1.log 2>&1
tail -n 3 gpurun_out/ncu_g1.log
