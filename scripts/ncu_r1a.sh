set -x
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4096 -c 2400 --csv --log-file gpurun_out/launches_r1a.csv $CMD > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_deform_corr|k_assemble_fit" -s 60 -c 2 -o gpurun_out/prof_r1a_corr_asm $CMD > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pcg_spmv|k_pcg_update|k_precond|k_node_terms" -s 1260 -c 4 -o gpurun_out/prof_r1a_pcg $CMD > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
