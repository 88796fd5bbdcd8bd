# Round-1 evidence (final code): launch list of the bench command + one --set full capture of the
# top kernels (first two GN iterations of the first timed solve: unseeded and seeded correspondence
# search, assembly, PCG) + the ARAP CG kernels.
set -x
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
$CMD > gpurun_out/plain_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4096 -c 900 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/ncu_f0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_deform_corr|k_assemble_fit|k_pcg_persistent" -s 90 -c 6 -o gpurun_out/r1_top_kernels $CMD > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_arap_spmv|k_arap_update|k_arap_dir|k_arap_local" -s 40 -c 4 -o gpurun_out/r1_arap_kernels python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --c2-points 100000 > gpurun_out/ncu_f2.log 2>&1
tail -n 2 gpurun_out/ncu_f0.log gpurun_out/ncu_f1.log gpurun_out/ncu_f2.log
cat gpurun_out/plain_final.log | cut -c1-300
