set -x
CMD="python scripts/per_iter_prof.py"
$CMD > gpurun_out/plain_d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_deform_corr" -s 20 -c 1 -o gpurun_out/prof_r1d_corr0 $CMD > gpurun_out/ncu_d1.log 2>&1
tail -n 3 gpurun_out/ncu_d1.log
