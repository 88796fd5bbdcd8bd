set -x
CMD="python scripts/per_iter_prof.py"
$CMD > gpurun_out/plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_deform_corr" -s 20 -c 2 -o gpurun_out/prof_r1b_corr $CMD > gpurun_out/ncu_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pcg_persistent|k_assemble_fit" -s 40 -c 2 -o gpurun_out/prof_r1b_pcg_asm $CMD > gpurun_out/ncu_b2.log 2>&1
tail -3 gpurun_out/ncu_b1.log gpurun_out/ncu_b2.log
