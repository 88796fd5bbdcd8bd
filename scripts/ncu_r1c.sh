set -x
CMD="python scripts/per_iter_prof.py"
$CMD > gpurun_out/plain_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_assemble_fit|k_pcg_persistent" -s 44 -c 2 -o gpurun_out/prof_r1c_asm_pcg $CMD > gpurun_out/ncu_c1.log 2>&1
tail -n 3 gpurun_out/ncu_c1.log
