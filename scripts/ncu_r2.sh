# Round-2 evidence: the default bench line, the launch list of the same command and one --set full
# capture of the top kernels (first and third GN iteration of the first timed solve: unseeded and seeded
# correspondence pass, assembly, PCG).  Numbers printed under ncu are never bench values.
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4250 -c 700 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_deform_corr|k_assemble_fit|k_pcg_persistent" -s 90 -c 9 -o gpurun_out/r2_top_kernels $CMD > gpurun_out/r2_ncu1.log 2>&1
tail -n 2 gpurun_out/r2_ncu1.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_ref.err
tail -c 600 gpurun_out/r2_bench_1gpu.err gpurun_out/r2_ref.err
