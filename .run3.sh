CMD="python bench.py --batch 32 --chunk 32 --steps 1 --warmup 1 --no-cpu-baseline --e2e-batch 32 --single-stream 1"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_dense|k_gap|k_mean|k_median" -s 9 -c 9 -o gpurun_out/prof_sel $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
