CMD="python bench.py --batch 64 --chunk 32 --steps 1 --warmup 1 --no-cpu-baseline --e2e-batch 32 --single-stream 1"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_support_match" -s 2 -c 1 -o gpurun_out/prof_match2 $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
