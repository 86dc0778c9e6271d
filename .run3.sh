CMD="python bench.py --batch 32 --chunk 32 --steps 1 --warmup 1 --no-cpu-baseline --e2e-batch 32 --single-stream 1"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_dense|k_support_match|k_descriptor|k_reproject|k_raster" -s 6 -c 6 -o gpurun_out/prof_sel2 $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
