python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err; echo "rc=$?"; tail -3 gpurun_out/bench_quick.err
CMD="python bench.py --batch 32 --chunk 32 --steps 1 --warmup 1 --no-cpu-baseline --e2e-batch 32 --single-stream 1"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -s 31 -c 32 -o gpurun_out/prof_all $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
