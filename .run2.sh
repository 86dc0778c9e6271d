python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err; echo "rc=$?"; tail -3 gpurun_out/bench_quick.err
