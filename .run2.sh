python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for cfg in "32 0" "32 1" "128 1"; do set -- $cfg; python bench.py --chunk $1 --single-stream $2 --no-cpu-baseline --steps 3 > gpurun_out/bench_c$1_s$2.json 2>>gpurun_out/bench_sweep.err; echo "rc=$?"; done
tail -5 gpurun_out/bench_sweep.err
