python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
for cfg in "64" "128"; do python bench.py --chunk $cfg --no-cpu-baseline --steps 3 > gpurun_out/bench_c$cfg.json 2>>gpurun_out/bench_sweep.err; echo "rc=$?"; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err
CMD="python bench.py --batch 64 --chunk 32 --steps 1 --warmup 1 --no-cpu-baseline --e2e-batch 32"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
