python tools/stage_times.py 3840 2160 511
python tools/stage_times.py 1920 1080 255
python tools/band_bench.py > gpurun_out/band_bench.jsonl 2> gpurun_out/band_bench.err; echo "rc=$?"; cat gpurun_out/band_bench.jsonl; tail -3 gpurun_out/band_bench.err
