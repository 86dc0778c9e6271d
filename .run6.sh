python -m pytest tests/test_gpu_band_split.py -m gpu -x -q > gpurun_out/pytest_band.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_band.log
