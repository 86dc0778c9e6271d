python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "4k" > gpurun_out/pytest_4k.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_4k.log
