python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "subsampling" > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_sub.log
