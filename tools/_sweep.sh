python -m pytest tests/test_ref_ops_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/tests_corr.log
python tools/time_correlation.py > gpurun_out/time_correlation_r01.log 2>&1
