python -m pytest tests/test_srfbn_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/tests_x2.log
python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q -k "c4" 2>&1 | tail -25 >> gpurun_out/tests_x2.log
