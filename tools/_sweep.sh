python -m pytest tests/test_flowimg_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/tests_flowimg.log
