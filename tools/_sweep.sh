python -m pytest tests/test_ops_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/tests_lab.log
python tools/time_ops.py 2>&1 | grep -i "label\|copy" > gpurun_out/time_label.log
