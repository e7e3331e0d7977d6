python -m pytest tests/test_srfbn_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/tests_dc.log
python tools/layer_times.py --no-bw --summary > gpurun_out/lt_dc2.log 2>&1
