python -m pytest tests/test_srfbn_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/tests_x2b.log
python tools/bench_configs.py --config c4 > gpurun_out/bench_c4_r01b.json 2> gpurun_out/bench_c4.err
