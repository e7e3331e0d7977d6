python -m pytest tests/test_ops_gpu.py tests/test_fullsize_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/tests_splat.log
python tools/bench_configs.py --config c3 > gpurun_out/bench_c3_r01b.json 2> gpurun_out/bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_splat.csv python tools/profile_splat.py > gpurun_out/ps_ncu.log 2>&1
