python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_seq.log
python tools/bench_configs.py --config c3 > gpurun_out/bench_c3_r01.json 2> gpurun_out/bench_c3.err
python tools/bench_configs.py --config c4 > gpurun_out/bench_c4_r01.json 2> gpurun_out/bench_c4.err
python tools/bench_configs.py --config c5 > gpurun_out/bench_c5_n1_r01.json 2> gpurun_out/bench_c5.err
