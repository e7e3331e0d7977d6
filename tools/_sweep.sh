python -m pytest tests/test_srfbn_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/tests_dc.log
python tools/layer_times.py --no-bw --summary > gpurun_out/lt_warp_arrive.log 2>&1
VSR_DECONV_DEBUG=7 python tools/layer_times.py --no-bw --summary 2>&1 | grep -E "deconv" >> gpurun_out/lt_warp_arrive.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s3c.json 2> gpurun_out/bench_s3c.err
