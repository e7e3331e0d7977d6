TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 tools/bench_configs.py --config c5 > gpurun_out/bench_c5_n8_r01.json 2> gpurun_out/bench_c5_n8.err
$TR --nproc-per-node 4 --master-port 29512 tools/bench_configs.py --config c5 > gpurun_out/bench_c5_n4_r01.json 2> gpurun_out/bench_c5_n4.err
$TR --nproc-per-node 2 --master-port 29513 tools/bench_configs.py --config c5 > gpurun_out/bench_c5_n2_r01.json 2> gpurun_out/bench_c5_n2.err
$TR --nproc-per-node 8 --master-port 29514 tools/bench_configs.py --config c5 --exact-chunks > gpurun_out/bench_c5_n8_exact_r01.json 2> gpurun_out/bench_c5_n8e.err
$TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/bench_r01_n8b.json 2> gpurun_out/bench_r01_n8b.err
$TR --nproc-per-node 8 --master-port 29516 tools/bench_configs.py --config c4 > gpurun_out/bench_c4_n8_r01.json 2> gpurun_out/bench_c4_n8.err
