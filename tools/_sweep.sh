for d in 0 8 1; do
echo "debug $d" >> gpurun_out/lt_dc8.log
VSR_DECONV_DEBUG=$d python tools/layer_times.py --no-bw --summary 2>&1 | grep -E "deconv" >> gpurun_out/lt_dc8.log
done
