run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$name', 'ms/step', round(d['ms_per_step'],2), 'e2e_ms', round(d['e2e']['ms_per_step'],2), 'deconv', k['deconv8x8s4']['ms'], 'fused', k['fused_downtran_conv8x8s4']['ms'], 'clk', d['clocks']['sm_mhz'])" >> gpurun_out/ab_s3b.log
}
for rep in 1 2; do
run base_4_2 X=1
run cps1_11 VSR_DECONV_CPS=1 VSR_DECONV_STAGES=11
run cps2_4 VSR_DECONV_CPS=2 VSR_DECONV_STAGES=4
run cps2_5 VSR_DECONV_CPS=2 VSR_DECONV_STAGES=5
run sleep32 VSR_FUSED_DEBUG=8192
run sleep256 VSR_FUSED_DEBUG=65536
done
