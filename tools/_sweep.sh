run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$name', 'ms/step', round(d['ms_per_step'],2), 'e2e_ms', round(d['e2e']['ms_per_step'],2), 'deconv', k['deconv8x8s4']['ms'], 'fused', k['fused_downtran_conv8x8s4']['ms'], 'clk', d['clocks']['sm_mhz'])" >> gpurun_out/ab_pdl.log
}
for rep in 1 2 3; do
run pdl1 VSR_PDL=1
run pdl0 VSR_PDL=0
done
