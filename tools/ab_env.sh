#!/bin/bash
# tools/ab_env.sh WORKLOAD "ENV=1 ..." : bench.py with extra environment, key numbers
WL=$1; shift
for e in "$@"; do
  env $e python bench.py --workload $WL --steps ${STEPS:-1500} --warmup ${WARM:-300} --no-ref-cuda --no-cpu-baseline > gpurun_out/abe.json 2> gpurun_out/abe.err || { echo "$e FAILED"; tail -5 gpurun_out/abe.err; continue; }
  python - "$e" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/abe.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(f"{sys.argv[1]:20s} value={d['value']:.0f} warm={d['value_l2_warm']:.0f} ({d['ms_per_step_l2_warm']*1e3:.2f} us) fps={d['render_fps']:.0f} e2e={d['e2e']['value']:.0f} psnr={d['psnr']:.3f} raster_b2b={(r.get('kernel_ms_back_to_back_l2_warm') or 0)*1e3:.2f}us frac={r['frac']:.3f} kern={ {k:round(v*1e3,1) for k,v in r['step_kernel_ms'].items()} }")
PY
done
