#!/bin/bash
# tools/ab_time.sh WORKLOAD name1 name2 ... : bench.py once per libgi2d variant (tools/build_variant.sh), key numbers per line
WL=$1; shift
for n in "$@"; do
  if [ "$n" = "main" ]; then unset GI2D_LIB; else export GI2D_LIB=$PWD/gaussianimage_plus_b200/csrc/build/libgi2d_$n.so; fi
  python bench.py --workload $WL --steps ${STEPS:-1500} --warmup ${WARM:-300} --no-ref-cuda --no-cpu-baseline > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err || { echo "$n FAILED"; tail -5 gpurun_out/ab_$n.err; continue; }
  python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(f"{n:8s} value={d['value']:.0f} warm={d['value_l2_warm']:.0f} ({d['ms_per_step_l2_warm']*1e3:.2f} us) fps={d['render_fps']:.0f} e2e={d['e2e']['value']:.0f} raster_cold={r['kernel_ms']*1e3:.1f}us frac={r['frac']:.3f} kern={ {k:round(v*1e3,1) for k,v in r['step_kernel_ms'].items()} } psnr={d['psnr']:.3f}")
PY
done
