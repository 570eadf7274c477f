#!/bin/bash
# Build kernel variants on the GPU box and bench each (experiment helper).
# usage: sweep_variants.sh "<nvcc -D flags>[;ENV=VAL ...]" ...   e.g. "-DFAB_PROF" "-DFAB_MIN_CTAS=2;FAB_FORCE_TILE=8"
mkdir -p gpurun_out
for spec in "$@"; do
  v="${spec%%;*}"; envs=""; [[ "$spec" == *";"* ]] && envs="${spec#*;}"
  echo "=== variant: $v  env: $envs"
  FAB_NVCC_FLAGS="$v" python fab_torch_b200/csrc/build.py --force > gpurun_out/build_variant.log 2>&1 || { tail -5 gpurun_out/build_variant.log; continue; }
  env $envs python -c "
import fab_torch_b200 as fb
f = fb.B200RealNVP(32,10,10); print('tile particles for n=2048:', fb._lib.lib().fab_tile_particles(f.desc(), 2048))"
  env $envs timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('particles/s %.0f  ms/step %.2f  kernel_ms %.3f  ffma_frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['pipe_frac']))"
done
