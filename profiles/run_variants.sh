#!/bin/bash
# developer tool (GPU box): bench every gpurun_variants/*.so, print the phase times
for so in gpurun_variants/*.so; do
  n=$(basename $so .so)
  for wl in "$@"; do
  AT3D_B200_LIB=$PWD/$so python bench.py --steps 3 --warmup 2 --no-cpu --workload $wl > gpurun_out/var_${n}_$wl.json 2> gpurun_out/var_${n}_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_${n}_$wl.json"))
    print("$n $wl step %.2f ms fwd %.2f adj %.2f render %.2f"%(d["ms_per_step"], d["phases_ms"]["forward"], d["phases_ms"]["adjoint"], d["render"]["kernel_ms"]))
except Exception as e:
    print("$n $wl failed", e)
PY
  done
done
