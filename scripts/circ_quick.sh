#!/bin/bash
# quick c4 timing of the current build (RTB200_LAB_LIB=1 selects the lab library); extra env via "$@"
tag=${1:-q}; shift
out=gpurun_out; mkdir -p $out
env "$@" python bench.py --steps 20 --warmup 3 --no-e2e --no-gmres --no-cpu --no-solve --no-amr --no-hp > $out/${tag}.json 2> $out/${tag}.err || tail -5 $out/${tag}.err
python - <<PY
import json
d=json.loads(open("$out/${tag}.json").read().strip().splitlines()[-1])
print("${tag} $*: %.4f ms  %.1f GDoF/s  frac %.4f  checksum %.16g  noscat %.4f ms" % (d["ms_per_step"], d["value"], d["roofline"]["frac"], d["checksum"], d["roofline"]["no_scattering"]["ms_per_step"]))
PY
