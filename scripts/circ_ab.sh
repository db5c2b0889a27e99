#!/bin/bash
# A/B of the k_apply_circ7 variants on the c4 workload + the parity tests that cover the kernel.
# usage (on the GPU box): bash scripts/circ_ab.sh <tag>
tag=${1:-ab}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_circ.py tests/test_gpu_apply.py tests/test_gpu_fullsize.py -x -q -m gpu > $out/${tag}_pytest.log 2>&1
tail -3 $out/${tag}_pytest.log
for v in "0 0" "1 0" "0 1" "1 1"; do
  set -- $v
  RTB200_CIRC7_EO=$1 RTB200_CIRC7_INNER=$2 python bench.py --steps 20 --warmup 3 --no-e2e --no-gmres --no-cpu --no-solve --no-amr --no-hp \
     > $out/${tag}_bench_eo$1_inner$2.json 2> $out/${tag}_bench_eo$1_inner$2.err
  python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_eo$1_inner$2.json").read().strip().splitlines()[-1])
print("eo=$1 inner=$2: %.4f ms  %.1f GDoF/s  frac %.4f  checksum %.16g  noscat %.4f ms" % (d["ms_per_step"], d["value"], d["roofline"]["frac"], d["checksum"], d["roofline"]["no_scattering"]["ms_per_step"]))
PY
done
