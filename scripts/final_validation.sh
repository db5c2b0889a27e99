#!/bin/bash
# Round-end validation on one B200: full GPU suite, default bench line, smoke, ncu launch list of the
# bench command and a full ncu capture of the dominant kernel.  usage: bash scripts/final_validation.sh <tag>
tag=${1:-final}
out=gpurun_out; mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; tail -2 $out/${tag}_pytest_gpu.log
python bench.py > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err; tail -c 400 $out/${tag}_bench_c4.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; tail -c 300 $out/${tag}_bench_reference.json; echo
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-solve --no-e2e --no-hp --no-amr > $out/${tag}_launches_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_apply_circ7 --launch-skip 6 --launch-count 1 -f -o $out/${tag}_circ7 python bench.py --steps 4 --warmup 3 --no-e2e --no-gmres --no-cpu --no-solve --no-amr --no-hp > $out/${tag}_ncu.log 2>&1; tail -1 $out/${tag}_ncu.log
