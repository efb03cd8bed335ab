#!/bin/bash
# Planner-knob sweep on the 4K bench step (one GPU): dataflow segment length and dependency-tile size.
mkdir -p gpurun_out; out=gpurun_out/knob_sweep.log; : > $out
run() { echo "== $*" >> $out; env "$@" timeout 120 python bench.py --no-cpu --steps 2 --warmup 3 2>>$out | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('parity'))" >> $out 2>&1; }
run PB_X=0
run PB_IMPRINT_SEGMENT=32
run PB_IMPRINT_SEGMENT=16
run PB_IMPRINT_SEGMENT=128
run PB_PLAN_TILE=16
run PB_IMPRINT_SEGMENT=32 PB_PLAN_TILE=16
cat $out
