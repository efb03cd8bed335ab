#!/bin/bash
# ncu evidence of a round (run under gpurun on ONE GPU): launch list of a short bench, --set full captures of the imprint
# kernel (r = 30, 112, 151), the compose kernel and the texture kernel. Reports land in gpurun_out/.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${1}_launches.csv python bench.py --strokes 300 --steps 2 --warmup 1 --no-cpu > gpurun_out/${1}_launches_bench.log 2>&1
for r in 30 112 151; do
  $NCU --set full --import-source on -k regex:imprint_kernel --launch-skip 1 -c 1 -f -o gpurun_out/${1}_imprint_r$r python scratch/imprint_micro.py $r 200 0.79 > gpurun_out/${1}_imprint_r$r.log 2>&1
done
$NCU --set full --import-source on -k regex:km_compose_kernel --launch-skip 1 -c 1 -f -o gpurun_out/${1}_compose_f32 python scratch/compose_only.py > gpurun_out/${1}_compose.log 2>&1
$NCU --set full --import-source on -k regex:texture_kernel -c 1 -f -o gpurun_out/${1}_texture python scratch/texture_only.py 2000 > gpurun_out/${1}_texture.log 2>&1
ls -la gpurun_out/*.ncu-rep
