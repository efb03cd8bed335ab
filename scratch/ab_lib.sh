#!/bin/bash
# A/B of two builds of the library on one box: painty_b200/libpainty_b200_base.so (A) against libpainty_b200.so (B).
mkdir -p gpurun_out; out=gpurun_out/ab_lib.log; : > $out
lib=painty_b200/libpainty_b200.so
cp $lib /tmp/new.so
for v in base new; do
  if [ $v = base ]; then cp painty_b200/libpainty_b200_base.so $lib; else cp /tmp/new.so $lib; fi
  echo "== $v micro" >> $out
  timeout 100 python scratch/imprint_micro.py 30,45,64,81,112,129,151 400 2>&1 | grep "r=" >> $out
  echo "== $v bench" >> $out
  timeout 120 python bench.py --no-cpu --steps 2 --warmup 3 2>>$out | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['imprint']['ms_each_step'], d.get('parity'))" >> $out 2>&1
done
echo "== pytest (new)" >> $out
timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 >> $out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> $out
cat $out
