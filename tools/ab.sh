#!/bin/bash
# A/B of two builds of the library on the SAME box (the boxes of the pool differ by several percent and sit in
# sw_power_cap to different degrees): tools/ab.sh <workload>... ; expects build/ab/old.so and build/ab/new.so
LIB=srcdsp_b200/lib/libsrcdsp_b200.so
for rep in 1 2 3; do
  for v in old new; do
    cp build/ab/$v.so $LIB
    for w in "$@"; do
      printf "%s %s rep%d: " $v $w $rep
      timeout 250 python bench.py --workload $w --steps 10 --warmup 3 --no-ddc --no-cpu --no-e2e 2>&1 | tail -1 | \
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
    done
  done
done
cp build/ab/new.so $LIB
