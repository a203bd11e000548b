#!/bin/sh
# A/B of the fused 2-D kernel: current build vs the early-refill variant (same box, interleaved)
for rep in 1 2; do
  for v in cur early; do
    if [ $v = early ]; then export LORASTENCIL_LIB=$PWD/lorastencil_b200/var/early/lib/liblorastencil_b200.so; else unset LORASTENCIL_LIB; fi
    echo "== $v rep $rep"
    python profiles/run_shapes.py --shapes star2d3r,star2d1r --tb 3 --launches 30 --reps 5
  done
done
export LORASTENCIL_LIB=$PWD/lorastencil_b200/var/early/lib/liblorastencil_b200.so
echo "== race test, early variant"
timeout 300 python profiles/debug/slab_mismatch2.py star2d3r 10240,10240 3 4 40 | tail -3
