# ncu --set full capture of one DDA launch per layout (numbers printed under ncu are never bench values)
# usage: bash tools/ncu_grid_layout.sh <workload> <layouts...>
WL=${1:-cfg4}; shift
for lay in ${@:-linear bricked texture}; do
  XN_GRID_LAYOUT=$lay ncu --set full --clock-control none --import-source on -k regex:dda --launch-skip 8 --launch-count 1 \
    -f -o gpurun_out/prof_dda_${WL}_${lay}_s2 python bench.py --workload $WL --no-extras --steps 10 --warmup 3 > gpurun_out/ncu_${WL}_${lay}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
