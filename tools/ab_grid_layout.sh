# A/B of the grid residency layouts (un-profiled bench runs)
# usage: LAYOUTS="linear texture" bash tools/ab_grid_layout.sh [steps] [workloads...]
STEPS=${1:-60}; shift
WLS=${@:-cfg1 cfg3 cfg4}
for wl in $WLS; do
  for lay in ${LAYOUTS:-linear bricked texture}; do
    XN_GRID_LAYOUT=$lay python bench.py --workload $wl --no-extras --steps $STEPS > gpurun_out/s2_${wl}_${lay}.json 2> gpurun_out/s2_${wl}_${lay}.err
    python -c "import json;d=json.load(open('gpurun_out/s2_${wl}_${lay}.json'));print('$wl $lay', d['value'], d['ms_per_step'], d['config'].get('grid_layout'), d['roofline']['frac'])"
  done
done
