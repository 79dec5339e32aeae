# A/B: skip-table brick size (run-time XN_SKIP_SHIFT: 3 = 8^3 voxels, 2 = 4^3, 1 = 2^3) on the smaller volumes
run() { tag=$1; wl=$2; trav=$3; shift 3
  env "$@" python bench.py --workload $wl $trav --steps 30 --warmup 3 --no-extras > gpurun_out/abv_${wl}_$tag.json 2> gpurun_out/abv_${wl}_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abv_${wl}_$tag.json'));print('$wl $trav $tag', d['value'], d['ms_per_step'])"; }
for s in 3 2 1; do run shift$s cfg1 "" XN_SKIP_SHIFT=$s; done
for s in 3 2 1; do run shift$s cfg2 "--traversal dda" XN_SKIP_SHIFT=$s; done
for s in 3 2; do run shift$s cfg3 "" XN_SKIP_SHIFT=$s; done
