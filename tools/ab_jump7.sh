set -x
run() { # tag env...
  tag=$1; shift
  env "$@" python bench.py --workload $WL $TRAV --steps 30 --warmup 3 --no-extras > gpurun_out/abp_${WL}_$tag.json 2> gpurun_out/abp_${WL}_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abp_${WL}_$tag.json'));r=d['roofline'];print('$WL $tag', d['value'], d['ms_per_step'])"
}
TRAV=
for WL in cfg4 cfg3 cfg1; do
  run default XN_X=0
  for v in look4 look6 look4e1 j4 j10; do run $v XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_$v.so; done
done
