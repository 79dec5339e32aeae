run() { tag=$1; wl=$2; shift 2
  env "$@" python bench.py --workload $wl --traversal esvo --steps 30 --warmup 3 --no-extras > gpurun_out/abu_${wl}_$tag.json 2> gpurun_out/abu_${wl}_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abu_${wl}_$tag.json'));print('$wl esvo $tag', d['value'], d['ms_per_step'])"; }
V=$PWD/xenodon_b200/variants/libxenodon_b200
for wl in cfg2 cfg4e; do
  run mb6 $wl XN_X=0
  for v in emb4 emb5 emb7; do run $v $wl XN_LIBRARY=${V}_$v.so; done
done
