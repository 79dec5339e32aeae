# always-fetch trips as the default: full GPU suite, default bench, knobs
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/abo_pytest.log 2>&1; tail -3 gpurun_out/abo_pytest.log
run() { # tag env...
  tag=$1; shift
  env "$@" python bench.py --workload $WL $TRAV --steps 30 --warmup 3 --no-extras > gpurun_out/abo_${WL}_$tag.json 2> gpurun_out/abo_${WL}_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abo_${WL}_$tag.json'));r=d['roofline'];print('$WL $tag', d['value'], d['ms_per_step'])"
}
TRAV=
for WL in cfg4 cfg3 cfg1; do
  run default XN_X=0
  for v in minb5 look1 look3 early1 early3; do run $v XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_$v.so; done
done
WL=cfg2; TRAV="--traversal dda"; run dda XN_X=0
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/abo_bench_default.json 2> gpurun_out/abo_bench_default.err
python -c "import json;d=json.load(open('gpurun_out/abo_bench_default.json'));print('default bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('parity_check'))"
