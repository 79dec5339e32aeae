# A/B of library variants built with `python -m xenodon_b200.build --variant NAME "-D..."`
# usage: bash tools/ab_variants.sh "<bench args>" name1 name2 ...
ARGS=$1; shift
for v in "$@"; do
  XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_$v.so python bench.py $ARGS --no-extras > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$v.json'));print('$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_check'] if 'parity_check' in d else '')"
done
