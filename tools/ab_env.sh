# A/B of an environment knob: bash tools/ab_env.sh "<bench args>" VAR v1 v2 ...
ARGS=$1; VAR=$2; shift; shift
for v in "$@"; do
  env $VAR=$v python bench.py $ARGS --no-extras > gpurun_out/abenv_$v.json 2> gpurun_out/abenv_$v.err
  python -c "import json;d=json.load(open('gpurun_out/abenv_$v.json'));print('$VAR=$v', d['value'], d['ms_per_step'])"
done
