# ncu --set full capture of one launch of an octree traversal on cfg2 (numbers under ncu are never bench values)
# usage: bash tools/ncu_svo.sh <traversal> <kernel regex> <tag>
T=$1; K=$2; TAG=$3
ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 8 --launch-count 1 \
  -f -o gpurun_out/prof_${TAG} python bench.py --workload cfg2 --traversal $T --no-extras --steps 10 --warmup 3 > gpurun_out/ncu_${TAG}.log 2>&1
