#!/usr/bin/env bash
# Same-box A/B of several builds of libvqb200.so.  Boxes of the pool differ by up to +-5 %, so small kernel effects
# must be compared inside ONE gpurun call.  Usage (in the build container):
#   1. build each variant and keep a copy:  cp vector_quantization_b200/libvqb200.so vector_quantization_b200/libvqb200_<name>.so
#   2. gpurun -- 'bash tools/ab_same_box.sh head nosharing ...'
# Every variant runs the assign micro-benchmark (random tokens) and the bench.py step (clustered tokens) twice,
# interleaved, through the VQB200_LIB override of vector_quantization_b200/_lib.py.
set -u
cd "$(dirname "$0")/.."
for rep in 1 2; do
  for v in "$@"; do
    export VQB200_LIB="$PWD/vector_quantization_b200/libvqb200_$v.so"
    a=$(timeout 100 python tools/bench_assign.py "cfg2 cos bf16x/pair" 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.readline())['ms_median'])")
    b=$(timeout 100 python tools/bench_assign.py "cfg3 l2 D8 1x1" 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.readline())['ms_median'])")
    c=$(timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), round(d['roofline']['kernel_ms'],5))")
    echo "$v  assign(pair, random tokens)=$a ms  assign(cfg3 shape)=$b ms  bench cfg2 step / kernel = $c ms"
  done
done
