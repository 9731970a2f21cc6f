#!/bin/bash
# Build variants of the sorted backward (msda_bwd_sorted.cu with extra -D flags) as separate libraries under build/exp/<name>/
# and, on the GPU box, time each one against the query-order kernel (tools/bwd_cell_check.py) in a process of its own.
#   tools/walker_variants.sh build  name1 "-DWALK_SLOTS=0" name2 "-DWALK_SLOTS=4 -DWALK_PREFETCH=1" ...
#   gpurun -- 'tools/walker_variants.sh run "B,S" "f32,bf16" name1 name2 ...'
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OBJ=$ROOT/vit-adapter_b200/lib/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mode=$1; shift
if [ "$mode" = build ]; then
  python $ROOT/vit-adapter_b200/build.py > /dev/null
  while [ $# -gt 0 ]; do
    name=$1; extra=$2; shift 2
    d=$ROOT/build/exp/$name; mkdir -p $d
    for tu in 0 1 2; do
      nvcc $FLAGS $extra -DMSDA_TU=$tu -c -o $d/sorted_$tu.o $ROOT/vit-adapter_b200/csrc/msda_bwd_sorted.cu &
    done
    wait
    others=$(ls $OBJ/*.o | grep -v msda_bwd_sorted)
    nvcc -shared -cudart static -o $d/libmsda_b200.so $others $d/sorted_0.o $d/sorted_1.o $d/sorted_2.o
    rm $d/*.o
    echo "built $name ($extra)"
  done
else
  variants=$1; dtypes=$2; shift 2
  mkdir -p $ROOT/gpurun_out
  for name in "$@"; do
    echo "== $name"
    MSDA_B200_LIB=$ROOT/build/exp/$name/libmsda_b200.so python $ROOT/tools/bwd_cell_check.py --mode sorted --variants $variants --dtypes $dtypes \
      --out $ROOT/gpurun_out/walker_$name.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    print(r['variant'], r['call'], r['dtype'], r['old_us'], r['cell_us'], r['speedup'], r['rel_err_gv_gl_ga'][0])"
  done
fi
