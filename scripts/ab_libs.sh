#!/bin/bash
# A/B of alternative builds of the library on ONE box (box-to-box variance is ~2 %): alternates the in-tree .so with
# build/libfoodrec_<variant>.so and prints step / e2e ms, the back-to-back propagation launch, and the HBM-regime launch.
# usage: scripts/ab_libs.sh base mb3 base mb3
L=multi-modal-food-recommendation_b200/libfoodrec_b200.so
cp $L /tmp/base.so
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so $L; else cp build/libfoodrec_$v.so $L; fi
  echo "== $v"
  python bench.py --no-extras --no-schgn --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'launch_us', round(r['avg_launch_us'],2), 'eval_ms', round(d['eval']['ms'],3))"
  python scripts/microbench_spmm_chunked.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v,3) for k,v in d.items() if k in ('full_ms','user_rows_ms','item_rows_ms')})"
done
cp /tmp/base.so $L
