P="python scripts/microbench_rank_probe.py"
for m in nomask mask; do
$P $m
FR_TOPK_BOUND_STRIDE=4 $P $m
FR_TOPK_BOUND_STRIDE=8 $P $m
FR_TOPK_BOUND_STRIDE=16 $P $m
FR_TOPK_TWO_PASS=0 $P $m
done
FR_TOPK_TWO_PASS=0 FR_TOPK_PROBE=1 $P nomask
