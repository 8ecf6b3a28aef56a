P="python scripts/microbench_rank_probe.py"
$P nomask
$P mask
FR_TOPK_BOUND_STRIDE=4 $P mask
FR_TOPK_BOUND_STRIDE=1 $P mask
FR_TOPK_TWO_PASS=0 FR_TOPK_PROBE=1 $P nomask
