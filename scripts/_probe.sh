P="python scripts/microbench_rank_probe.py"
$P nomask
$P mask 32 75776 20
$P mask 32 75776 80
