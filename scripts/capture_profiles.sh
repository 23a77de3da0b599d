#!/bin/bash
# Round profile capture on the GPU box (run under gpurun, one GPU).  Every ncu pass follows a plain run of the same
# command that exited 0.  Outputs land in gpurun_out/; scripts/make_profile_summary.py turns the reports into the text
# summaries committed under profiles/.
set -u
TAG=${1:-r01}
OUT=gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --targets 100000 --cpu-seconds 0 --no-micro"
$BENCH > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain bench failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_list.log 2>&1
# k = 3 launch of the first sweep (third lbfgs launch)
ncu --set full --clock-control none --import-source on -k regex:lbfgs_kernel --launch-skip 2 --launch-count 1 -f -o $OUT/${TAG}_lbfgs_k3 $BENCH > $OUT/${TAG}_ncu_lbfgs.log 2>&1
python scripts/cov_one.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:coverage_kernel --launch-count 2 -f -o $OUT/${TAG}_coverage python scripts/cov_one.py > $OUT/${TAG}_ncu_cov.log 2>&1
python scripts/smush_one.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:smush_loss_grad --launch-count 2 -f -o $OUT/${TAG}_smush_adj python scripts/smush_one.py > $OUT/${TAG}_ncu_smush.log 2>&1
SLAM_B200_LPP=4 python scripts/k2_one.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:loss_grad_kernel --launch-count 2 -f -o $OUT/${TAG}_loss_grad python scripts/k2_one.py > $OUT/${TAG}_ncu_k2.log 2>&1
ls -la $OUT | grep $TAG
