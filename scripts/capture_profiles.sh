#!/bin/bash
# Round profile capture on the GPU box (run under gpurun, one GPU).  Every ncu pass follows a plain run of the same
# command that exited 0.  Outputs land in gpurun_out/; scripts/make_profile_summary.py turns the reports into the text
# summaries committed under profiles/.
#   usage: scripts/capture_profiles.sh <tag> [what ...]     what = list k5 k5k1 k5c cov smush k2 weyl   (default: all)
set -u
TAG=${1:-r02}
shift || true
WHAT=${*:-list k5 k5k1 k5c cov smush k2 weyl}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --targets 100000 --cpu-seconds 0 --no-micro"
FULL="ncu --set full --clock-control none --import-source on -f"
has() { [[ " $WHAT " == *" $1 "* ]]; }
$BENCH > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain bench failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
has list && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_list.log 2>&1
# k = 3 launch of the first sweep (third lbfgs launch), and the k = 1 launch
has k5 && $FULL -k regex:lbfgs_kernel --launch-skip 2 --launch-count 1 -o $OUT/${TAG}_lbfgs_k3 $BENCH > $OUT/${TAG}_ncu_lbfgs.log 2>&1
has k5k1 && $FULL -k regex:lbfgs_kernel --launch-skip 0 --launch-count 1 -o $OUT/${TAG}_lbfgs_k1 $BENCH > $OUT/${TAG}_ncu_lbfgs1.log 2>&1
has k5c && python scripts/k5c_one.py > /dev/null 2>&1 && $FULL -k regex:adj_lbfgs --launch-count 1 -o $OUT/${TAG}_k5c python scripts/k5c_one.py > $OUT/${TAG}_ncu_k5c.log 2>&1
has cov && python scripts/cov_one.py > /dev/null 2>&1 && $FULL -k regex:coverage_kernel --launch-count 2 -o $OUT/${TAG}_coverage python scripts/cov_one.py > $OUT/${TAG}_ncu_cov.log 2>&1
has smush && python scripts/smush_one.py > /dev/null 2>&1 && $FULL -k regex:smush_loss_grad --launch-count 2 -o $OUT/${TAG}_smush_adj python scripts/smush_one.py > $OUT/${TAG}_ncu_smush.log 2>&1
has k2 && python scripts/k2_one.py > /dev/null 2>&1 && $FULL -k regex:loss_grad_kernel --launch-count 2 -o $OUT/${TAG}_loss_grad python scripts/k2_one.py > $OUT/${TAG}_ncu_k2.log 2>&1
has weyl && python scripts/weyl_traj_one.py > /dev/null 2>&1 && $FULL -k "regex:weyl_kernel|traj" --launch-count 2 -o $OUT/${TAG}_weyl_traj python scripts/weyl_traj_one.py > $OUT/${TAG}_ncu_weyl.log 2>&1
# summarise on the box (gpurun copies back at most 64 MiB: the reports themselves stay here, except the K5 k = 3 one)
for r in $OUT/${TAG}_*.ncu-rep; do
  [ -f "$r" ] || continue
  python scripts/make_profile_summary.py $r ${r%.ncu-rep}_summary.txt > /dev/null 2>&1
done
for r in $OUT/${TAG}_*.ncu-rep; do
  case "$r" in *lbfgs_k3*) ;; *) rm -f "$r" ;; esac
done
ls -la $OUT | grep $TAG
