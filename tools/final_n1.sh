#!/bin/bash
# round-end evidence on one GPU: GPU tests, smoke, the default bench line, the reference arm, then (each only after the same command
# exited 0 without ncu) the launch list of the timed V-cycles and one full capture of each smoother.  Outputs under gpurun_out/.
set -u
T=${1:-r02z}
O=gpurun_out
mkdir -p $O
( time timeout 300 python -m pytest tests -x -q -m gpu ) > $O/${T}_pytest.log 2>&1; tail -3 $O/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
timeout 420 python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "ref rc=$?"
P="bench.py --steps 2 --warmup 3 --profile --no-cpu --no-gap --no-small --no-valley --e2e-steps 0"
timeout 200 python $P > $O/${T}_profcmd.json 2> $O/${T}_profcmd.err; rc=$?; echo "profile command rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${T}_launches_amr.csv python $P > $O/${T}_ncu.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gsrb_twin --launch-count 1 -o $O/${T}_k_gsrb_twin python $P > $O/${T}_ncu_twin.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gsrb_patch --launch-count 1 -o $O/${T}_k_gsrb_patch python $P > $O/${T}_ncu_patch.log 2>&1
fi
ls -la $O | grep $T
