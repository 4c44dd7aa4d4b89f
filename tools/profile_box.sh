#!/usr/bin/env bash
# Runs ON THE GPU BOX (via gpurun, one GPU): the ncu evidence of this round, following
# /opt/skills/guides/B200_PROFILING.md — every ncu command only after the same command exited 0 without ncu.
# Outputs land in gpurun_out/; tools/ncu_extract.py turns them into the committed summaries under profiles/.
set -u
R=${ROUND:-r01}
O=gpurun_out
mkdir -p $O

# 1. launch list of the bench command (device time of every launch; shares, not absolutes)
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --extras C3 --tto C2"
$BENCH > $O/${R}_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_bench.csv $BENCH > $O/${R}_launches_bench.log 2>&1
echo "launch list rc=$?"

# 2. the persistent kernel, full metric set, steady-state launch (3rd of 3): C4 (the bench workload) and C3
for CFG in "32768x65536 8 c4" "8192x16384 24 c3"; do
	set -- $CFG
	T="python tools/prof_target.py --lp $1 --pivots $2 --launches 3"
	$T > $O/${R}_prof_plain_$3.log 2>&1 &&
	ncu --set full --clock-control none --import-source on -k regex:simplex_persistent -s 2 -c 1 -f -o $O/${R}_prof_persistent_$3 $T > $O/${R}_prof_ncu_$3.log 2>&1
	echo "persistent $3 rc=$?"
done

# 3. the two streaming phases as stand-alone kernels (one launch per phase mode), full metric set, C3
T="python tools/prof_target.py --lp 8192x16384 --pivots 6 --launches 1 --phases"
$T > $O/${R}_prof_plain_phases.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_price|k_update_ftran" -s 6 -c 4 -f -o $O/${R}_prof_phases_c3 $T > $O/${R}_prof_ncu_phases.log 2>&1
echo "phases rc=$?"
ls -la $O | grep ${R}_
