#!/usr/bin/env bash
# Runs ON THE GPU BOX (via gpurun, one GPU): the ncu evidence of this round, following
# /opt/skills/guides/B200_PROFILING.md — every ncu command only after the same command exited 0 without ncu.
# Outputs land in gpurun_out/; tools/ncu_extract.py turns them into the committed summaries under profiles/.
set -u
R=${ROUND:-r02}
O=gpurun_out
mkdir -p $O

# 1. launch list of the bench command (device time of every launch; shares, not absolutes)
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --extras C2,C3 --tto C2 --tto-se C2,C3"
$BENCH > $O/${R}_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_bench.csv $BENCH > $O/${R}_launches_bench.log 2>&1
echo "launch list rc=$?"

# 2. full metric set, steady-state launch (3rd of 3) of every loop kernel:
#    general persistent kernel at C4 (bench workload) and C3 (roofline config), steepest-edge variant at C3,
#    shared-memory-resident kernel at C2, tiny kernel (Klee-Minty 20), sharded loop (2 ranks emulated on one device)
# The reports are ~50 MB each (gpurun brings back 64 MiB at most): the raw page and the hottest source lines are
# exported to CSV on the box and the .ncu-rep is dropped.
export_rep() { # $1 = report stem (without .ncu-rep)
	ncu -i $1.ncu-rep --page raw --csv > $1_raw.csv 2> /dev/null
	ncu -i $1.ncu-rep --page source --csv > $1_source_full.csv 2> /dev/null
	python tools/ncu_hot_lines.py $1_source_full.csv 60 > $1_source_top.txt 2> /dev/null
	rm -f $1.ncu-rep $1_source_full.csv
}
prof() { # $1 tag  $2 kernel regex  $3.. target arguments
	local tag=$1 rx=$2; shift 2
	local T="python tools/prof_target.py $*"
	$T > $O/${R}_prof_plain_$tag.log 2>&1 &&
	ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o $O/${R}_prof_$tag $T > $O/${R}_prof_ncu_$tag.log 2>&1
	echo "$tag rc=$?"
	export_rep $O/${R}_prof_$tag
}
prof persistent_c4 "simplex_persistent" --lp 32768x65536 --pivots 8 --launches 3
prof persistent_c3 "simplex_persistent" --lp 8192x16384 --pivots 24 --launches 3
prof steepest_c3 "simplex_persistent" --lp 8192x16384 --pivots 16 --launches 3 --rule 1
prof resident_c2 "simplex_resident" --lp 1024x2048 --pivots 200 --launches 3
prof tiny_km20 "simplex_tiny" --km 20 --pivots 4000 --launches 3
prof sharded_emu2_c3 "simplex_persistent_sharded_emu" --lp 8192x16384 --pivots 16 --launches 3 --emu 2

# 3. the two streaming phases as stand-alone kernels (one launch per phase mode), full metric set, C3
T="python tools/prof_target.py --lp 8192x16384 --pivots 6 --launches 1 --phases"
$T > $O/${R}_prof_plain_phases.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_price|k_update_ftran" -s 6 -c 4 -f -o $O/${R}_prof_phases_c3 $T > $O/${R}_prof_ncu_phases.log 2>&1
echo "phases rc=$?"
export_rep $O/${R}_prof_phases_c3
ls -la $O | grep ${R}_
