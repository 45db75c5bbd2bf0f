# gpurun_out/<tag> artefacts of scripts/gpu_final.sh -> profiles/<tag>_* (tracked summaries).  usage: bash scripts/make_profile_summaries.sh <tag>
T=${1:?tag}; P=profiles; G=gpurun_out; K='dcb_step_kernel_704ILb1ELb0ELb0'; LIB=deepcomp_b200/libdeepcomp_b200.so
cp $G/bench_$T.json $P/${T}_bench.json
cp $G/bench_${T}_driver.json $P/${T}_bench_driver_flags.json
grep '^{' $G/bench_ref_$T.json | tail -1 > $P/${T}_bench_reference.json
cp $G/launches_$T.csv $P/${T}_launches.csv
[ -f $G/timeline_$T.txt ] && cp $G/timeline_$T.txt $P/${T}_timeline.txt
(tail -n 4 $G/racecheck_$T.log; echo "--- wide kernel (DCB_FORCE_WIDE=1)"; tail -n 4 $G/racecheck_wide_$T.log) > $P/${T}_racecheck.log
tail -n 3 $G/pytest_$T.log > $P/${T}_pytest_gpu.log
python scripts/ncu_summary.py $G/prof_$T.ncu-rep $P/${T}_step_kernel_ncu_summary.csv > /dev/null
python scripts/ncu_summary.py $G/prof_f20_$T.ncu-rep $P/${T}_step_kernel_f20_ncu_summary.csv > /dev/null
python scripts/ncu_summary.py $G/prof_wide_$T.ncu-rep $P/${T}_wide_kernel_ncu_summary.csv > /dev/null
python scripts/ncu_regions.py $G/prof_$T.ncu-rep $K > $P/${T}_step_kernel_regions.txt
(TOP=30 python scripts/ncu_roles.py $G/prof_$T.ncu-rep $LIB $K; python scripts/ncu_stalls.py $G/prof_$T.ncu-rep $LIB $K) > $P/${T}_step_kernel_roles.txt
python scripts/ncu_hotlines.py $G/prof_wide_$T.ncu-rep $LIB 'dcb_wide_kernelILb0ELb0' 40 > $P/${T}_wide_kernel_hotlines.txt
python scripts/make_configs_table.py $T
