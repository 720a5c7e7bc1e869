#!/bin/bash
# copy the evidence of a GPU visit (gpurun_out/*_<tag>.*) into profiles/ under the round's names, condense the ncu captures
# usage: bash scripts/collect_profiles.sh <visit tag> <round tag>      e.g.  r02 r02
V=${1:-r02}; R=${2:-r02}; O=gpurun_out; P=profiles
cp $O/bench_${V}.json $P/${R}_bench_n1.json
cp $O/bench_reference_${V}.json $P/${R}_bench_reference_n1.json
for p in mixed f32 f64; do cp $O/launches_${p}_${V}.csv $P/${R}_launches_${p}.csv; done
cp $O/sweep_${V}.jsonl $P/${R}_sweep_configs.jsonl
cp $O/reduce_timeline_${V}.txt $P/${R}_reduce_timeline.txt
cp $O/closed_loop_${V}.txt $P/${R}_closed_loop_probe.txt
cp $O/host_timing_${V}.txt $P/${R}_host_timing.txt
cp $O/pipes_${V}.txt $P/${R}_pipes.txt
cp $O/gen_${V}.txt $P/${R}_gen.txt
cp $O/gpu_${V}.txt $P/${R}_gpu_box.txt
cp $O/pytest_gpu_${V}.log $P/${R}_pytest_gpu.log
cp $O/sanitizer_memcheck_${V}.log $P/${R}_sanitizer_memcheck.log
cp $O/sanitizer_racecheck_${V}.log $P/${R}_sanitizer_racecheck.log
python profiles/ncu_summarize.py $P/${R}_ncu_summary.json \
  "rollout_lean_sm_kernel<DIFF_DRIVE,SCREEN,nogrid> (precision mixed; K 65536 T 64)=$O/prof_rollout_mixed_${V}.ncu-rep" \
  "reduce_screen_kernel + finalizer block (precision mixed; K 65536 T 64)=$O/prof_reduce_mixed_${V}.ncu-rep" \
  "rollout_lean_kernel<DIFF_DRIVE,SCREEN,nogrid,128> (precision mixed; K 2097152 T 128)=$O/prof_rollout_c5_${V}.ncu-rep"
python profiles/sass_extract.py ${R}
ls -la $P | grep ${R}_ | wc -l
