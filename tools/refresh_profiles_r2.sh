#!/bin/bash
# Round-2 ncu evidence (run on the GPU box through gpurun; outputs under gpurun_out/, summarised into
# profiles/ by tools/ncu_summary.py here).  Every command first runs WITHOUT ncu and must exit 0.
set -x
B5="python bench.py --total-spectra 256 --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-superposition --no-smooth-saturation --no-small-spectra --no-config3 --no-one-call"
B3="python bench.py --workload config3 --total-spectra 512 --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-superposition --no-smooth-saturation --no-small-spectra --no-config3 --no-one-call"
export MDB_CHUNK_SPECTRA=64
$B5 > gpurun_out/p5.json 2> gpurun_out/p5.err || { tail -5 gpurun_out/p5.err; exit 1; }
# launch list of the same command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 200 --csv --log-file gpurun_out/launches_r2.csv $B5 > gpurun_out/ncu_launches_r2.log 2>&1
full() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s $skip -c 1 -o gpurun_out/prof_${name}_r2 -f "$@" > gpurun_out/ncu_${name}_r2.log 2>&1
  tail -1 gpurun_out/ncu_${name}_r2.log
}
full fit_iter fit_iter_kernel 25 $B5
full mse_superposition "superposition_kernel<.int.2," 2 $B5
full mse_partials mse_partials_kernel 2 $B5
full detect detect_kernel 2 $B5
full smooth smooth_lanes_kernel 2 $B5
full select select_kernel 2 $B5
export MDB_CHUNK_SPECTRA=256
$B3 > gpurun_out/p3.json 2> gpurun_out/p3.err || { tail -5 gpurun_out/p3.err; exit 1; }
full fit_iter_config3 fit_iter_kernel 25 $B3
unset MDB_CHUNK_SPECTRA
python tools/run_sup_once.py > gpurun_out/sup_once.log 2>&1 && full superposition_vec "superposition_kernel<.int.0," 1 python tools/run_sup_once.py
python tools/run_blood_once.py > gpurun_out/blood_once.log 2>&1 && full smooth_split smooth_split_kernel 1 python tools/run_blood_once.py
MDB_SMOOTH_SPLIT=0 full smooth_stream smooth_stream_kernel 1 python tools/run_blood_once.py
full fit_wide2 fit_wide2_superpose_kernel 12 python tools/run_blood_once.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_blood_r2.csv python tools/run_blood_once.py > /dev/null 2>&1
python tools/run_sim_once.py > gpurun_out/sim_once.log 2>&1 && full small_fused small_fused_kernel 1 python tools/run_sim_once.py
ls -la gpurun_out/*_r2.ncu-rep
