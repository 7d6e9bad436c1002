set -x
B5="python bench.py --total-spectra 256 --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-superposition --no-smooth-saturation --no-small-spectra --no-config3"
B3="python bench.py --workload config3 --total-spectra 512 --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-superposition --no-smooth-saturation --no-small-spectra --no-config3"
export MDB_CHUNK_SPECTRA=64
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:superposition_kernel<2" -s 2 -c 1 -o gpurun_out/prof_mse_superposition_r2 -f $B5 > gpurun_out/ncu_mse_superposition_r2.log 2>&1; tail -n 2 gpurun_out/ncu_mse_superposition_r2.log
export MDB_CHUNK_SPECTRA=256
ncu --set full --clock-control none --import-source on -k regex:fit_iter_kernel -s 25 -c 1 -o gpurun_out/prof_fit_iter_config3_warp_r2 -f $B3 > gpurun_out/ncu_fit_iter_config3_warp_r2.log 2>&1; tail -n 2 gpurun_out/ncu_fit_iter_config3_warp_r2.log
unset MDB_CHUNK_SPECTRA
ncu --set full --clock-control none --import-source on -k regex:superposition_kernel -s 3 -c 1 -o gpurun_out/prof_superposition_vec_r2 -f python tools/run_sup_once.py > gpurun_out/ncu_superposition_vec_r2.log 2>&1; tail -n 2 gpurun_out/ncu_superposition_vec_r2.log
