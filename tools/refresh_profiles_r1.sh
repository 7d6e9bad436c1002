set -x
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_final.log 2>&1; tail -3 gpurun_out/gpu_tests_final.log
python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -c 300 gpurun_out/bench_final2.err
B="python bench.py --spectra 256 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-superposition --no-smooth-saturation --no-small-spectra"
MDB_CHUNK_SPECTRA=256 $B > gpurun_out/b256.json 2> gpurun_out/b256.err || exit 1
MDB_CHUNK_SPECTRA=256 ncu --metrics gpu__time_duration.sum --clock-control none -s 1024 -c 170 --csv --log-file gpurun_out/launches_r1b.csv $B > gpurun_out/ncu_launches_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:superposition_kernel -s 258 -c 1 -o gpurun_out/prof_mse_fast_r1 -f $B > gpurun_out/ncu_mse_fast.log 2>&1; tail -2 gpurun_out/ncu_mse_fast.log
