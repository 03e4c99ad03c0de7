set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 300 gpurun_out/bench_r2d.err; head -c 200 gpurun_out/bench_r2d.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err; head -c 400 gpurun_out/bench_ref_r2.json; echo
python bench.py --steps 1 --warmup 1 --frames 1024 --no-cpu-baseline --no-e2e --no-extra --lanes 1 > gpurun_out/plain_r2d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 1 --warmup 1 --frames 1024 --no-cpu-baseline --no-e2e --no-extra --lanes 1 > gpurun_out/ncu_r2d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lst_ncc_tc|lst_preprocess_win|lst_epilogue" -s 9 -c 9 -o gpurun_out/prof_r2d -f python bench.py --steps 1 --warmup 1 --frames 1024 --no-cpu-baseline --no-e2e --no-extra --lanes 1 > gpurun_out/ncu_r2d_full.log 2>&1
ls -la gpurun_out/prof_r2d.ncu-rep gpurun_out/r2d_launches.csv
