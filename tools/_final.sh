mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02u_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02u_smoke.log 2>&1
python bench.py > gpurun_out/r02u_bench_n1.json 2> gpurun_out/r02u_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02u_bench_ref.json 2> gpurun_out/r02u_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02u_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02u_b.log 2>&1
tail -3 gpurun_out/r02u_pytest_gpu.log; tail -2 gpurun_out/r02u_smoke.log; cut -c1-600 gpurun_out/r02u_bench_n1.json; cut -c1-400 gpurun_out/r02u_bench_ref.json
