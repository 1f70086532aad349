mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02v_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.log 2>&1
python bench.py > gpurun_out/r02v_bench_n1.json 2> gpurun_out/r02v_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02v_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02v_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ertb_render_pool_kernel -s 2 -c 1 -f -o gpurun_out/r02v_c2 python bench.py --steps 2 --warmup 1 > gpurun_out/r02v_ncu.log 2>&1
tail -3 gpurun_out/r02v_pytest_gpu.log; tail -1 gpurun_out/r02v_smoke.log; cut -c1-300 gpurun_out/r02v_bench_n1.json
