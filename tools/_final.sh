ncu --set full --clock-control none --import-source on -k regex:ertb_render_pool_kernel -s 2 -c 1 -f -o gpurun_out/r02x_c2 python bench.py --steps 2 --warmup 1 > gpurun_out/r02x_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02x_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02x_b.log 2>&1
python tools/bench_spectral.py > gpurun_out/r02x_spectral.json 2>&1
tail -2 gpurun_out/r02x_ncu.log
