cd $GRAFT_REPO_ROOT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -s 140 -c 80 --log-file gpurun_out/s3_launches_bench_b36.csv python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/s3_ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_pm_kernel -s 6 -c 3 -o gpurun_out/s3_conv_pm python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/s3_ncu_b.log 2>&1
ncu -i gpurun_out/s3_conv_pm.ncu-rep --page raw --csv > gpurun_out/s3_conv_pm_raw.csv 2>/dev/null
python bench.py --steps 700 --no-extra --no-cpu-baseline > gpurun_out/s3_bench_n1_sustained.json 2> gpurun_out/s3_bench_sus.err
python bench.py --steps 100 > gpurun_out/s3_bench_n1.json 2> gpurun_out/s3_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/s3_bench_reference.json 2> gpurun_out/s3_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s3_smoke.log 2>&1; tail -2 gpurun_out/s3_smoke.log
ls -la gpurun_out | tail -12
