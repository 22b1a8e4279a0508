for B in 10 16 25 32; do python bench.py --batch $B --steps 8 --warmup 3 --skip pixel_sum,c3,c4,c5,e2e,cpu > gpurun_out/r5_batch_$B.json 2> gpurun_out/r5_batch_$B.err; done
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r5_step_counts.csv python bench.py --steps 1 --warmup 0 --skip pixel_sum,c3,c4,c5,e2e,cpu > gpurun_out/r5_ncu.log 2>&1
tail -c 300 gpurun_out/r5_batch_32.json
