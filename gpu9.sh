python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:nf_forward_kernel -s 1 -c 1 -o gpurun_out/prof_fwd_r1_c python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_fwd_c.log 2>&1
tail -2 gpurun_out/ncu_fwd_c.log | cut -c1-200
