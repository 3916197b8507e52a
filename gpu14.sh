python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python benchmarks/micro_bench.py > gpurun_out/micro_r1.jsonl 2> gpurun_out/micro_r1.err; tail -3 gpurun_out/micro_r1.err; wc -l gpurun_out/micro_r1.jsonl
python benchmarks/solve_bench.py --robots 1 --poses 100 --landmarks 4 > gpurun_out/solve_manhattan100_r1.json 2> gpurun_out/solve1.err; tail -3 gpurun_out/solve1.err; cat gpurun_out/solve_manhattan100_r1.json
python benchmarks/solve_bench.py --robots 8 --poses 16 --landmarks 4 > gpurun_out/solve_mr8x16_g1_r1.json 2> gpurun_out/solve2.err; tail -3 gpurun_out/solve2.err; cat gpurun_out/solve_mr8x16_g1_r1.json
