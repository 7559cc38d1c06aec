mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final/gputests.log 2>&1; tail -3 gpurun_out/final/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; tail -2 gpurun_out/final/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final/bench_n1_reference_arm.json 2> gpurun_out/final/ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench.err; tail -c 300 gpurun_out/final/bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --sweep 2,4,8,32,64,96,128,256 --secondary none --no-cpu-baseline --no-parity > gpurun_out/final/bench_n1_sweep.json 2> gpurun_out/final/sweep.err
