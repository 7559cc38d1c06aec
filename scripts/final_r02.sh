mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final/gputests.log 2>&1; tail -4 gpurun_out/final/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; tail -5 gpurun_out/final/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final/bench_n1_reference_arm.json 2> gpurun_out/final/ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench.err; tail -c 300 gpurun_out/final/bench.err
bash scripts/launches_r02.sh > gpurun_out/final/launches.log 2>&1; tail -3 gpurun_out/final/launches.log
timeout 900 python scripts/bench_create_index.py > gpurun_out/final/create_index_10Mx768.jsonl 2> gpurun_out/final/ci.err; cat gpurun_out/final/create_index_10Mx768.jsonl | cut -c1-260
du -sh gpurun_out
