# Round-2 ncu launch lists of bench.py (durations + DRAM bytes per launch) at batch 1 / 16 / 1024, one GPU.
# Only this library's kernels are listed (-k regex:wb::): the generator's torch kernels are skipped by name.
mkdir -p gpurun_out/launches
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
run() {  # batch skip count
  CMD="python bench.py --batch $1 --steps 2 --warmup 3 --secondary none --no-cpu-baseline --no-parity"
  $CMD > gpurun_out/launches/plain_b$1.json 2> gpurun_out/launches/plain_b$1.err && \
  ncu $M -k regex:$4 -s $2 -c $3 --log-file gpurun_out/launches/launches_bench_n1_batch$1.csv $CMD > gpurun_out/launches/ncu_b$1.log 2>&1
}
if [ "$1" = "b1" ]; then run 1 60 30 scan_topk; else
run 1 60 30 scan_topk
run 16 1900 400 'wb::'
run 1024 380 200 'wb::'
fi
ls -la gpurun_out/launches
