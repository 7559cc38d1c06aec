# Round-2 closing run on ONE B200 (`gpurun --timeout 1500 -- 'bash scripts/run_r02_final.sh'`): the new fused-coarse test
# first (own process, own timeout: a protocol bug there must not take the rest of the run with it), then the whole GPU
# suite, the IVF call-latency A/B, the driver's bench commands, and one ncu capture of the fused IVF kernel.
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 420 python -m pytest tests/test_ivf_gpu.py -m gpu -q -k fused_coarse > $O/test_fused_coarse.log 2>&1
FUSED_RC=$?
echo "fused_coarse rc=$FUSED_RC" | tee $O/status.txt
tail -5 $O/test_fused_coarse.log
DESEL=""
if [ $FUSED_RC -ne 0 ]; then export WB_IVF_FUSE_COARSE=0; DESEL="--deselect tests/test_ivf_gpu.py::test_ivf_fused_coarse_matches_two_launch_path"; fi
timeout 900 python -m pytest tests -m gpu -q --durations=12 $DESEL > $O/test_gpu.log 2>&1
echo "gpu suite rc=$?" | tee -a $O/status.txt
tail -25 $O/test_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/status.txt
if [ $FUSED_RC -eq 0 ]; then
  timeout 300 python scripts/call_latency.py --ivf-only > $O/ivf_call_latency.jsonl 2> $O/ivf_call_latency.err
  echo "ivf latency rc=$?" | tee -a $O/status.txt
  cat $O/ivf_call_latency.jsonl
fi
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" | tee -a $O/status.txt
cut -c1-600 $O/bench_n1.json
if [ $FUSED_RC -eq 0 ]; then
  P="python scripts/prof_r02.py ivf_fused"
  timeout 200 $P > $O/plain_ivf_fused.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none -k regex:scan_topk -s 5 -c 1 -o $O/ivf_fused $P > $O/ncu_ivf_fused.log 2>&1
  if [ -f $O/ivf_fused.ncu-rep ]; then
    ncu -i $O/ivf_fused.ncu-rep --page raw --csv > $O/ncu_full_ivf_fused_raw.csv 2>/dev/null
    ncu -i $O/ivf_fused.ncu-rep --page details > $O/ncu_full_ivf_fused_details.txt 2>/dev/null
    rm -f $O/ivf_fused.ncu-rep
  fi
  cat $O/plain_ivf_fused.log
fi
cat $O/status.txt
