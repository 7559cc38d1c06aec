mkdir -p gpurun_out/prof
P="python scripts/prof_r02.py"
N="ncu --set full --clock-control none"
cap() {  # name, kernel regex, skip, count, mode
  $P $5 > gpurun_out/prof/plain_$1.log 2>&1 && $N -k "regex:$2" -s $3 -c $4 -o gpurun_out/prof/$1 $P $5 > gpurun_out/prof/ncu_$1.log 2>&1
  if [ -f gpurun_out/prof/$1.ncu-rep ]; then
    ncu -i gpurun_out/prof/$1.ncu-rep --page raw --csv > gpurun_out/prof/$1_raw.csv 2>/dev/null
    ncu -i gpurun_out/prof/$1.ncu-rep --page details > gpurun_out/prof/$1_details.txt 2>/dev/null
    rm -f gpurun_out/prof/$1.ncu-rep
  fi
}
cap flat1 'scan_topk' 3 1 flat1
cap flat16 'gemm_topk_kernel|rescore_candidates|compact_topk' 19 3 flat16
cap flat1024 'filter2_topk|rescore_candidates|compact_topk' 16 3 flat1024
ls -la gpurun_out/prof
