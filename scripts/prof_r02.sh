set -x
mkdir -p gpurun_out/prof
P="python scripts/prof_r02.py"
N="ncu --set full --clock-control none --import-source on"
$P flat1 > gpurun_out/prof/plain_flat1.log 2>&1 && $N -k regex:scan_topk -s 3 -c 1 -o gpurun_out/prof/flat1 $P flat1 > gpurun_out/prof/ncu_flat1.log 2>&1
$P flat16 > gpurun_out/prof/plain_flat16.log 2>&1 && $N -k 'regex:gemm_topk_kernel|rescore_candidates|compact_topk' -s 19 -c 3 -o gpurun_out/prof/flat16 $P flat16 > gpurun_out/prof/ncu_flat16.log 2>&1
$P flat1024 > gpurun_out/prof/plain_flat1024.log 2>&1 && $N -k 'regex:rescore_candidates|compact_topk' -s 5 -c 2 -o gpurun_out/prof/flat1024_aux $P flat1024 > gpurun_out/prof/ncu_flat1024.log 2>&1
$P ivf_search > gpurun_out/prof/plain_ivf_search.log 2>&1 && $N -k 'regex:scan_topk_kernel|ivf_listmajor' -s 7 -c 5 -o gpurun_out/prof/ivf_search $P ivf_search > gpurun_out/prof/ncu_ivf_search.log 2>&1
$P ivf_train > gpurun_out/prof/plain_ivf_train.log 2>&1 && $N -k 'regex:filter2_topk|segment_sum|csr_scatter|csr_hist' -s 4 -c 4 -o gpurun_out/prof/ivf_train $P ivf_train > gpurun_out/prof/ncu_ivf_train.log 2>&1
$P ivf_add > gpurun_out/prof/plain_ivf_add.log 2>&1 && $N -k 'regex:gemm2_topk_kernel' -s 2 -c 1 -o gpurun_out/prof/ivf_add $P ivf_add > gpurun_out/prof/ncu_ivf_add.log 2>&1
cat gpurun_out/prof/plain_*.log | grep -v "^+" | tail -20
ls -la gpurun_out/prof
