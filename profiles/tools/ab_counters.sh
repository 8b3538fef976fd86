# Shared-memory wavefronts, bank conflicts, executed instructions and SM cycles of the feedback kernel for several builds
# (ncu metrics pass; numbers under the profiler are for comparison only).
# Usage (GPU box): bash profiles/tools/ab_counters.sh name1 name2 ...   (crazyflie_nmpc_b200/variants/libcfnmpc_<name>.so)
for v in "$@"; do
  CFNMPC_LIB=$PWD/crazyflie_nmpc_b200/variants/libcfnmpc_$v.so ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:'cf_rti_kernel' -s 3 -c 1 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | grep -E '^"[0-9]' | awk -F'","' -v v=$v '{gsub(/"/,"",$NF); printf "%s %s %s\n", v, $(NF-2), $NF}'
done
