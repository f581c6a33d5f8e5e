#!/bin/bash
# evidence run: launch lists, full ncu captures of the kernels on timed paths (reduced to text on the box: the reports
# themselves exceed what gpurun copies back), compute-sanitizer logs
mkdir -p gpurun_out
O=gpurun_out
T=/tmp/ncu_tmp; mkdir -p $T
summarise() {  # $1 report (without extension), $2 output prefix
  ncu -i $T/$1.ncu-rep --page raw --csv > $T/$1_raw.csv 2>/dev/null
  python - "$T/$1_raw.csv" > $O/$2_metrics.txt <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "nvlrx__bytes.sum", "nvltx__bytes.sum")
for r in data:
    for h, u, v in zip(hdr, units, r):
        if h in keep or h.startswith("dram__bytes") or "nvl" in h:
            print(f"{h} [{u}] = {v}")
    print("--")
PY
  if [ -n "$3" ]; then
    ncu -i $T/$1.ncu-rep --page source --csv > $T/$1_src.csv 2>/dev/null
    python tools/ncu_source_summary.py $T/$1_src.csv 25 > $O/$2_source_summary.txt 2>&1
  fi
}
# 1. launch list of the C3 timed region
timeout 600 ncu --nvtx --nvtx-include "dualip_timed_region/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c3.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > $O/launches_c3.log 2>&1
# 2. full capture of the slab kernel at a late iterate of C3 (one launch)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matching_slab -s 150 -c 1 -o $T/prof_c3 python bench.py --steps 4 --warmup 3 --warm-start-iters 145 --no-cpu --no-e2e > $O/prof_c3.log 2>&1
summarise prof_c3 ncu_c3_slab_kernel src
# 3. c1: the slab kernel's warp-per-column path and the CTA-per-column kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:matching_slab -s 70 -c 1 -o $T/prof_c1_slab python bench.py --workload c1 --steps 4 --warmup 3 --warm-start-iters 65 --no-cpu > $O/prof_c1.log 2>&1
summarise prof_c1_slab ncu_c1_mid_path src
timeout 600 ncu --set full --clock-control none -k regex:matching_long_cta -s 70 -c 1 -o $T/prof_c1_long python bench.py --workload c1 --steps 4 --warmup 3 --warm-start-iters 65 --no-cpu >> $O/prof_c1.log 2>&1
summarise prof_c1_long ncu_c1_long_cta_kernel
# 4. launch lists of the small objectives (LP, fairness, operators) and the update kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lp_|agd_|fair_|epilogue" -c 40 --csv --log-file $O/launches_c5.csv python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fair_|epilogue|left_multiply|row_sums|gather_block|scatter_block|project_block" -c 60 --csv --log-file $O/launches_ops.csv python -m pytest tests/test_gpu_operators.py -m gpu -q -x > /dev/null 2>&1
# 5. compute-sanitizer
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $O/sanitizer_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_midcols.py tests/test_gpu_graph.py tests/test_gpu_peer.py tests/test_gpu_operators.py tests/test_gpu_properties.py -m gpu -q -x > $O/sanitizer_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file $O/sanitizer_racecheck.log python -m pytest "tests/test_gpu_parity.py::test_fixtures_generated_by_the_reference" tests/test_gpu_peer.py -m gpu -q -x > $O/sanitizer_racecheck_pytest.log 2>&1; echo "racecheck rc=$?"
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file $O/sanitizer_synccheck.log python -m pytest "tests/test_gpu_parity.py::test_fixtures_generated_by_the_reference" tests/test_gpu_midcols.py -m gpu -q -x > $O/sanitizer_synccheck_pytest.log 2>&1; echo "synccheck rc=$?"
for f in memcheck racecheck synccheck; do echo "== $f"; tail -3 $O/sanitizer_$f.log; tail -2 $O/sanitizer_${f}_pytest.log; done
rm -rf $T; du -sh $O
