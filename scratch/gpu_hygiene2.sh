#!/bin/bash
# evidence for the kernels added late in the round: compute-sanitizer over the all-CTA tail, the team-per-column kernel, the
# peer evaluation; launch list of the final c1 iteration
mkdir -p gpurun_out
O=gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $O/sanitizer2_memcheck.log python -m pytest tests/test_gpu_graph.py tests/test_gpu_midcols.py tests/test_gpu_peer.py tests/test_gpu_configs.py -m gpu -q -x > $O/sanitizer2_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file $O/sanitizer2_racecheck.log python -m pytest "tests/test_gpu_graph.py::test_all_cta_tail_equals_the_last_cta_tail" "tests/test_gpu_midcols.py::test_team_per_column_kernel_equals_the_in_kernel_path" "tests/test_gpu_midcols.py::test_long_columns_cta_per_column_kernel_against_the_oracle" -m gpu -q -x > $O/sanitizer2_racecheck_pytest.log 2>&1; echo "racecheck rc=$?"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file $O/sanitizer2_synccheck.log python -m pytest tests/test_gpu_graph.py tests/test_gpu_midcols.py -m gpu -q -x > $O/sanitizer2_synccheck_pytest.log 2>&1; echo "synccheck rc=$?"
for f in memcheck racecheck synccheck; do echo "== $f"; tail -2 $O/sanitizer2_$f.log; tail -1 $O/sanitizer2_${f}_pytest.log; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:matching -c 18 --csv --log-file $O/c1_launches_final.csv python bench.py --workload c1 --steps 5 --warmup 3 --warm-start-iters 60 --no-cpu --no-e2e > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/c1_launches_final.csv")) if len(r)>5]
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if r[0]=="ID": hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    try: agg[d["Kernel Name"][:70]].append(float(d["Metric Value"].replace(",","")))
    except Exception: pass
for k,v in agg.items(): print("  %-72s n=%3d mean %.1f us"%(k,len(v),sum(v)/len(v)/1e3))
PY
