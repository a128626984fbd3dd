#!/bin/bash
# ONE-GPU session in the README's order, every step logged under gpurun_out/ (which gpurun copies back):
#     gpurun --timeout 2400 -- 'bash tools/first_gpu_session.sh'
# Stops at the first failing step: nothing is launched after a fault.  ncu passes run only after the same command exited 0 without ncu.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
step() { echo "=== $1"; }
step "1 pre-flight (compute-sanitizer, one process per path)"
timeout 1500 python tools/preflight.py > gpurun_out/r02_preflight.log 2>&1; rc=$?; tail -30 gpurun_out/r02_preflight.log
[ $rc -eq 0 ] || { echo "preflight failed (rc=$rc)"; exit 1; }
step "1b cluster mode finder probe (its own process; bench.py and the GPU tests repeat it and fall back to one CTA per mode if it fails)"
timeout 600 python tools/probe_cluster_finder.py > gpurun_out/r02_cluster_probe.json 2> gpurun_out/r02_cluster_probe.err; echo "probe rc=$?"; tail -c 600 gpurun_out/r02_cluster_probe.json
timeout 600 python tools/probe_cluster_finder.py --part many_mode > gpurun_out/r02_many_mode_probe.json 2> gpurun_out/r02_many_mode_probe.err; echo "many-mode probe rc=$?"; tail -c 300 gpurun_out/r02_many_mode_probe.json
step "2 pytest -m gpu"
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gpu_tests.log 2>&1; rc=$?; tail -15 gpurun_out/r02_gpu_tests.log
[ $rc -eq 0 ] || { echo "gpu tests failed (rc=$rc)"; exit 2; }
step "3 smoke + bench (c3 default), reference arm"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1 || { tail -20 gpurun_out/r02_smoke.log; exit 3; }
timeout 900 python bench.py > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; rc=$?; tail -c 3000 gpurun_out/r02_bench_c3.json
[ $rc -eq 0 ] || { tail -20 gpurun_out/r02_bench_c3.err; exit 4; }
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>&1
step "4 other workloads as main lines"
for w in c1 c2 c4 c5; do
  timeout 900 python bench.py --workload $w --no-extra --steps 5 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err || { echo "bench $w failed"; tail -5 gpurun_out/r02_bench_$w.err; }
done
step "5 ncu: launch list of the default command, then one full capture of the chain kernel"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --no-modes > gpurun_out/r02_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:logistic_staged_kernel -s 3 -c 1 -o gpurun_out/r02_c3_chain \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-modes --iters-per-step 100 > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out | tail -20
echo "session complete"
