set -x
cd $GRAFT_REPO_ROOT
timeout 300 python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/bench.err; echo bench rc=$?
timeout 300 python bench.py --impl reference > gpurun_out/r1_bench_reference_n1.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo launches rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate -c 1 -f -o gpurun_out/prof_c2full_v7 python tools/gpu_probe.py c2full > gpurun_out/ncu_full.log 2>&1; echo full rc=$?
timeout 300 python tools/gpu_probe.py peak c1big c1bigomega c2big c3 c5 c1one c2full c2fullcap1 c2heavy2 > gpurun_out/probe_final.log 2>&1
timeout 300 python tools/dropin_latency.py >> gpurun_out/probe_final.log 2>&1
timeout 600 python tools/gpu_probe.py c4full c5full > gpurun_out/probe_full.log 2>&1
tail -n 30 gpurun_out/probe_final.log; tail -n 8 gpurun_out/probe_full.log
