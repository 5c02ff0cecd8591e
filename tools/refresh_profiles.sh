# How the round-2 files under profiles/ are produced (each line is one gpurun call; gpurun merges at most 64 MiB back, so
# tools/gpu_job.sh prunes gpurun_out/ and tools/ncu_capture.sh exports CSVs instead of keeping the .ncu-rep files).
set -x
G=/usr/local/graft/bin/gpurun
$G --timeout 2700 -- 'tools/gpu_job.sh "timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_full.log 2>&1; timeout 500 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; timeout 300 python bench.py --impl reference > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_ref.err; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-north-star --no-blocks > gpurun_out/r2_ncu_launch.log 2>&1; tools/ncu_capture.sh c2full r2_c2full; tools/ncu_capture.sh c1bigomega r2_c1bigomega; tools/ncu_capture.sh c5 r2_c5; timeout 1500 python tests/parity_sweep.py 12000 gpurun_out/r2_parity_sweep.json 3000 32 > gpurun_out/r2_parity_sweep.log 2>&1; timeout 600 python tests/parity_sweep.py 0 gpurun_out/r2_parity_sweep_np3.json 0 0 4000 > gpurun_out/r2_parity_sweep_np3.log 2>&1"'
for n in 2 4 8; do
  $G --gpus $n --timeout 900 -- "tools/gpu_job.sh \"timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err\""
done
$G --timeout 1200 -- 'tools/gpu_job.sh "python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_driver.json; python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_driver_ref.json; python tools/gpu_probe.py peak c1big c1bigomega c3 c5 c2full > gpurun_out/r2_probe_a.log; python tools/overlap_probe.py; python tools/e2e_pools_probe.py"'
# then: cp the JSON / CSV files into profiles/, and
for c in c2full c1bigomega c5; do python tools/sass_mix.py gpurun_out/r2_${c}_source.csv 26 > profiles/r2_ncu_${c}_sass_mix.txt; cp gpurun_out/r2_${c}_raw.csv profiles/r2_ncu_${c}_raw.csv; done
