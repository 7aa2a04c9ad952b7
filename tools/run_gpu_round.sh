# One GPU call that produces everything the round's profiles/ are built from (run under gpurun from the repo root).
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
timeout 600 python tools/gpu_joint.py > gpurun_out/joint.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --missions 2368 --no-cpu-baseline --jacobi-missions 0 --joint-missions 0 > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pdip1 -s 2 -c 1 -o gpurun_out/pdip1_full -f python bench.py --steps 1 --warmup 3 --missions 2368 --no-cpu-baseline --jacobi-missions 0 --joint-missions 0 > gpurun_out/b_ncu2.log 2>&1
JOINT_CASES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pdip_kernel -s 2 -c 1 -o gpurun_out/pdip_joint_b16 -f python tools/gpu_joint.py > gpurun_out/b_ncu3.log 2>&1
ls -la gpurun_out
