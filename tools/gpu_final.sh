#!/bin/bash
# Round-end evidence: bench (both arms), smoke, ncu launch list of the bench command, step profile.
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_arm.json 2> $out/bench_ref.err; echo "ref rc=$?"
PF_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $out/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
gzip -f $out/launches.csv
# one step out of the launch list: shares + DRAM bytes of the GEMM launches (bench.py's roofline.traffic);
# the JSON records the kernel-source hash it was measured at, bench.py refuses a stale one
python tools/ncu_launches_step.py $out/launches.csv.gz $out/ncu_launches_summary.txt $out/gemm_traffic.json "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python bench.py --steps 2 --warmup 1 --no-cpu" > /dev/null; echo "traffic rc=$?"
timeout 300 python tools/profile_step.py 64 > $out/step_profile_b64.txt 2>&1
head -12 $out/step_profile_b64.txt
cat $out/bench_n1.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d['clocks'])"
cat $out/bench_reference_arm.json | cut -c1-300
