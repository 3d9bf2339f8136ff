#!/bin/bash
# round-end style validation + profiles (run under gpurun)
mkdir -p gpurun_out
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json; cat gpurun_out/bench_n1.json | cut -c1-1200; tail -3 gpurun_out/bench_n1.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.json; cut -c1-400 gpurun_out/bench_ref.json
echo "== bench frame workload"; timeout 900 python bench.py --workload frame --frame-blocks 64 --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_frame.json; cut -c1-900 gpurun_out/bench_frame.json
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --quick 2>&1 | tail -1
echo "== ncu full K1/K2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_compress -s 2 -c 1 -f -o gpurun_out/r1_k1 python bench.py --steps 1 --warmup 3 --quick 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_decompress -s 2 -c 1 -f -o gpurun_out/r1_k2 python bench.py --steps 1 --warmup 3 --quick 2>&1 | tail -1
