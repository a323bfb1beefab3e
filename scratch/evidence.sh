#!/bin/bash
# One evidence round on the GPU box (run through gpurun): tests, smoke, bench (both arms), stream, launch list, full ncu capture.
tag=${1:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --copy-ceiling > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 1500 gpurun_out/${tag}_bench.json; echo
timeout 300 python bench.py --stream 10000 --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench_stream.json 2> gpurun_out/${tag}_bench_stream.err
python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_stream.json')); print('stream', d.get('stream'))"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; tail -c 600 gpurun_out/${tag}_bench_ref.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --frames 32 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1; tail -2 gpurun_out/${tag}_ncu_launch.log | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none -o gpurun_out/${tag}_prof -f python scratch/prof_run.py 3840 2160 37 > gpurun_out/${tag}_ncu_full.log 2>&1; tail -2 gpurun_out/${tag}_ncu_full.log
