# The one-call evidence set of a round (run on the GPU box):  bash tools/evidence_run.sh <tag>
T=${1:-r02w}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'comb_|logo_|fill_pairs|scan_|erase' -c 40 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/${T}_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:comb_ws -s 1 -c 1 -o gpurun_out/${T}_comb_ws python tools/prof_step.py --frames 1800 --steps 2 --comb-only > gpurun_out/${T}_ncu_ws.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scan_frames_bit_exact or analyze_frames or comb_bit_exact_small or fused_and_host or comb_thresholds or logoscan_accumulate or erase_logo or weave or unaligned or comb_16bit or empty_single" > gpurun_out/${T}_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.txt
tail -n 3 gpurun_out/${T}_pytest_gpu.txt gpurun_out/${T}_memcheck.txt
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['parity']['scores_bitexact'],d['parity']['counts_equal']);print(json.dumps(d['secondary'].get('single_frame_1440')))"
