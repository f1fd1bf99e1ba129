set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02h_ref.json 2> gpurun_out/r02h_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'comb_|logo_|fill_pairs|scan_|erase' -c 40 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/r02h_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:comb_ws -s 1 -c 1 -o gpurun_out/r02h_comb_ws python tools/prof_step.py --frames 1800 --steps 2 --comb-only > gpurun_out/r02h_ncu_ws.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logo_scores -s 1 -c 1 -o gpurun_out/r02h_logo_scores python tools/prof_step.py --frames 1800 --steps 2 > gpurun_out/r02h_ncu_logo.log 2>&1
AMTK_COMB_MMA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:comb_mma -s 1 -c 1 -o gpurun_out/r02h_comb_mma1 python tools/prof_step.py --frames 1800 --steps 2 --comb-only > gpurun_out/r02h_ncu_mma.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scan_frames_bit_exact or analyze_frames or comb_bit_exact_small or fused_and_host or comb_thresholds or logoscan_accumulate or erase_logo or weave or unaligned or comb_16bit or empty_single" > gpurun_out/r02h_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02h_memcheck.txt
AMTK_COMB_MMA=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "comb_bit_exact_small or comb_thresholds" > gpurun_out/r02h_memcheck_mma.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02h_memcheck_mma.txt
tail -3 gpurun_out/r02h_memcheck.txt gpurun_out/r02h_memcheck_mma.txt
