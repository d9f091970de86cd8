# Regenerates the evidence under profiles/ on a B200 box (run through gpurun; outputs land in gpurun_out/).
set -x
for w in c2 c3 c4 c5; do timeout 400 python bench.py --workload $w > gpurun_out/r1_bench_$w.json 2> gpurun_out/r1_bench_$w.err; done
timeout 400 python bench.py --workload c5 --dense-augment --no-cpu-baseline > gpurun_out/r1_bench_c5_dense.json 2> gpurun_out/r1_bench_c5_dense.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference_arm.json 2> gpurun_out/r1_bench_reference_arm.err
timeout 200 python tools/time_fused_aug.py 4 > gpurun_out/r1_fused_aug_classes.txt 2>&1
timeout 200 python tools/time_fused_aug.py 4 rot+scale 5 c4 > gpurun_out/r1_fused_aug_dense_scene.txt 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bin_fixed|k_tile_store|k_overflow_red' -c 3 -o gpurun_out/r1_prof_tiled -f python tools/prof_one.py 0 32 c2 1 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tile_aug|k_rebin_aug' -c 2 -o gpurun_out/r1_prof_tile_aug -f python tools/time_fused_aug.py 4 rotate 1 > gpurun_out/ncu_aug.log 2>&1
timeout 200 python tools/time_fused_aug.py 4 "" 5 empty > gpurun_out/r1_fused_aug_empty_input.txt 2>&1
timeout 200 python tools/time_strict.py > gpurun_out/r1_strict_signature.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.txt 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.txt 2>&1
echo done
