# Regenerates the round-2 evidence under profiles/ on a B200 box (run through gpurun; outputs land in gpurun_out/, copy the
# ones listed in profiles/README.md into profiles/).
R=r2
set -x
for w in c1 c2 c3 c4 c5; do timeout 400 python bench.py --workload $w > gpurun_out/${R}_bench_$w.json 2> gpurun_out/${R}_bench_$w.err; done
timeout 400 python bench.py --points-dtype f32 --no-cpu-baseline > gpurun_out/${R}_bench_c2_f32rows.json 2> gpurun_out/${R}_bench_c2_f32rows.err
timeout 400 python bench.py --workload c4 --algo tiled --no-cpu-baseline > gpurun_out/${R}_bench_c4_tiled.json 2> gpurun_out/${R}_bench_c4_tiled.err
timeout 400 python bench.py --workload c5 --dense-augment --no-cpu-baseline > gpurun_out/${R}_bench_c5_dense.json 2> gpurun_out/${R}_bench_c5_dense.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err
timeout 200 python tools/time_fused_aug.py 4 > gpurun_out/${R}_fused_aug_classes.txt 2>&1
timeout 200 python tools/time_fused_aug.py 4 "" 5 empty > gpurun_out/${R}_fused_aug_empty_input.txt 2>&1
timeout 200 python tools/time_fused_aug.py 4 rot+scale 5 c4 > gpurun_out/${R}_fused_aug_dense_scene.txt 2>&1
timeout 300 python tools/time_batch_sizes.py > gpurun_out/${R}_batch_sizes.txt 2>&1
timeout 300 python tools/feed_vit_consumer.py 16 20 > gpurun_out/${R}_feed_vit_consumer.txt 2>&1
timeout 300 python tools/time_av2_dir.py > gpurun_out/${R}_av2_directory_from_disk.txt 2>&1
timeout 200 python tools/time_latency.py > gpurun_out/${R}_call_latency.txt 2>&1
timeout 200 python tools/time_strict.py > gpurun_out/${R}_strict_signature.txt 2>&1
# launch list of the default bench command (after the same command has run clean without ncu)
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_bench.log 2>&1
# full captures: default path at batch 32 on AV2 columns, fused augmentation (rotate; rotate + scale), map rasterizer
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bin_fixed|k_tile_store|k_overflow_red' -c 3 -o gpurun_out/${R}_prof_tiled -f python tools/prof_one.py 0 32 c2 1 av2 > gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tile_aug|k_rebin_aug' -c 2 -o gpurun_out/${R}_prof_tile_aug -f python tools/time_fused_aug.py 4 rotate 1 > gpurun_out/${R}_ncu_aug.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tile_aug' -c 1 -o gpurun_out/${R}_prof_tile_aug_both -f python tools/time_fused_aug.py 4 rot+scale 1 > gpurun_out/${R}_ncu_aug2.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.txt 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.txt 2>&1
echo done
