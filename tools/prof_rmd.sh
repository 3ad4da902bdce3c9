mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rmd_batched|k_enum' -s 8 -c 4 -f -o gpurun_out/r01c_prof_rmd python bench.py --steps 3 --warmup 3 --pool 4 --no-cpu-baseline > gpurun_out/r01c_ncu_rmd.log 2>&1
tail -3 gpurun_out/r01c_ncu_rmd.log
