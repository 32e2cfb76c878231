mkdir -p gpurun_out
python tools/sanity_step.py fp32 > gpurun_out/sanity_fp32.txt 2>&1
python tools/sanity_step.py tf32 > gpurun_out/sanity_tf32.txt 2>&1
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python tools/sanity_step.py fp32 --no-oracle > gpurun_out/san_init_fp32.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanity_step.py fp32 --no-oracle > gpurun_out/san_race_fp32.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanity_step.py tf32 --no-oracle > gpurun_out/san_mem_tf32.txt 2>&1
timeout 300 python tools/step_profile.py --config shapes --res 64 --batch 32 --out gpurun_out/step_shapes64.md > gpurun_out/step_shapes64.log 2>&1
timeout 600 python tools/step_profile.py --config taichi --res 256 --batch 8 --out gpurun_out/step_taichi256.md > gpurun_out/step_taichi256.log 2>&1
cat gpurun_out/sanity_fp32.txt gpurun_out/sanity_tf32.txt; tail -5 gpurun_out/san_init_fp32.txt gpurun_out/san_race_fp32.txt gpurun_out/san_mem_tf32.txt
