mkdir -p gpurun_out
bash tools/gpu_r02.sh r03g bench tests smoke layers layers6 ops
python tools/layer_report.py --precision fp16x3 > gpurun_out/layers_x3_r03g.log 2>&1
python tools/layer_report.py --batch 1024 --size 96 > gpurun_out/layers_96_r03g.log 2>&1
python tools/fuzz_shapes.py 60 2 > gpurun_out/fuzz_r03g.log 2>&1; tail -1 gpurun_out/fuzz_r03g.log
