mkdir -p gpurun_out
python tools/layer_report.py --batch 1024 --size 96 > gpurun_out/layers_96_r02p.log 2>&1; cat gpurun_out/layers_96_r02p.log
