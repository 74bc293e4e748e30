mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:'conv_x3_kernel|conv_first_x3|conv_last_rows_x3|adain_nhwc' -s 63 -c 21 -f -o gpurun_out/prof_x3_r03m python tools/layer_report.py --iters 1 --batch 32 --precision fp16x3 > gpurun_out/ncu_x3_r03m.log 2>&1
echo rc=$?
python tools/ncu_summary.py gpurun_out/prof_x3_r03m.ncu-rep > gpurun_out/ncu_x3_summary_r03m.txt 2>&1
cut -c1-190 gpurun_out/ncu_x3_summary_r03m.txt
rm -f gpurun_out/prof_x3_r03m.ncu-rep
