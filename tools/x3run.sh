mkdir -p gpurun_out
bash tools/gpu_r02.sh r02v san ncu_all ncu_list
