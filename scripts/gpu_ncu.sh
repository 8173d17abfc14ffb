# one ncu --set full capture of the fused kernel (2 launches after 3 warm-up launches); $1 = output tag
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fused_augment -s 3 -c 2 -o gpurun_out/$1 python bench.py --steps 4 --warmup 3 --spin-s 0 --no-cpu-baseline > gpurun_out/$1.log 2>&1
tail -3 gpurun_out/$1.log
