# e2e upload-mode comparison on the GPU box; $1 = tag
mkdir -p gpurun_out
python neuralnet-tracker-traincode_b200/build.py > /dev/null
timeout 200 python -m pytest tests/test_gpu_standalone.py -m gpu -x -q -k upload 2>&1 | tail -3
for v in "" 1; do
  echo "ONE_BY_ONE=$v"
  env ${v:+B200AUG_UPLOAD_ONE_BY_ONE=1} timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print('bands', e['value'], 'whole', e['whole_frames']['value'], 'sync', e['whole_frames_synchronised_every_step']['value'], 'h2d', e['h2d_bytes_per_step'])"
done 2>&1 | tee gpurun_out/$1_e2e.log
