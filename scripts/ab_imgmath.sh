set -x
for v in cur img0 img2; do
  if [ $v = cur ]; then unset RFI_B200_LIB; else export RFI_B200_LIB=$PWD/rfi_toolbox_b200/_lib/ab/librfi_$v.so; fi
  python scripts/img_error.py > gpurun_out/r02_imgerr_$v.txt 2>&1
  python bench.py --steps 20 --warmup 5 > gpurun_out/r02_ab_$v.json 2> gpurun_out/r02_ab_$v.err
  python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/r02_ab_c5_$v.json 2> gpurun_out/r02_ab_c5_$v.err
done
unset RFI_B200_LIB
python -m pytest tests -m gpu -x -q > gpurun_out/r02_a_tests.log 2>&1; tail -5 gpurun_out/r02_a_tests.log
