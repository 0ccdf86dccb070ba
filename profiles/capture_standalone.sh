set -e
N="ncu --set full --clock-control none --import-source on -f"
timeout 100 $N -k regex:gauss_march --launch-skip 2 -c 1 -o gpurun_out/r1f_gauss python benchmarks/op_once.py gauss 1 > gpurun_out/r1f_gauss.log 2>&1 || echo gauss-fail
timeout 100 $N -k regex:median3x3_packed --launch-skip 2 -c 1 -o gpurun_out/r1f_median3 python benchmarks/op_once.py median2d 1 > gpurun_out/r1f_median3.log 2>&1 || echo median-fail
timeout 100 $N -k regex:clahe16 --launch-skip 4 -c 2 -o gpurun_out/r1f_clahe16 python benchmarks/op_once.py clahe16 1 > gpurun_out/r1f_clahe16.log 2>&1 || echo clahe16-fail
ls -la gpurun_out/r1f_*
