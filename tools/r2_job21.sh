#!/bin/bash
# ncu --set full of the two transform kernels of the three-kernel path at B = 4096; summarised on the box
mkdir -p gpurun_out /tmp/ncu21
for k in k_input_rfft k_irfft_out; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 45 -c 1 -o /tmp/ncu21/$k -f python bench.py --workload C5-4096 --steps 4 --warmup 41 --no-cpu --e2e-steps 3 > gpurun_out/ncu_$k.log 2>&1; echo "$k rc=$?"
done
{
  echo "# ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 45 -c 1 python bench.py --workload C5-4096 --steps 4 --warmup 41 --no-cpu --e2e-steps 3"
  for k in k_input_rfft k_irfft_out; do
    echo; echo "##### workload C5-4096 $k"
    python tools/ncu_summary.py /tmp/ncu21/$k.ncu-rep
    echo "--- top source lines by warp-stall samples"
    python tools/ncu_lines.py /tmp/ncu21/$k.ncu-rep k_ 16
    echo "--- stall reasons"
    python tools/ncu_stalls.py /tmp/ncu21/$k.ncu-rep 2>&1 | head -40
  done
} > gpurun_out/r02_c5_4096_split.txt
tail -5 gpurun_out/r02_c5_4096_split.txt
