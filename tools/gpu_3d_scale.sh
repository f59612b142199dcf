#!/bin/bash
# C5G7 3-D (BASELINE.json config 4) through the plugin with the planes sharded over 1, 2, 4, 8 devices:
# wall-clock of the MoC sweeper and of the whole solve, with the plugin's own split (upload / enqueue / host work
# between the inners / device wait + download + post-processing / device sweep kernels).
#   OUTERS=3 DEVS="0 0,1 0,1,2,3 0,1,2,3,4,5,6,7" tools/gpu_3d_scale.sh
mkdir -p gpurun_out/3d && cd gpurun_out/3d
cp ../../mocc_b200/bin/inputs/c5g7.xsl .
python ../../tools/make_c5g7_3d.py ../../mocc_b200/bin/inputs/c5g7_2d.xml c5g7_3d.xml --max-iter ${OUTERS:-3} --sn-inner 10 --moc-attrs 'tl_splitting="t"' > /dev/null
for d in ${DEVS:-0}; do
  n=$(echo $d | tr ',' '\n' | wc -l)
  OMP_NUM_THREADS=${THREADS:-$(nproc)} ../../mocc_b200/bin/mocc_b200_solve c5g7_3d.xml out_$n.arrays --set solver/sweeper@type=2d3d_cuda --set solver/sweeper/moc_sweeper/cuda@devices=$d > log_$n.txt 2>&1
  echo "devices=$n $(grep '^mocc_b200_solve:' log_$n.txt)"
  grep '^CudaMoCSweeper:' log_$n.txt
done
