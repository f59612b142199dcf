#!/bin/bash
W=/tmp/c53d; mkdir -p $W; cp mocc_b200/bin/inputs/c5g7.xsl $W/
python tools/make_c5g7_3d.py mocc_b200/bin/inputs/c5g7_2d.xml $W/c5g7_3d.xml --max-iter ${ITERS:-1} > /dev/null
cd $W
for dev in ${DEVS:-0}; do
  echo "== devices $dev"
  OMP_NUM_THREADS=${THREADS:-1} /root/repo/mocc_b200/bin/mocc_b200_solve c5g7_3d.xml out.arrays --set solver/sweeper@type=2d3d_cuda --set solver/sweeper/moc_sweeper/cuda@devices=$dev > solve_$dev.log 2>&1
  grep -v "^ *[0-9.]* [0-9]* " solve_$dev.log | grep -i -E "time|sweep|Sn|CMFD|mocc_b200_solve|source|ray|CudaMoCSweeper" | head -40
done
