#!/bin/bash
W=/tmp/s9; mkdir -p $W; cp tests/golden/inputs/* $W/; cd $W
S=/root/repo/mocc_b200/bin/mocc_b200_solve
export OMP_NUM_THREADS=1
run() { name=$1; shift; $S "$@" > $name.log 2>&1; tail -1 $name.log; }
run a_ref mini3d.xml a_ref.arrays
run a_cuda mini3d.xml a_cuda.arrays --set solver/sweeper@type=moc_cuda
run b_ref mini3d.xml b_ref.arrays --set solver@cmfd=t
run b_cuda mini3d.xml b_cuda.arrays --set solver@cmfd=t --set solver/sweeper@type=moc_cuda
run c_ref mini2d3d.xml c_ref.arrays --set solver@max_iter=3
run c_cuda mini2d3d.xml c_cuda.arrays --set solver@max_iter=3 --set solver/sweeper@type=2d3d_cuda
grep -E "^ +[0-9.]+ +[0-9]+ " c_ref.log | head -5; grep -E "^ +[0-9.]+ +[0-9]+ " c_cuda.log | head -5
