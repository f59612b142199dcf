#!/bin/bash
# C5G7-2D / 3x3 eigenvalue solves: reference CPU sweeper vs the plugin (default, group-batched, Jacobi boundary),
# k-eff (pcm), FSR flux (max rel.), outers, time to converge. Output: gpurun_out/solve_parity.jsonl
W=/tmp/solve; mkdir -p $W gpurun_out; cp mocc_b200/bin/inputs/* $W/; cd $W
S=/root/repo/mocc_b200/bin/mocc_b200_solve
: > /root/repo/gpurun_out/solve_parity.jsonl
for c in 3x3 c5g7_2d; do
  OMP_NUM_THREADS=$(nproc) $S $c.xml ${c}_ref.arrays > ${c}_ref.log 2>&1; tail -1 ${c}_ref.log
  $S $c.xml ${c}_cuda.arrays --set solver/sweeper@type=moc_cuda > ${c}_cuda.log 2>&1; tail -1 ${c}_cuda.log
  $S $c.xml ${c}_cudab.arrays --set solver/sweeper@type=moc_cuda --set solver/sweeper/cuda@group_batch=t > ${c}_cudab.log 2>&1; tail -1 ${c}_cudab.log
  $S $c.xml ${c}_cudaj.arrays --set solver/sweeper@type=moc_cuda --set solver/sweeper@boundary_update=jacobi > ${c}_cudaj.log 2>&1; tail -1 ${c}_cudaj.log
  for v in cuda cudab cudaj; do python /root/repo/tools/compare_solves.py ${c}_ref.arrays ${c}_$v.arrays >> /root/repo/gpurun_out/solve_parity.jsonl; done
done
cat /root/repo/gpurun_out/solve_parity.jsonl | cut -c1-600
