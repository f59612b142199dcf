#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
W=/tmp/solve; mkdir -p $W; cp mocc_b200/bin/inputs/* $W/; cd $W
S=/root/repo/mocc_b200/bin/mocc_b200_solve
export OMP_NUM_THREADS=$(nproc)
for c in 3x3 c5g7_2d; do
  $S $c.xml ${c}_ref.arrays > ${c}_ref.log 2>&1; tail -1 ${c}_ref.log
  $S $c.xml ${c}_cuda.arrays --set solver/sweeper@type=moc_cuda > ${c}_cuda.log 2>&1; tail -1 ${c}_cuda.log
  $S $c.xml ${c}_cudab.arrays --set solver/sweeper@type=moc_cuda --set solver/sweeper/cuda@group_batch=t > ${c}_cudab.log 2>&1; tail -1 ${c}_cudab.log
  $S $c.xml ${c}_cudaj.arrays --set solver/sweeper@type=moc_cuda --set solver/sweeper@boundary_update=jacobi > ${c}_cudaj.log 2>&1; tail -1 ${c}_cudaj.log
done
cd /root/repo
for c in 3x3 c5g7_2d; do for v in cuda cudab cudaj; do python tools/compare_solves.py $W/${c}_ref.arrays $W/${c}_$v.arrays; done; done > gpurun_out/solve_parity.jsonl 2>&1
cat gpurun_out/solve_parity.jsonl
cp $W/*.log gpurun_out/ 2>/dev/null
cp $W/3x3_ref.arrays gpurun_out/
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_pergroup.json 2> gpurun_out/bench_pergroup.err; cat gpurun_out/bench_pergroup.json; tail -3 gpurun_out/bench_pergroup.err
python bench.py --steps 5 --warmup 3 --mode batched --no-cpu-baseline > gpurun_out/bench_batched.json 2> gpurun_out/bench_batched.err; cat gpurun_out/bench_batched.json; tail -3 gpurun_out/bench_batched.err
