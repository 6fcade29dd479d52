#!/bin/bash
# Round-2 profiling captures (run under gpurun on one B200; every ncu run follows the same command exiting 0 without ncu).
set -x
B="python bench.py --steps 2 --warmup 3 --kernels-only --workload c3"
timeout 120 $B > gpurun_out/r02_plain_c3.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
B1="python bench.py --steps 1 --warmup 3 --kernels-only --workload c3"
timeout 120 $B1 > gpurun_out/r02_plain_c3b.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gravity_allpairs -s 3 -c 1 -o gpurun_out/r02_allpairs -f $B1 > gpurun_out/r02_ncu_allpairs.log 2>&1
B2="python bench.py --workload c3 --gravity tree --steps 1 --warmup 3 --kernels-only"
timeout 120 $B2 > gpurun_out/r02_plain_c3tree.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_tree_walk|k_cell_neighbors|k_density|k_pressure_grad|k_lbvh_nodes|k_permute_cells|k_integrate" -s 21 -c 7 -o gpurun_out/r02_tree_sph -f $B2 > gpurun_out/r02_ncu_tree.log 2>&1
B3="python bench.py --workload c4 --steps 1 --warmup 2 --kernels-only"
timeout 200 $B3 > gpurun_out/r02_plain_c4.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv $B3 > gpurun_out/r02_ncu_launches_c4.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 5 > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_final_ref.json 2> gpurun_out/r02_bench_final_ref.err
timeout 60 ./planetmodel-sph_b200/host_cpp/host_demo 3000 5 tree > gpurun_out/r02_host_demo.txt 2>&1
ls -la gpurun_out | tail -20
