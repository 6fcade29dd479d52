#!/bin/bash
# Round-1 profiling captures (run under gpurun on one B200; every ncu run follows the same command exiting 0 without ncu).
set -x
B="python bench.py --steps 2 --warmup 3 --kernels-only"
timeout 120 $B > gpurun_out/r01_plain_c3.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_c3.csv $B > gpurun_out/r01_ncu_launches.log 2>&1
B1="python bench.py --steps 1 --warmup 3 --kernels-only"
timeout 120 $B1 > gpurun_out/r01_plain_c3b.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gravity_allpairs -s 3 -c 1 -o gpurun_out/r01_allpairs -f $B1 > gpurun_out/r01_ncu_allpairs.log 2>&1
B2="python bench.py --gravity tree --steps 1 --warmup 3 --kernels-only"
timeout 120 $B2 > gpurun_out/r01_plain_c3tree.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_tree_walk|k_cell_neighbors|k_density|k_pressure_grad|k_lbvh_nodes|k_permute_cells|k_integrate" -s 21 -c 7 -o gpurun_out/r01_tree_sph -f $B2 > gpurun_out/r01_ncu_tree.log 2>&1
timeout 200 python bench.py --workload c4 --steps 2 --warmup 3 --kernels-only > gpurun_out/r01_c4.json 2> gpurun_out/r01_c4.err
make -C profiles/tools tune_allpairs > /dev/null 2>&1
timeout 120 ./profiles/tools/tune_allpairs 262144 > gpurun_out/r01_tune_allpairs.txt 2>&1
timeout 60 ./planetmodel-sph_b200/host_cpp/host_demo 3000 5 tree > gpurun_out/r01_host_demo.txt 2>&1
ls -la gpurun_out | tail -20
