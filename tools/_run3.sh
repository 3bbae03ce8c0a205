mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mesh_forward -c 1 -o gpurun_out/r2_c2_nodes_b -f python tools/profile_case.py c2 1 > gpurun_out/r2_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mesh_jacobian -c 1 -o gpurun_out/r2_c4_jac_b -f python tools/profile_case.py c4 1 --rows 8192 >> gpurun_out/r2_ncu2.log 2>&1
tail -5 gpurun_out/r2_ncu2.log
