set -x
cd $GRAFT_REPO_ROOT
NCU="ncu --clock-control none"
timeout 300 python tools/profile_step.py > gpurun_out/r02_plain_step.log 2>&1 && \
timeout 1200 $NCU --metrics gpu__time_duration.sum -c 120000 --csv --log-file gpurun_out/r02_launches_c2_full.csv python tools/profile_step.py > gpurun_out/r02_ncu_step.log 2>&1
echo launches rc=$?
timeout 600 $NCU --set full --import-source on -k regex:k_gemm_i8limb -s 40 -c 2 -o gpurun_out/r02_gemm python tools/profile_step.py --rows 100000 > gpurun_out/r02_ncu_gemm.log 2>&1
echo gemm rc=$?
SPASM_B200_SCHUR_DENSE=0 timeout 600 $NCU --set full --import-source on -k regex:"k_solve_global|k_solve_smem" -c 2 -o gpurun_out/r02_solve python tools/sparse_probe.py 64 > gpurun_out/r02_ncu_solve.log 2>&1
echo solve rc=$?
timeout 600 $NCU --set full --import-source on -k regex:k_sptrsm_seq -s 1 -c 1 -o gpurun_out/r02_sptrsm python tools/sparse_probe.py 64 > gpurun_out/r02_ncu_sptrsm.log 2>&1
echo sptrsm rc=$?
ls -la gpurun_out | tail -12
