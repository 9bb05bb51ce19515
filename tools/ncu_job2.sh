set -x
cd $GRAFT_REPO_ROOT
timeout 200 python tools/gemm_probe.py > gpurun_out/r02_plain_gemm.log 2>&1 && \
timeout 400 ncu --clock-control none --set full --import-source on -k regex:k_gemm_i8limb -c 1 -o gpurun_out/r02_gemm_k4096 python tools/gemm_probe.py > gpurun_out/r02_ncu_gemm2.log 2>&1
echo rc=$?
timeout 200 python tools/gemm_probe.py 16384 8192 4096 4294967291 >> gpurun_out/r02_plain_gemm.log 2>&1
cat gpurun_out/r02_plain_gemm.log
