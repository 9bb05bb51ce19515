#!/bin/bash
# final evidence of the round on one GPU: the default bench line, a complete ncu launch list of one step, and a full
# ncu capture of the tile Gauss-Jordan kernel (numbers printed under ncu are never bench values)
cd /root/repo
timeout 600 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
echo "bench rc=$?"
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 120000 --csv --log-file gpurun_out/r02_final_launches.csv python tools/profile_step.py > gpurun_out/r02_final_ncu_step.log 2>&1
echo "launch list rc=$?"
timeout 300 $NCU --set full --import-source on -k regex:k_tile_gauss_cluster -s 200 -c 1 -o gpurun_out/r02_tile_gauss python tools/profile_step.py --rows 60000 > gpurun_out/r02_ncu_tile.log 2>&1
echo "tile capture rc=$?"
tail -c 600 gpurun_out/r02_final_bench.json | head -c 600; echo
ls -la gpurun_out/r02_final* gpurun_out/r02_tile*
