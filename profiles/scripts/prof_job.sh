set -x
cd $GRAFT_REPO_ROOT
NCU="ncu --clock-control none"
# 1. launch list of the default bench command (after it exited 0 without ncu)
python bench.py --steps 1 --warmup 1 --sweeps 20 --no-cpu-baseline > gpurun_out/r02_bench_short.json 2>/dev/null || exit 1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --sweeps 20 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
# 2. full capture of both instantiations of the sweep kernel
python profiles/scripts/prof_c3.py 7 6 > /dev/null 2>&1 || exit 1
$NCU --set full --import-source on -k regex:k_sweep_rows -s 4 -c 4 -f -o gpurun_out/r02_sweep python profiles/scripts/prof_c3.py 7 6 > gpurun_out/ncu2.log 2>&1
# 3. DRAM traffic of config 5, two launches per sweep and the fused pass
for v in "" "ISING_STRIP_FUSE=1"; do
  env $v $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct -k regex:k_strip_ -c 4 --csv --log-file gpurun_out/r02_c5_traffic_${v:-phases}.csv python profiles/scripts/prof_c5.py > gpurun_out/ncu3.log 2>&1
done
# 4. config 4: launch list + full capture of the general sweep kernel
python profiles/scripts/prof_c4.py > /dev/null 2>&1 || exit 1
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/r02_c4_launches.csv python profiles/scripts/prof_c4.py > gpurun_out/ncu4.log 2>&1
$NCU --set full --import-source on -k regex:k_sweep_general -s 8 -c 2 -f -o gpurun_out/r02_c4_general python profiles/scripts/prof_c4.py > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out | tail -20
