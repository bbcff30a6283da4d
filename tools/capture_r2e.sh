tag=r2e
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --also "" --tool-files 0 > gpurun_out/${tag}_launch_bench.log 2>&1
k=gmm_tc4_kernel; n=gmm_tc4
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -o gpurun_out/${tag}_$n -f python bench.py --steps 2 --warmup 3 --no-cpu --also "" --tool-files 0 > gpurun_out/${tag}_ncu_$n.log 2>&1
if [ -f gpurun_out/${tag}_$n.ncu-rep ]; then
   ncu -i gpurun_out/${tag}_$n.ncu-rep --page details --csv > gpurun_out/${tag}_${n}_details.csv 2>/dev/null
   python tools/ncu_summary.py gpurun_out/${tag}_$n.ncu-rep > gpurun_out/${tag}_${n}_raw.txt 2>/dev/null
   python tools/ncu_lines.py gpurun_out/${tag}_$n.ncu-rep 14 >> gpurun_out/${tag}_${n}_raw.txt 2>/dev/null
   rm -f gpurun_out/${tag}_$n.ncu-rep
fi
ls -la gpurun_out/ | grep r2e
