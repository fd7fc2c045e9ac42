set -x
python bench.py --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu-baseline --no-graph > gpurun_out/r3_plain.json 2> gpurun_out/r3_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c5_final.csv python bench.py --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu-baseline --no-graph > gpurun_out/r3_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fit_contract -s 6 -c 1 -o gpurun_out/r02_fit_contract_final python bench.py --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu-baseline --no-graph > gpurun_out/r3_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
