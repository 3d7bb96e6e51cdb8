set -x
ncu --set full --clock-control none --import-source on -k regex:ln_fwd_kernel -s 2 -c 1 -f -o /tmp/r02j_ln_fwd python tools/profile_rows.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:ln_bwd_kernel -s 2 -c 1 -f -o /tmp/r02j_ln_bwd python tools/profile_rows.py 2>&1 | tail -1
python tools/ncu_summary.py /tmp/r02j_ln_fwd.ncu-rep "ln_fwd (T=262144, H=768, bf16), ncu --set full --clock-control none" > gpurun_out/r02j_ln_fwd_ncu.txt
python tools/ncu_summary.py /tmp/r02j_ln_bwd.ncu-rep "ln_bwd (T=262144, H=768, bf16), ncu --set full --clock-control none" > gpurun_out/r02j_ln_bwd_ncu.txt
ncu -i /tmp/r02j_ln_fwd.ncu-rep --page source --csv > gpurun_out/r02j_ln_fwd_source.csv 2>/dev/null
ncu -i /tmp/r02j_ln_bwd.ncu-rep --page source --csv > gpurun_out/r02j_ln_bwd_source.csv 2>/dev/null
ls -la gpurun_out/ | tail -5
