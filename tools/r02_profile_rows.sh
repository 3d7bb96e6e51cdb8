# ncu --set full of the LayerNorm kernels at the passage-tower shape (one encoder layer forward + backward, T = 262 144)
set -x
TAG=${TAG:-r02w}
ncu --set full --clock-control none --import-source on -k regex:ln_fwd -s 2 -c 1 -f -o /tmp/${TAG}_ln_fwd python tools/profile_rows.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:ln_bwd_kernel -s 2 -c 1 -f -o /tmp/${TAG}_ln_bwd python tools/profile_rows.py 2>&1 | tail -1
python tools/ncu_summary.py /tmp/${TAG}_ln_fwd.ncu-rep "ln_fwd (T=262144, H=768, bf16), ncu --set full --clock-control none" > gpurun_out/${TAG}_ln_fwd_ncu.txt
python tools/ncu_summary.py /tmp/${TAG}_ln_bwd.ncu-rep "ln_bwd (T=262144, H=768, bf16), ncu --set full --clock-control none" > gpurun_out/${TAG}_ln_bwd_ncu.txt
