# round-2 ncu captures (run under gpurun on one B200); text summaries go to gpurun_out/, the .ncu-rep of the two kernels
# worked on next are kept for the source page
set -x
python -m pytest "tests/test_gpu_large_hidden.py" -q 2>&1 | grep -E "assert|Error|passed|failed" | head -8 | cut -c1-300
for c in ffn1_fwd qkv_fwd attnout_fwd attnout_dgrad qkv_dgrad ffn2_dgrad; do
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -f -o /tmp/r02_$c python tools/profile_gemm.py $c --tokens 262144 --iters 2 2>&1 | tail -1
  python tools/ncu_summary.py /tmp/r02_$c.ncu-rep "$c (M=262144, bench shape of the passage tower), ncu --set full --clock-control none" > gpurun_out/r02_${c}_ncu.txt
done
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 2 -f -o /tmp/r02_attn python tools/profile_attn.py --iters 1 2>&1 | tail -2
python tools/ncu_summary.py /tmp/r02_attn.ncu-rep "attention fwd / bwd, 2048 x 12 heads x S=128" > gpurun_out/r02_attn_ncu.txt
cp /tmp/r02_ffn1_fwd.ncu-rep /tmp/r02_attn.ncu-rep gpurun_out/
