# round-2 ncu captures AFTER the half-box epilogue (run under gpurun on one B200); text summaries go to gpurun_out/
set -x
for c in ffn1_fwd qkv_fwd attnout_fwd attnout_dgrad qkv_dgrad ffn2_dgrad ffn2_fwd ffn1_dgrad; do
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -f -o /tmp/r02f_$c python tools/profile_gemm.py $c --tokens 262144 --iters 2 2>&1 | tail -1
  python tools/ncu_summary.py /tmp/r02f_$c.ncu-rep "$c (M=262144, bench shape of the passage tower), half-box epilogue build, ncu --set full --clock-control none" > gpurun_out/r02f_${c}_ncu.txt
done
cp /tmp/r02f_attnout_fwd.ncu-rep /tmp/r02f_ffn1_fwd.ncu-rep gpurun_out/
