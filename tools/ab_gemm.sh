# same-box A/B of two builds of the library: simxns_b200/libsimx_b200_old.so (built from the previous commit) against the
# current one, op-level GEMM timings at the passage-tower shapes (tools/profile_gemm.py), interleaved twice
CASES="${CASES:-qkv_fwd attnout_fwd ffn1_fwd ffn2_fwd ffn2_dgrad ffn1_dgrad qkv_dgrad}"
cp simxns_b200/libsimx_b200.so /tmp/new.so
cp simxns_b200/libsimx_b200_old.so /tmp/old.so
for rep in 1 2; do
for lib in old new; do
  cp /tmp/$lib.so simxns_b200/libsimx_b200.so
  for c in $CASES; do
    echo -n "$lib "; python tools/profile_gemm.py $c --tokens 262144 --iters 20 2>&1 | tail -1
  done
done
done
cp /tmp/new.so simxns_b200/libsimx_b200.so
