# SM clock and board power while one GEMM epilogue variant runs back to back for several seconds
for e in 0 7 2; do
  (python tools/profile_gemm.py ffn1_fwd --tokens 262144 --iters 8000 --epi $e 2>&1 | tail -1) &
  pid=$!
  sleep 7
  for i in 1 2 3 4; do nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.sw_power_cap --format=csv,noheader; sleep 0.5; done
  wait $pid
done
