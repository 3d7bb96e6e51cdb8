# ln_fwd ring-parameter sweep.  Build the variants first (from simxns_b200/csrc, after `make`):
#   for v in "4 2" "2 4" "4 1" "3 4"; do set -- $v; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
#     -Xcompiler -fPIC -DSIMX_LNF_SLOTS=$1 -DSIMX_LNF_G=$2 -c rowops.cu -o /tmp/rowops_$1_$2.o && nvcc -gencode \
#     arch=compute_100a,code=sm_100a -shared -o ../libsimx_b200_lnf_$1_$2.so $(ls build/*.o | grep -v rowops.o) \
#     /tmp/rowops_$1_$2.o -cudart static; done
cp simxns_b200/libsimx_b200.so /tmp/base.so
for v in base lnf_4_2 lnf_2_4 lnf_4_1 lnf_3_4 base; do
  if [ $v = base ]; then cp /tmp/base.so simxns_b200/libsimx_b200.so; else cp simxns_b200/libsimx_b200_$v.so simxns_b200/libsimx_b200.so; fi
  python bench.py --layers 4 --steps 10 --no-extras --no-cpu-baseline --no-dropout-run 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks={k['name']:k for k in d['kernels']}
print('$v', 'ln_fwd ms', round(ks['ln_fwd']['ms'],3), 'GB/s', round(ks['ln_fwd']['gbs']), 'step', round(d['ms_per_step'],2))"
done
cp /tmp/base.so simxns_b200/libsimx_b200.so
