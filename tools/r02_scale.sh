# round-2 scaling evidence on one 8-GPU box: N=1 and N=8 of bench.py back to back (FlatGradSync and stock DDP), then the
# PROD / LEAD workloads at N=8
set -x
python bench.py --steps 10 --warmup 3 --no-extras --no-profile --no-cpu-baseline > gpurun_out/r02_scale_n1.json 2>/dev/null
for mode in 0 1; do
  SIMX_BENCH_DDP=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$mode bench.py --gpus 8 --steps 10 --warmup 3 --no-profile > gpurun_out/r02_scale_n8_ddp$mode.json 2> gpurun_out/r02_scale_n8.err
done
python - <<'PY'
import json
for f in ("r02_scale_n1", "r02_scale_n8_ddp0", "r02_scale_n8_ddp1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "dropout", d["dropout"] and round(d["dropout"]["ms_per_step"], 2), d["clocks"], d["config"].get("gradient_exchange"))
    except Exception as e:
        print(f, "FAILED", e)
PY
for w in lead prod_kd; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 tools/bench_workloads.py --workload $w --steps 5 --out gpurun_out/r02_workloads_8gpu.jsonl 2>&1 | grep '^{' | cut -c1-420
done
