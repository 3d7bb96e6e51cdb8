# round-2 scaling evidence on one 8-GPU box: N=1 then N=2,4,8 of bench.py back to back (FlatGradSync), per-rank compute table
set -x
python bench.py --steps 10 --warmup 3 --no-extras --no-profile --no-cpu-baseline > gpurun_out/r02e_scale_n1.json 2>/dev/null
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 10 --warmup 3 --no-profile > gpurun_out/r02e_scale_n$n.json 2> gpurun_out/r02e_scale.err
done
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/r02e_scale_n{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "dropout", d["dropout"] and round(d["dropout"]["ms_per_step"], 2), d["clocks"]["sm_mhz"], d.get("rank_compute"))
    except Exception as e:
        print(n, "FAILED", e)
PY
