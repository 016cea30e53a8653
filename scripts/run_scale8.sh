set -x
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
nproc >> gpurun_out/r2_topo.txt; numactl -H >> gpurun_out/r2_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r2_topo.txt
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-alt --sustain-seconds 0 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
  else
    timeout 400 $TR --nproc-per-node $N --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-alt --sustain-seconds 0 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
  fi
  tail -c 300 gpurun_out/r2_scale_n$N.err
done
timeout 600 $TR --nproc-per-node 8 --master-port 29520 scripts/coach_loop.py --iters 3 --games 65536 --batch 8192 --ddp-train > gpurun_out/r2_coach_loop_n8.jsonl 2> gpurun_out/r2_coach_loop_n8.err
tail -c 600 gpurun_out/r2_coach_loop_n8.err
python - <<'PY'
import json
for N in (1,2,4,8):
    try:
        d = json.loads(open(f"gpurun_out/r2_scale_n{N}.json").read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        ec = d.get("e2e_coach") or {}
        print(N, "value", round(d["value"]/1e6,2), "e2e", round((e.get("value") or 0)/1e6,2), "issue_frac", e.get("host_issue_fraction"), "cpu", e.get("host_cpu_cores_busy"), "pcie", e.get("pcie_gbs_h2d"), e.get("pcie_gbs_d2h"), "e2e_coach", round((ec.get("value") or 0)/1e6,2), "gather", d.get("example_gather"))
    except Exception as ex:
        print(N, "failed", ex)
for l in open("gpurun_out/r2_coach_loop_n8.jsonl"):
    print(l.strip()[:900])
PY
