cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/r2_pcie_chain.jsonl
for N in 1 2 4 8; do
  for AG in 4 16; do
    timeout 120 $TR --nproc-per-node $N --master-port $((29600+N)) scripts/pcie_chain_probe.py --agents $AG 2>/dev/null | grep '^{' >> gpurun_out/r2_pcie_chain.jsonl
  done
done
cat gpurun_out/r2_pcie_chain.jsonl
timeout 600 $TR --nproc-per-node 8 --master-port 29620 scripts/coach_loop.py --iters 3 --games 65536 --batch 8192 --ddp-train local > gpurun_out/r2_coach_loop_n8_local.jsonl 2> gpurun_out/r2_coach_loop_n8_local.err
tail -c 400 gpurun_out/r2_coach_loop_n8_local.err; cut -c1-700 gpurun_out/r2_coach_loop_n8_local.jsonl
