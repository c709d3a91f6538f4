for r in 4 2 1; do echo "PGIBBS_HEAD_ROWS=$r"; PGIBBS_HEAD_ROWS=$r timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02zq_tmp.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/r02zq_tmp.json").read().strip().splitlines()[-1])
print("  C2", round(d["value"],2), "head_sample share", d["roofline"]["time_share_by_kernel"]["head_sample"])
for k,v in d["other_configs"].items():
    if k in ("C4_shard","C3","C3_all_positions","C5_shard_p25"): print("  ",k, round(v["iters_per_sec"],2), {kk:vv["avg_launch_us"] for kk,vv in v["kernels"].items() if "head_s" in kk})
PY
done
timeout 600 python -m pytest tests -m gpu -x -q -k "forward_logits_vs_oracle or device_scoring or golden or graph_replay or full_size" 2>&1 | tail -3
