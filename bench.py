"""Gibbs-iteration throughput benchmark (BASELINE.json metric) -- one JSON line on stdout from rank 0.

Workload at every N: BASELINE.json configs[1] per GPU -- ESM-1b 650M (33 x 1280, 20 heads, FFN 5120), 64
independent chains of L=256 (T=258 tokens), top_k=3 with burnin=0, all 256 positions resampled each iteration,
<mask> scatter on.  A "step" is one Gibbs iteration over the batch.  Chains shard across GPUs with no data-path
collective ("weak" scaling: 64 chains per GPU); NCCL is used once, to broadcast the weights from rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W]                 # this engine
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # the reference's CPU loop (oracle port)
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "esm1b_t33_650M_UR50S"
CHAINS_PER_GPU = 64
SEQ_LEN = 256
TOP_K = 3
BURNIN = 0
METRIC = "gibbs_iters_per_sec"
UNIT = "iters/s (one iter = 64 chains x 256 residues resampled, per GPU share)"
AA = "ACDEFGHIKLMNPQRSTVWY"


def workload_config(n_gpus):
    return {
        "workload": "BASELINE configs[1]: ESM-1b 650M single-seq Gibbs, batch=64 chains/GPU, L=256, top_k=3, "
                    "burnin=0, num_positions=0 (all 256 positions), mask=True",
        "chains_per_gpu": CHAINS_PER_GPU, "total_chains": CHAINS_PER_GPU * n_gpus, "seq_len": SEQ_LEN,
        "tokens_per_chain": SEQ_LEN + 2, "top_k": TOP_K, "burnin": BURNIN,
        "parallelism": "chains sharded, %d x 64" % n_gpus,
        "l2": "no flush: per-iteration working set (1.3 GB fp16 weights + >1 GB activations) >> 126 MB L2",
        "weights": "synthetic seeded N(0,0.02) (no pretrained checkpoints offline)",
    }


def algorithmic_flops_per_iter(cfg, B, T):
    d, F, V, L = cfg["embed_dim"], cfg["ffn_dim"], cfg["vocab"], cfg["layers"]
    f_token = L * (2 * (4 * d * d + 2 * d * F) + 4 * T * d) + 2 * (d * d + d * V)
    return B * T * f_token


def seeds(n, length, seed=1234):
    rng = random.Random(seed)
    return ["".join(rng.choice(AA) for _ in range(length)) for _ in range(n)]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tensor": d["bf16_tflops_sustained"], "tensor_burst": d["bf16_tflops"], "hbm": d["hbm_gbs"],
                "source": "measured"}
    return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------- CPU arms
_CPU_MODEL = None


def cpu_reference_run(n_chains, n_iters):
    """The reference's loop (oracle port: per-residue generate_step on the host + fp32 eager forward),
    on all the host threads torch uses by default."""
    global _CPU_MODEL
    import torch
    from oracle.fair_esm import OracleModel
    from oracle.gibbs_loop import esm_generate
    from protein_gibbs_sampler_b200.config import get_config
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    if _CPU_MODEL is None:
        cfg = get_config(MODEL)
        _CPU_MODEL = OracleModel(cfg, synthetic_state_dict(cfg, 0))
    model = _CPU_MODEL
    random.seed(0)
    torch.manual_seed(0)
    t0 = time.perf_counter()
    esm_generate(model, n_chains, seeds(n_chains, SEQ_LEN), batch_size=n_chains, num_iters=n_iters, top_k=TOP_K,
                 burnin=BURNIN)
    dt = time.perf_counter() - t0
    return dt, torch.get_num_threads()


def cpu_baseline_sample():
    """Bounded sample: 4 of the 64 chains for 1 iteration (~1.4 TFLOP of fp32 GEMM + 1024 generate_step calls)."""
    n_chains, n_iters = 4, 1
    dt, threads = cpu_reference_run(n_chains, n_iters)
    iters_per_s = (n_iters / dt) * (n_chains / CHAINS_PER_GPU)  # cost is linear in the number of chains
    return {"value": iters_per_s, "unit": "iters/s (64-chain iterations)", "cores": threads, "kind": "port",
            "sample": "%d of 64 chains x %d iteration(s) of the same workload in %.1f s, scaled linearly in chains; "
                      "reference sampler loop restated in oracle/gibbs_loop.py + fp32 eager forward "
                      "(fair-esm is not installable offline)" % (n_chains, n_iters, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_chains = 4
    times = []
    for i in range(args.warmup + args.steps):
        dt, threads = cpu_reference_run(n_chains, 1)
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    # one 64-chain iteration costs 16x the 4-chain sample; N GPUs' worth of chains cost N x that on the same host
    value = (1.0 / dt) * (n_chains / CHAINS_PER_GPU)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "each step = %d of 64 chains x 1 iteration (%.1f s), scaled linearly in chains"
                                   % (n_chains, dt)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.config import get_config
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = get_config(MODEL)

    # ---- weights: rank 0 creates them, one NCCL broadcast ships them (the only collective on this path)
    if world > 1:
        sd = synthetic_state_dict(cfg, 0) if rank == 0 else None
        meta = [{k: tuple(v.shape) for k, v in sd.items() if k != "lm_head.weight"}] if rank == 0 else [None]
        dist.broadcast_object_list(meta, src=0)
        gsd = {}
        for k, shape in meta[0].items():
            t = sd[k].to(dev) if rank == 0 else torch.empty(shape, dtype=torch.float32, device=dev)
            dist.broadcast(t, src=0)
            gsd[k] = t
        sd = gsd
    else:
        sd = synthetic_state_dict(cfg, 0)
    model = models.ESM1b(state_dict=sd)
    sampler = ESM_sampler(model, device="cuda:%d" % local, rng="device")
    del sd
    engine = model.model.engine
    # time on the stream the kernels are launched on: a dedicated torch stream shared with the engine
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    engine.set_stream(stream.cuda_stream)

    B, T = CHAINS_PER_GPU, SEQ_LEN + 2
    K, W = args.steps, args.warmup
    my_seeds = seeds(B * world, SEQ_LEN)[rank * B:(rank + 1) * B]
    tokens = model.batch_converter([(str(i), s) for i, s in enumerate(my_seeds)])[2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: tokens + schedule already in HBM when the timed region starts
    indexes, _ = sampler.calculate_indexes(None, 0, SEQ_LEN, False)
    plan, _ = sampler.plan_positions(B, indexes, -1, 0, False, W + K)
    engine.set_tokens(tokens)
    engine.set_schedule(plan.positions, plan.n_iters, plan.P, plan.iter_stride, plan.chain_stride)
    engine.set_noise(None)
    engine.set_device_rng(1234 + rank)
    engine.run(0, W, BURNIN, TOP_K, None, True, sampler.valid_aa_idx)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = engine.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    torch.cuda.profiler.start()   # `ncu --profile-from-start off` then captures exactly the timed region
    ev0.record(stream)
    engine.run(W, K, BURNIN, TOP_K, None, True, sampler.valid_aa_idx)
    ev1.record(stream)
    barrier()
    torch.cuda.profiler.stop()
    host_ms = (time.perf_counter() - t_host0) * 1000.0
    ms = ev0.elapsed_time(ev1)
    # the device interval can never exceed the host wall clock around it by more than jitter
    assert ms <= host_ms * 1.05 + 1.0 and ms >= 0.5 * host_ms, (ms, host_ms)
    launches = engine.launch_count() - launches0
    clock_info = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * K / (ms / 1000.0)

    # ---- end to end through the public API: host strings in, host strings out (H2D + D2H inside)
    random.seed(rank)
    torch.manual_seed(rank)
    sampler.generate(B, my_seeds, batch_size=B, num_iters=min(W, 3), top_k=TOP_K, burnin=BURNIN,
                     show_progress_bar=False)
    barrier()
    t0 = time.perf_counter()
    out = sampler.generate(B, my_seeds, batch_size=B, num_iters=K, top_k=TOP_K, burnin=BURNIN,
                           show_progress_bar=False)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert len(out) == B and all(len(s) == SEQ_LEN for s in out)
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_value = world * K / e2e_s
    h2d = (B * T * 4 + plan.P * 4 + len(sampler.valid_aa_idx) * 4) / K   # tokens + schedule + candidate ids, once per call
    d2h = (B * T * 4) / K

    # ---- per-kernel-class timing (separate pass with CUDA events around every launch) for the roofline
    roofline = None
    if rank == 0:
        n_prof = min(K, 5)
        engine.set_tokens(tokens)
        engine.set_schedule(plan.positions, plan.n_iters, plan.P, plan.iter_stride, plan.chain_stride)
        engine.profile_enable(True)
        engine.run(W, n_prof, BURNIN, TOP_K, None, True, sampler.valid_aa_idx)
        engine.sync()
        prof = engine.profile_read()
        engine.profile_enable(False)
        peaks = measured_peaks()
        M, d, F = B * T, cfg["embed_dim"], cfg["ffn_dim"]
        gemm_flops = {"gemm_qkv": 2.0 * M * 3 * d * d, "gemm_out": 2.0 * M * d * d, "gemm_fc1": 2.0 * M * d * F,
                      "gemm_fc2": 2.0 * M * d * F, "gemm_head": 2.0 * (B * SEQ_LEN) * d * d}
        total_ms = sum(v[0] for v in prof.values())
        shares = {k: round(v[0] / total_ms, 4) for k, v in prof.items()}
        dom = max((k for k in prof if k in gemm_flops), key=lambda k: prof[k][0])
        avg_ms = prof[dom][0] / prof[dom][1]
        achieved = gemm_flops[dom] / (avg_ms / 1000.0) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom)
        roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["tensor"],
                    "unit": "TFLOP/s", "frac": achieved / peaks["tensor"], "traffic": traffic,
                    "peak_source": peaks["source"] + " bf16_tflops_sustained",
                    "avg_launch_ms": avg_ms, "algorithmic_flops_per_launch": gemm_flops[dom],
                    "time_share_by_kernel": shares,
                    "step": {"algorithmic_tflop_per_iter": algorithmic_flops_per_iter(cfg, B, T) / 1e12,
                             "achieved_tflops": algorithmic_flops_per_iter(cfg, B, T) * (value / world) / 1e12,
                             "frac_of_sustained_peak": algorithmic_flops_per_iter(cfg, B, T) * (value / world) / 1e12
                             / peaks["tensor"]}}

    cpu = cpu_baseline_sample() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate/residual/softmax/LayerNorm", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "one ESM_sampler.generate call of K iterations: host strings -> tokens -> H2D -> K "
                            "on-device iterations -> D2H -> strings; bytes are per-call totals / K"},
            "gpu_launches": launches, "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu,
            "chain_iters_per_sec": value * CHAINS_PER_GPU, "residue_updates_per_sec": value * CHAINS_PER_GPU * SEQ_LEN,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
