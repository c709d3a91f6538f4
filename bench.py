"""Gibbs-iteration throughput benchmark (BASELINE.json metric) -- one JSON line on stdout from rank 0.

Headline workload at every N: BASELINE.json configs[1] per GPU -- ESM-1b 650M (33 x 1280, 20 heads, FFN 5120), 64
independent chains of L=256 (T=258 tokens), top_k=3 with burnin=0, all 256 positions resampled each iteration,
<mask> scatter on.  A "step" is one Gibbs iteration over the batch.  Chains shard across GPUs with no data-path
collective ("weak" scaling: 64 chains per GPU); NCCL is used once, to broadcast the packed weight blob from rank 0.
The per-GPU shards of the other BASELINE configs (3, 4, 5) and the split-operand precision mode are timed in the same
run and reported under "other_configs" (5 timed steps each).

  python bench.py [--gpus N] [--steps K] [--warmup W]                 # this engine
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # the reference's CPU loop on the host cores
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
        "precision": "fast (one tensor-core pass of fp16 operands per GEMM); split-operand mode under other_configs",
        "launch": "one stream, programmatic dependent launch, CUDA-graph replay from the 2nd iteration of a run, LayerNorm "
                  "forked next to the residual GEMMs' last wave (switches: PGIBBS_PDL / PGIBBS_GRAPH / PGIBBS_TAIL_OVERLAP)",
    }


def algorithmic_flops_per_iter(cfg, B, T):
    """GEMMs + attention matmuls, 2 FLOP per MAC, LM head over all tokens as the reference computes it."""
    d, F, V, L = cfg["embed_dim"], cfg["ffn_dim"], cfg["vocab"], cfg["layers"]
    f_token = L * (2 * (4 * d * d + 2 * d * F) + 4 * T * d) + 2 * (d * d + d * V)
    return B * T * f_token


def msa_flops_per_iter(cfg, B, R, C):
    d, F, V, L = cfg["embed_dim"], cfg["ffn_dim"], cfg["vocab"], cfg["layers"]
    return B * R * C * (L * (2 * (8 * d * d + 2 * d * F) + 4 * C * d + 4 * R * d) + 2 * (d * d + d * V))


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
        return self

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
REF_SAMPLE_CHAINS = 4          # of the 64 chains of one GPU's batch; the loop's cost is linear in the chain count


def _host_threads():
    """All the host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to every rank)."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _reference_sampler_class():
    """The UNMODIFIED reference sampler (`pgen.esm_sampler.ESM_sampler`, installed with pip from /root/reference into
    baseline/_ref by __graft_entry__.build(); git-ignored, ships to the GPU box), or None -> the oracle's port of the
    same loop.  Either way the forward is the fp32 eager restatement in oracle/fair_esm.py: fair-esm itself cannot be
    installed offline."""
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"):
        if os.path.isdir(os.path.join(p, "pgen")):
            sys.path.insert(0, p)
            try:
                from pgen.esm_sampler import ESM_sampler as RefSampler
                return RefSampler, p
            except Exception:
                sys.path.remove(p)
    return None, None


def cpu_reference_run(n_chains, n_iters):
    """One bounded sample of the C2 workload on the host: `n_chains` chains x `n_iters` iterations of the reference
    loop (per-residue generate_step + one fp32 forward per iteration).  Returns (seconds, threads, kind)."""
    global _CPU_MODEL
    import torch
    from oracle.fair_esm import OracleModel
    from protein_gibbs_sampler_b200.config import get_config
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    threads = _host_threads()
    if _CPU_MODEL is None:
        cfg = get_config(MODEL)
        _CPU_MODEL = OracleModel(cfg, synthetic_state_dict(cfg, 0))
    model = _CPU_MODEL
    ref_cls, _ = _reference_sampler_class()
    random.seed(0)
    torch.manual_seed(0)
    kw = dict(batch_size=n_chains, num_iters=n_iters, top_k=TOP_K, burnin=BURNIN)
    t0 = time.perf_counter()
    if ref_cls is not None:
        out = ref_cls(model, device="cpu").generate(n_chains, seeds(n_chains, SEQ_LEN), show_progress_bar=False, **kw)
        kind = "reference"
    else:
        from oracle.gibbs_loop import esm_generate
        out = esm_generate(model, n_chains, seeds(n_chains, SEQ_LEN), **kw)
        kind = "port"
    dt = time.perf_counter() - t0
    assert len(out) == n_chains and all(len(s) == SEQ_LEN for s in out)
    return dt, threads, kind


def _cpu_kind_note(kind):
    return ("unmodified reference sampler loop (pgen.esm_sampler.ESM_sampler.generate from baseline/_ref) driving the "
            "fp32 eager forward of oracle/fair_esm.py" if kind == "reference" else
            "oracle/gibbs_loop.py port of the reference sampler loop + the fp32 eager forward of oracle/fair_esm.py") + \
        " (fair-esm is not installable offline)"


def cpu_baseline_sample():
    """Bounded sample: 4 of the 64 chains for 1 iteration (~1.4 TFLOP of fp32 GEMM + 1024 generate_step calls)."""
    n_chains, n_iters = REF_SAMPLE_CHAINS, 1
    cpu_reference_run(n_chains, n_iters)                  # page in the weights / warm the thread pool
    dt, threads, kind = cpu_reference_run(n_chains, n_iters)
    iters_per_s = (n_iters / dt) * (n_chains / CHAINS_PER_GPU)
    return {"value": iters_per_s, "unit": "iters/s (64-chain iterations)", "cores": threads, "kind": kind,
            "sample": "%d of 64 chains x %d iteration(s) of the same workload measured in %.2f s and scaled x%d "
                      "(linear in chains); %s" % (n_chains, n_iters, dt, CHAINS_PER_GPU // n_chains, _cpu_kind_note(kind))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_chains = REF_SAMPLE_CHAINS
    scale = CHAINS_PER_GPU // n_chains
    times = []
    for i in range(args.warmup + args.steps):
        dt, threads, kind = cpu_reference_run(n_chains, 1)
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    # one 64-chain iteration costs `scale` x the sample on the same host cores
    value = (1.0 / dt) / scale
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        # what was actually timed: one step = the bounded sample; the 64-chain figure is `value`
        "ms_per_step": 1000.0 * dt, "ms_per_step_is": "measured sample (%d of 64 chains x 1 iteration)" % n_chains,
        "sample_scale": scale, "ms_per_64_chain_iteration_extrapolated": 1000.0 * dt * scale,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "each step = %d of 64 chains x 1 iteration, measured %.2f s, value scaled x%d "
                                   "(linear in chains); %s" % (n_chains, dt, scale, _cpu_kind_note(kind))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
class Timer:
    """Device-timed region on the engine's stream, max over ranks."""

    def __init__(self, torch, dist, world, dev, stream):
        self.torch, self.dist, self.world, self.dev, self.stream = torch, dist, world, dev, stream

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn, profiler=False):
        torch = self.torch
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        if profiler:
            torch.cuda.profiler.start()   # `ncu --profile-from-start off` then captures exactly the timed region
        ev0.record(self.stream)
        fn()
        ev1.record(self.stream)
        self.barrier()
        if profiler:
            torch.cuda.profiler.stop()
        host_ms = (time.perf_counter() - t0) * 1000.0
        ms = ev0.elapsed_time(ev1)
        # the device interval can never exceed the host wall clock around it by more than jitter
        assert ms <= host_ms * 1.05 + 1.0 and ms >= 0.5 * host_ms - 1.0, (ms, host_ms)
        return self.max_over_ranks(ms)

    def max_over_ranks(self, v):
        if self.world > 1:
            t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            v = t.item()
        return v


def _model_on_ranks(models_mod, cls, cfg, world, rank, dev, precision="fast"):
    """Rank 0 creates the synthetic weights; one NCCL broadcast of the packed blob ships them (the only collective on
    the data path).  Split-operand precision needs the fp32 weights, so its blob is not packed to fp16."""
    from protein_gibbs_sampler_b200.parallel import broadcast_weights
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    if world > 1:
        sd = synthetic_state_dict(cfg, 0) if rank == 0 else None
        sd = broadcast_weights(cfg, sd, src=0, device=dev, gemm_fp16=(precision == "fast"))
    else:
        sd = synthetic_state_dict(cfg, 0)
    return cls(state_dict=sd, precision=precision)


def _time_single(timer, sampler_cls, model, local, rank, B, L, top_k, burnin, num_positions, W, K, profiler=False):
    """Device-resident Gibbs iterations of a single-sequence config: tokens + schedule in HBM before the clock starts."""
    s = sampler_cls(model, device="cuda:%d" % local, rng="device")
    eng = model.model.engine
    eng.set_stream(timer.stream.cuda_stream)
    toks = model.batch_converter([(str(i), q) for i, q in enumerate(seeds(B, L, 1234 + rank))])[2]
    idx, _ = s.calculate_indexes(None, 0, L, False)
    random.seed(rank)
    plan, _ = s.plan_positions(B, idx, -1, num_positions, False, W + K)
    eng.set_tokens(toks)
    eng.set_schedule(plan.positions, plan.n_iters, plan.P, plan.iter_stride, plan.chain_stride)
    eng.set_noise(None)
    eng.set_device_rng(1234 + rank)
    eng.run(0, W, burnin, top_k, None, True, s.valid_aa_idx)
    launches0 = eng.launch_count()
    ms = timer.run(lambda: eng.run(W, K, burnin, top_k, None, True, s.valid_aa_idx), profiler)
    launches = eng.launch_count() - launches0
    _time_single.kernels = _kernel_shares(eng, lambda: eng.run(W + K - 1, 1, burnin, top_k, None, True, s.valid_aa_idx)) \
        if (rank == 0 and not profiler) else None
    return s, plan, toks, ms, launches


def _kernel_shares(engine, run_one):
    """Share of the step per kernel class: one extra iteration with CUDA events around every launch (rank 0)."""
    engine.profile_enable(True)
    run_one()
    engine.sync()
    prof = engine.profile_read()
    engine.profile_enable(False)
    total = sum(v[0] for v in prof.values()) or 1.0
    return {k: {"share": round(v[0] / total, 4), "avg_launch_us": round(1000.0 * v[0] / v[1], 2), "launches": v[1]}
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}


def _other_config_entry(name, workload, ms, K, flops, peaks, world, clock, kernels=None):
    tf = flops / (ms / K / 1000.0) / 1e12
    return {"workload": workload, "steps": K, "ms_per_step": ms / K, "iters_per_sec": world * K / (ms / 1000.0),
            "n_gpus": world, "algorithmic_tflop_per_iter_per_gpu": flops / 1e12, "achieved_tflops_per_gpu": tf,
            "frac_of_sustained_peak": tf / peaks["tensor"], "clocks": clock, "kernels": kernels}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.config import get_config
    from protein_gibbs_sampler_b200.esm_msa_sampler import ESM_MSA_sampler
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    from protein_gibbs_sampler_b200.parallel import shard_sampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # time on the stream the kernels are launched on: a dedicated torch stream shared with the engines
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    timer = Timer(torch, dist, world, dev, stream)
    peaks = measured_peaks()
    cfg = get_config(MODEL)
    model = _model_on_ranks(models, models.ESM1b, cfg, world, rank, dev)

    B, T = CHAINS_PER_GPU, SEQ_LEN + 2
    K, W = args.steps, args.warmup

    # ---- headline: device-resident throughput of config 2
    clocks = ClockSampler(local).start() if rank == 0 else None
    sampler, plan, tokens, ms, launches = _time_single(timer, ESM_sampler, model, local, rank, B, SEQ_LEN, TOP_K, BURNIN,
                                                       0, W, K, profiler=True)
    clock_info = clocks.stop() if rank == 0 else None
    engine = model.model.engine
    value = world * K / (ms / 1000.0)

    # ---- end to end through the public API: host strings in, host strings out (H2D + D2H inside).  At N > 1 this is
    # the product's sharded path: every rank calls the SAME generate(N*64 chains) on a parallel.shard_sampler, runs
    # its 64-chain slice and the final tokens are all-gathered once.
    all_seeds = seeds(B * world, SEQ_LEN)
    if world > 1:
        shard_sampler(sampler)

    def e2e_generate(n_iters):
        random.seed(0)
        torch.manual_seed(0)
        return sampler.generate(B * world, all_seeds, batch_size=B * world, num_iters=n_iters, top_k=TOP_K, burnin=BURNIN,
                                show_progress_bar=False)

    e2e_generate(min(W, 3))
    timer.barrier()
    t0 = time.perf_counter()
    out = e2e_generate(K)
    torch.cuda.synchronize()
    e2e_s = timer.max_over_ranks(time.perf_counter() - t0)
    assert len(out) == B * world and all(len(s) == SEQ_LEN for s in out)
    e2e_value = world * K / e2e_s
    h2d = (B * T * 4 + plan.P * 4 + len(sampler.valid_aa_idx) * 4) / K   # tokens + schedule + candidate ids, once per call
    d2h = (B * T * 4) / K
    sharded_check = None
    if world > 1:
        # the sharded run returns what ONE GPU computes for all N*64 chains: rank 0 re-runs a short job unsharded
        short = e2e_generate(3)
        sampler.shard = None
        if rank == 0:
            sharded_check = {"iters": 3, "chains": B * world, "equal_to_single_gpu": e2e_generate(3) == short}
            assert sharded_check["equal_to_single_gpu"], "sharded generate differs from the single-GPU run"
        timer.barrier()

    # ---- per-kernel-class timing (separate pass with CUDA events around every launch) for the roofline
    roofline = None
    if rank == 0:
        n_prof = min(K, 5)
        engine.set_tokens(tokens)
        engine.set_schedule(plan.positions, plan.n_iters, plan.P, plan.iter_stride, plan.chain_stride)
        engine.set_chain_offset(0)
        engine.profile_enable(True)
        engine.run(W, n_prof, BURNIN, TOP_K, None, True, sampler.valid_aa_idx)
        engine.sync()
        prof = engine.profile_read()
        engine.profile_enable(False)
        M, d, F = B * T, cfg["embed_dim"], cfg["ffn_dim"]
        gemm_flops = {"gemm_qkv": 2.0 * M * 3 * d * d, "gemm_out": 2.0 * M * d * d, "gemm_fc1": 2.0 * M * d * F,
                      "gemm_fc2": 2.0 * M * d * F, "gemm_head": 2.0 * (B * SEQ_LEN) * d * d,
                      "attention": 4.0 * M * T * d}
        total_ms = sum(v[0] for v in prof.values())
        shares = {k: round(v[0] / total_ms, 4) for k, v in prof.items()}
        per_kernel = {k: {"avg_launch_ms": prof[k][0] / prof[k][1], "launches_per_iter": prof[k][1] // n_prof,
                          "tflops": gemm_flops[k] / (prof[k][0] / prof[k][1] / 1000.0) / 1e12,
                          "frac_of_sustained_peak": gemm_flops[k] / (prof[k][0] / prof[k][1] / 1000.0) / 1e12 / peaks["tensor"]}
                      for k in prof if k in gemm_flops}
        dom = max((k for k in prof if k in gemm_flops), key=lambda k: prof[k][0])
        avg_ms = prof[dom][0] / prof[dom][1]
        achieved = gemm_flops[dom] / (avg_ms / 1000.0) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom)
        step_tf = algorithmic_flops_per_iter(cfg, B, T) * (value / world) / 1e12
        roofline = {"step_frac_of_sustained_peak": step_tf / peaks["tensor"],
                    "bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["tensor"],
                    "unit": "TFLOP/s", "frac": achieved / peaks["tensor"], "traffic": traffic,
                    "peak_source": peaks["source"] + " bf16_tflops_sustained",
                    "avg_launch_ms": avg_ms, "algorithmic_flops_per_launch": gemm_flops[dom],
                    "time_share_by_kernel": shares, "per_kernel": per_kernel,
                    "step": {"algorithmic_tflop_per_iter": algorithmic_flops_per_iter(cfg, B, T) / 1e12,
                             "achieved_tflops": step_tf, "frac_of_sustained_peak": step_tf / peaks["tensor"],
                             "frac_of_burst_peak": step_tf / peaks["tensor_burst"]}}

    # ---- the other BASELINE configs, per-GPU shard each (weak scaling like the headline): 3 warm-up + 5 timed steps
    other = {}
    if not args.no_other_configs:
        OW, OK = 3, 5

        def clocked(fn):
            c = ClockSampler(local).start() if rank == 0 else None
            r = fn()
            return r, (c.stop() if rank == 0 else None)

        # config 5: ESM-1b, 16 of 128 chains x L=1022, num_positions_percent 5 / 10 / 25 (same engine as the headline)
        for pct in (5, 10, 25):
            P = int(1022 * pct / 100)
            (_, _, tk, ms5, _), ck = clocked(lambda: _time_single(timer, ESM_sampler, model, local, rank, 16, 1022, 0,
                                                                   float("inf"), P, OW, OK))
            other["C5_shard_p%d" % pct] = _other_config_entry(
                "C5", "BASELINE configs[4] shard: ESM-1b, 16 of 128 chains x L=1022, num_positions_percent=%d (P=%d)" % (pct, P),
                ms5, OK, algorithmic_flops_per_iter(cfg, 16, 1024), peaks, world, ck, _time_single.kernels)
        engine.close()
        model.model.engine = None
        del model, sampler
        # config 4: ESM-2 650M, 64 of 512 chains x L=512, top_k=5, burnin=50 (at N=8 this IS config 4)
        cfg4 = get_config("esm2_t33_650M_UR50D")
        m4 = _model_on_ranks(models, models.ESM2_t33_650M, cfg4, world, rank, dev)
        (_, _, _, ms4, _), ck = clocked(lambda: _time_single(timer, ESM_sampler, m4, local, rank, 64, 512, 5, 50, 0, OW, OK))
        other["C4_shard"] = _other_config_entry(
            "C4", "BASELINE configs[3] shard: ESM-2 650M, 64 chains/GPU x L=512 (512 chains at N=8), top_k=5, burnin=50, "
                  "all positions", ms4, OK, algorithmic_flops_per_iter(cfg4, 64, 514), peaks, world, ck, _time_single.kernels)
        m4.model.engine.close()
        del m4
        # config 4 again in split-operand precision (logits within 1e-3 of fp32 per row)
        m4s = _model_on_ranks(models, models.ESM2_t33_650M, cfg4, world, rank, dev, precision="split")
        (_, _, _, ms4s, _), ck = clocked(lambda: _time_single(timer, ESM_sampler, m4s, local, rank, 64, 512, 5, 50, 0, OW, 3))
        other["C4_shard_split_precision"] = _other_config_entry(
            "C4", "as C4_shard with precision='split' (fp16 hi+lo operands, three tensor-core passes per GEMM)", ms4s, 3,
            algorithmic_flops_per_iter(cfg4, 64, 514), peaks, world, ck, _time_single.kernels)
        m4s.model.engine.close()
        del m4s
        # ... and with only the weights carried as hi + lo pairs (two passes per GEMM: inside 1e-3 on the batch metric)
        m4w = _model_on_ranks(models, models.ESM2_t33_650M, cfg4, world, rank, dev, precision="split_weights")
        (_, _, _, ms4w, _), ck = clocked(lambda: _time_single(timer, ESM_sampler, m4w, local, rank, 64, 512, 5, 50, 0, OW, 3))
        other["C4_shard_split_weights"] = _other_config_entry(
            "C4", "as C4_shard with precision='split_weights' (fp16 hi+lo weights, two tensor-core passes per GEMM)", ms4w, 3,
            algorithmic_flops_per_iter(cfg4, 64, 514), peaks, world, ck, _time_single.kernels)
        m4w.model.engine.close()
        del m4w
        # config 3: MSA-1b, 16 MSAs x 32 rows x L=128, 10 % of the positions of every row per iteration
        cfg3 = get_config("esm_msa1b_t12_100M_UR50S")
        m3 = _model_on_ranks(models, models.ESM_MSA1, cfg3, world, rank, dev)
        s3 = ESM_MSA_sampler(m3, device="cuda:%d" % local, rng="device")
        e3 = m3.model.engine
        e3.set_stream(stream.cuda_stream)
        rows = seeds(32, 128, 99 + rank)
        toks3 = m3.batch_converter([[(str(i), q) for i, q in enumerate(rows)]] * 16)[2]
        idx3, _ = s3.calculate_indexes(None, 0, 128, False)
        for label, P3 in (("C3", 12), ("C3_all_positions", 0)):
            random.seed(rank)
            plan3, _ = s3.plan_positions(16, 32, idx3, -1, P3, False, OW + OK)
            e3.set_tokens(toks3)
            e3.set_schedule(plan3.positions, plan3.n_iters, plan3.P, plan3.iter_stride, plan3.chain_stride)
            e3.set_noise(None)
            e3.set_device_rng(7 + rank)
            e3.run(0, OW, float("inf"), 0, None, True, s3.valid_aa_idx)
            ms3, ck = clocked(lambda: timer.run(lambda: e3.run(OW, OK, float("inf"), 0, None, True, s3.valid_aa_idx)))
            k3 = _kernel_shares(e3, lambda: e3.run(OW + OK - 1, 1, float("inf"), 0, None, True, s3.valid_aa_idx)) if rank == 0 else None
            other[label] = _other_config_entry(
                "C3", "BASELINE configs[2]: MSA-1b, 16 MSAs/GPU x 32 rows x L=128, %s positions per row and iteration"
                      % ("10 %% (P=12)" if P3 else "all 128"), ms3, OK, msa_flops_per_iter(cfg3, 16, 32, 129), peaks, world, ck, k3)
        e3.close()

    cpu = cpu_baseline_sample() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate/residual/softmax/LayerNorm", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "one ESM_sampler.generate call of K iterations: host strings -> tokens -> H2D -> K "
                            "on-device iterations -> D2H -> strings; bytes are per-call totals / K"
                            + ("; sharded over the ranks by parallel.shard_sampler (every rank calls generate for all "
                               "N*64 chains, runs its slice, final tokens all-gathered once)" if world > 1 else ""),
                    "sharded_equals_single_gpu": sharded_check},
            "gpu_launches": launches, "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu,
            "other_configs": other,
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
    ap.add_argument("--no-other-configs", action="store_true")
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
