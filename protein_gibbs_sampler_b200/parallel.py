"""Multi-GPU plumbing: chains (and whole MSAs) are independent Markov chains, so they shard across ranks with
no data-path collective.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only to ship the weights
from rank 0 once and to gather the final token tensors."""
import torch
import torch.distributed as dist


def shard_range(n_items, world_size, rank):
    """Contiguous split of ``n_items`` chains (or MSAs -- never split one MSA's rows) over ranks."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_state_dict(sd, src=0, device="cpu"):
    """Rank ``src`` holds ``sd``; every rank returns the same dict of fp32 tensors on ``device``."""
    rank = dist.get_rank()
    meta = [{k: tuple(v.shape) for k, v in sd.items()} if rank == src else None]
    dist.broadcast_object_list(meta, src=src)
    out = {}
    for k, shape in meta[0].items():
        t = sd[k].to(device=device, dtype=torch.float32).contiguous() if rank == src else \
            torch.empty(shape, dtype=torch.float32, device=device)
        dist.broadcast(t, src=src)
        out[k] = t
    return out


def _blob_plan(cfg, gemm_fp16):
    from .weights import is_gemm_weight, state_dict_layout
    plan, off = [], 0
    for key, shape in state_dict_layout(cfg):
        numel = 1
        for n in shape:
            numel *= n
        half = gemm_fp16 and is_gemm_weight(key)
        nbytes = numel * (2 if half else 4)
        plan.append((key, shape, half, off, nbytes))
        off += (nbytes + 15) // 16 * 16          # keep every tensor 16-byte aligned inside the blob
    return plan, off


def pack_weights(cfg, sd, gemm_fp16=True):
    """One contiguous byte blob holding the whole model: the GEMM weights as fp16 (what the tensor cores consume --
    the engine would round them to exactly these values anyway) when ``gemm_fp16``, everything else fp32."""
    plan, total = _blob_plan(cfg, gemm_fp16)
    blob = torch.zeros(total, dtype=torch.uint8)
    for key, shape, half, off, nbytes in plan:
        t = sd[key].detach().to("cpu", torch.float16 if half else torch.float32).contiguous()
        assert tuple(t.shape) == tuple(shape), (key, tuple(t.shape), shape)
        blob[off:off + nbytes] = t.view(-1).view(torch.uint8)
    return blob


def unpack_weights(cfg, blob, gemm_fp16=True):
    """fp32 state dict (on the blob's device) out of a packed blob."""
    plan, total = _blob_plan(cfg, gemm_fp16)
    assert blob.numel() == total and blob.dtype == torch.uint8
    out = {}
    for key, shape, half, off, nbytes in plan:
        raw = blob[off:off + nbytes]
        out[key] = (raw.view(torch.float16).to(torch.float32) if half else raw.view(torch.float32)).reshape(shape)
    return out


def broadcast_weights(cfg, sd, src=0, device="cpu", gemm_fp16=True):
    """The ONE collective of a sharded job: rank ``src`` packs the model into a single blob (1.3 GB for the 650M
    models with fp16 GEMM weights), one ``dist.broadcast`` ships it over NCCL / NVLink, and every rank -- ``src``
    included, so that all ranks run bit-identical weights -- unpacks the same bytes.  The layout comes from ``cfg``
    alone, so no metadata is exchanged.  Pass ``gemm_fp16=False`` for split-operand precision (the lo halves need the
    fp32 weights)."""
    _, total = _blob_plan(cfg, gemm_fp16)
    if dist.get_rank() == src:
        blob = pack_weights(cfg, sd, gemm_fp16).to(device)
    else:
        blob = torch.empty(total, dtype=torch.uint8, device=device)
    dist.broadcast(blob, src=src)
    return unpack_weights(cfg, blob, gemm_fp16)


def gather_sequences(local_seqs):
    """All ranks' output strings in rank order (rank 0 gets the full list, others too)."""
    bucket = [None] * dist.get_world_size()
    dist.all_gather_object(bucket, list(local_seqs))
    return [s for part in bucket for s in part]


def all_gather_tokens(local):
    """List of every rank's int64 token tensor [n_r, R, T] in rank order (ranks may hold different n_r, even 0)."""
    bucket = [None] * dist.get_world_size()
    dist.all_gather_object(bucket, local.cpu())
    return bucket


def shard_sampler(sampler, rank=None, world_size=None, all_gather=all_gather_tokens):
    """Make ``sampler.generate`` run its chains sharded over the ranks of the default process group.

    Every rank must call ``generate`` with the same arguments and the same ``random`` / ``torch`` RNG state (seed them
    alike, or broadcast rank 0's state): each pre-draws the full position schedule (and, in replay mode, the Exp(1)
    variates) in the reference's order, runs only its contiguous slice of chains on its own GPU -- no collective on the
    iteration path -- and the final tokens are all-gathered once, so every rank returns the sequences a single GPU
    would have produced.  ``all_gather`` is injectable for tests."""
    if rank is None:
        rank, world_size = dist.get_rank(), dist.get_world_size()
    sampler.shard = (rank, world_size, all_gather)
    return sampler
