"""Scene-batch sharding over the GPUs of one box (one process per GPU).

The backbone forward has no cross-frame dependency in eval mode (BN uses
running statistics, LN is per token, attention per window), so the path shards
by scene batch with no data-path collective.  The only
communication is the gathering of results / timings after the forward, the
analogue of the reference's `collect_results_gpu` (mmdet/apis/test.py:278-308).
Works with the `nccl` backend on GPUs and `gloo` on CPU (tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; no-op for 1 process.
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def _cpu_ranges(cpus):
    cpus = sorted(cpus)
    out, i = [], 0
    while i < len(cpus):
        j = i
        while j + 1 < len(cpus) and cpus[j + 1] == cpus[j] + 1:
            j += 1
        out.append(str(cpus[i]) if i == j else f'{cpus[i]}-{cpus[j]}')
        i = j + 1
    return ','.join(out)


def bind_to_gpu_numa(local_rank):
    """Restrict this process to the CPUs NVML reports as local to GPU `local_rank` (its NUMA
    node), so the pinned host buffers it allocates afterwards are first-touched next to the
    GPU's PCIe root and the H2D / D2H copies of the end-to-end path do not cross the socket
    interconnect.  With 8 ranks each moving ~90 MB per 3 ms step that link is what saturates
    first.  Call it before allocating pinned memory.  Returns a description of what was
    done ('cpus 0-55' / the reason nothing was); never raises."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = None
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (max(os.cpu_count() or 64, 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        near = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = near & allowed
        if not cpus:
            return f'unbound (GPU-local cpus {_cpu_ranges(near)} not in this cpuset)'
        if cpus == allowed:
            return f'cpus {_cpu_ranges(cpus)} (whole cpuset is GPU-local)'
        os.sched_setaffinity(0, cpus)
        return f'cpus {_cpu_ranges(cpus)}'
    except Exception as e:      # no NVML, no permission: run unbound
        return f'unbound ({type(e).__name__}: {e})'


def shard_bounds(n_frames, rank, world):
    """Contiguous, balanced [lo, hi) slice of a global batch for `rank`."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank, world):
    """Slice every tensor's leading (frame) dimension for this rank."""
    lo, hi = shard_bounds(tensors[0].shape[0], rank, world)
    return [t[lo:hi] for t in tensors]


def gather_frames(t, n_frames=None):
    """all_gather per-rank results (B_r, ...) back into global frame order.
    Ranks may hold different B_r; pads to the max and trims."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    world = dist.get_world_size()
    counts = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    all_counts = [int(c) for c in all_counts]
    mx = max(all_counts)
    if t.shape[0] < mx:
        t = torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))])
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return torch.cat([p[:c] for p, c in zip(parts, all_counts)])


def max_over_ranks(value, device):
    """max of a python float over all ranks (timing rule: slowest rank counts)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
