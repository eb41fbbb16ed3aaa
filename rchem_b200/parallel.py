"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch;
gloo for the CPU tests).  The J/K build has exactly one exchange step: every rank digests its
share of the screened quartet list into private partial J and K, and the partials are summed
with ONE all-reduce over the contiguous [J | K] buffer.  The dense tensor
needs no collective (each rank could keep its own slab)."""
import os


def init_distributed(device_backend="nccl"):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun) and joins the process group.
    Returns (rank, world_size, local_rank)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if device_backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(device_backend, rank=rank, world_size=world)
    return rank, world, local


def block_owner(block, nranks):
    """Thread blocks (4 warps = 128 shell quartets of one bra-ket task) are dealt round-robin:
    rank r runs blocks b with b % nranks == r (eri_kernel.cuh, EriTask::rank/nranks)."""
    return block % nranks


def blocks_of_rank(nblocks, rank, nranks):
    """How many of nblocks thread blocks rank `rank` launches (engine.cu run_tasks)."""
    return (nblocks - rank + nranks - 1) // nranks if nblocks > rank else 0


def allreduce_jk(jk_tensor, group=None):
    """Sum the per-rank partial [J | K] buffers in place (one collective per J/K build)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(jk_tensor, op=dist.ReduceOp.SUM, group=group)
    return jk_tensor


def jk_direct_distributed(basis, D_dev, JK_dev, rank, nranks, group=None):
    """This rank's share of JK_direct on device tensors (torch, float64, cuda), then the
    all-reduce.  D_dev: (N,N); JK_dev: (2,N,N) receives J and K."""
    import torch

    # the build and the collective must be ordered on ONE stream: torch's current stream (it
    # produced D_dev, and NCCL's all-reduce is enqueued behind it)
    basis.set_stream(torch.cuda.current_stream(D_dev.device).cuda_stream)
    basis.jk_direct_device(D_dev.data_ptr(), JK_dev.data_ptr(), rank, nranks)
    return allreduce_jk(JK_dev, group)
