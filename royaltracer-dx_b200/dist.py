"""Multi-GPU plumbing: the scene and its BVH are replicated on every GPU, samples are partitioned across
ranks, and the accumulation buffers (gPermanentData, float4 per pixel) are combined with ONE reduce per progressive pass.
The path shards with no data-path collective other than that reduce.  On the GPU the reduce lives behind the C ABI
(rtx_comm_init / rtx_reduce_accum: ncclReduce straight from gPermanentData on a side stream, overlapped with the next pass);
torch.distributed only carries the 128-byte NCCL id between the ranks.  reduce_accum() below is the same plumbing on torch
tensors for the CPU tests (gloo)."""


def init_engine_comm(ctx, rank, world):
    """Collective: rank 0 draws the NCCL unique id through the engine, every rank joins the engine's communicator.
    Needs an initialised torch.distributed process group (any backend) for the broadcast of the id."""
    import torch.distributed as dist
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)


def sample_for(step, rank, world):
    """Global sample index rendered by `rank` in progressive pass `step`: s = step * world + rank, i.e. rank r renders
    every sample with s mod world == r.  Seeds depend on the global sample index only (shaders/Pass_init_di_v7.hlsl:76-77
    with uint(time) := s), so the image is independent of the GPU count up to fp32 summation order."""
    return step * world + rank


def samples_of_rank(n_samples, rank, world):
    return list(range(rank, n_samples, world))


def reduce_accum(local_accum, scratch, dst=0):
    """One reduce per pass.  `local_accum` stays this rank's private partial sum (a later pass keeps adding to it);
    the global sum lands in `scratch` on rank dst.  Both are float32 tensors of shape (H, W, 4): rgb sums + sample count."""
    import torch.distributed as dist
    scratch.copy_(local_accum)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(scratch, dst=dst, op=dist.ReduceOp.SUM)
    return scratch


def wrap_device_buffer(ptr, shape, typestr="<f4"):
    """A torch view of device memory owned by the engine (no copy), through __cuda_array_interface__."""
    import torch

    class _W:
        pass
    w = _W()
    w.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(w, device="cuda")
