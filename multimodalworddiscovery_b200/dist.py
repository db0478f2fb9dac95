"""Cross-rank reduction of the packed count buffers (the only collective on the path).

Each EM iteration every rank contributes ONE packed float64 buffer
[translation counts | init counts | transition counts | sum log-lik | posterior gradient]
(~0.3 MB at K=65, P=49, D=512).  It is combined by all_gather + a fixed-rank-order reduction, so
the result is bitwise independent of NCCL's internal reduction order (and identical on all
ranks); at this size the collective is latency-bound, so the extra bytes over all_reduce are free.
"""


def is_distributed(group=None):
    try:
        import torch.distributed as dist
    except ImportError:
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def fixed_order_allreduce(buf, group=None, log_domain=False, ll_index=None):
    """In-place sum (or log-sum-exp) of ``buf`` over ranks in rank order.

    ``ll_index``: position of a plain-sum entry inside a log-domain buffer (the log-likelihood)."""
    if not is_distributed(group):
        return buf
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    flat = torch.empty((world * buf.numel(),), dtype=buf.dtype, device=buf.device)
    dist.all_gather_into_tensor(flat, buf.reshape(-1), group=group)
    if buf.is_cuda:
        # product path: one kernel of libmwd_b200.so, rank order fixed
        import ctypes as C
        from . import _lib
        st = C.c_void_p(torch.cuda.current_stream(buf.device).cuda_stream)
        _lib.check(_lib.load().mwd_rank_reduce(C.c_void_p(flat.data_ptr()), world, buf.numel(), 1 if log_domain else 0,
                                               -1 if ll_index is None else int(ll_index),
                                               C.c_void_p(buf.data_ptr()), st))
        return buf
    # CPU tensors: only the gloo host-logic tests (tests/test_dist_gloo.py) come through here
    g = flat.view(world, buf.numel())
    if log_domain:
        ll = g[:, ll_index].sum() if ll_index is not None else None
        torch.logsumexp(g, dim=0, out=buf.reshape(-1))
        if ll is not None:
            buf.reshape(-1)[ll_index] = ll
    else:
        torch.sum(g, dim=0, out=buf.reshape(-1))
    return buf
