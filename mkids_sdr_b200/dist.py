"""Multi-GPU host logic: static sharding of the work units and the single reduction of the
per-pixel products.  One process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests).

  * channelize -> detect : shard by board / feedline (independent streams, no collective)
  * decode / histogram    : shard by packet-file chunk (bundle-aligned ranges of each roach stream);
                            a chunk carries the number of seconds closed before it, taken from the
                            positions of the end-of-second words (PacketMaster.c:329-368)
  * one all-reduce (sum) of counts_raw[sec][pixel] and hist[pixel][bin]; the 2500-event cap quirk
    (PacketMaster.c:373-380) is applied after the reduction.
"""
import numpy as np

BUNDLE = 8192


def assign_boards(n_boards_total, world, rank):
    """Boards of this rank: contiguous blocks, remainder spread over the first ranks."""
    base, rem = divmod(n_boards_total, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def shard_streams(stream_lens, eos_positions, world, chunk_words=BUNDLE * 16):
    """Split every roach stream into bundle-aligned chunks and deal them round-robin to the ranks.

    stream_lens[r]: words in roach r's stream; eos_positions[r]: sorted indices of its EOS words.
    Returns per rank a list of (roach, start, length, sec_start)."""
    assert chunk_words % BUNDLE == 0
    out = [[] for _ in range(world)]
    k = 0
    for r, n in enumerate(stream_lens):
        eos = np.asarray(eos_positions[r], dtype=np.int64)
        for start in range(0, n, chunk_words):
            length = min(chunk_words, n - start)
            sec_start = int(np.searchsorted(eos, start, side='left'))     # EOS words strictly before `start`
            out[k % world].append((r, start, length, sec_start))
            k += 1
    return out


def find_eos(words):
    """Indices of end-of-second words (channel field 255) in a host stream."""
    w = np.asarray(words, dtype=np.uint64)
    return np.nonzero((w >> np.uint64(56)) == np.uint64(255))[0]


def reduce_products(tensors, group=None):
    """Sum the per-pixel products over all ranks in place (torch tensors, int32/int64)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


class ProductReducer:
    """The one collective of the path through the C ABI (mkid_hist_allreduce / mkid_hist_reduce): an NCCL communicator
    owned by the library, its unique id handed from rank 0 to the others through the existing torch.distributed group
    (any backend) or through any callable `bcast(bytes_or_None) -> bytes`.  The reduce is queued on the context's own
    stream: decode -> reduce -> next batch need no host synchronisation."""

    def __init__(self, ctx, rank, world, bcast=None):
        import ctypes
        from . import _lib
        self.ctx, self.rank, self.world = ctx, rank, world
        self.comm = ctypes.c_void_p()
        if world == 1:
            return
        uid = (ctypes.c_uint8 * 128)()
        if rank == 0:
            ctx._check(ctx.lib.mkid_nccl_unique_id(ctx.h, uid))
        if bcast is None:
            import torch
            import torch.distributed as dist
            dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
            t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
            dist.broadcast(t, 0)
            data = bytes(t.cpu().tolist())
        else:
            data = bcast(bytes(uid) if rank == 0 else None)
        uid = (ctypes.c_uint8 * 128).from_buffer_copy(data)
        ctx._check(ctx.lib.mkid_nccl_init(ctx.h, uid, world, rank, ctypes.byref(self.comm)))

    def allreduce(self, buf, n):
        """In-place sum over all ranks of n uint32 values at device address `buf` (DeviceBuffer / tensor / int)."""
        if self.world == 1:
            return
        from . import _lib
        c = self.ctx
        c._check(c.lib.mkid_hist_allreduce(c.h, self.comm, _lib.ptr(buf), int(n)))

    def reduce(self, buf, n, root=0):
        if self.world == 1:
            return
        from . import _lib
        c = self.ctx
        c._check(c.lib.mkid_hist_reduce(c.h, self.comm, _lib.ptr(buf), int(n), int(root)))

    def close(self):
        if self.comm:
            self.ctx.lib.mkid_nccl_destroy(self.ctx.h, self.comm)
            self.comm = None
