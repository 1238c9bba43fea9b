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
