"""N>1 host logic on CPU: two gloo ranks shard a photon file by chunk, bin their shards (with the
oracle standing in for the GPU kernel) and all-reduce; the result equals the single-process oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mkids_sdr_b200 import dist as mdist
    from mkids_sdr_b200 import synth
    from oracle import decode as odec
    R, npix, secs, cap = 4, 253, 3, 2500
    streams, eos = synth.photon_streams(300000, R, npix, secs, seed=5, n_hot=1, hot_rate=3000)
    assert all(np.array_equal(mdist.find_eos(s), e) for s, e in zip(streams, eos))
    shards = mdist.shard_streams([len(s) for s in streams], eos, world, chunk_words=8192 * 2)
    n_pix = R * npix
    counts = torch.zeros(secs * n_pix, dtype=torch.int64)
    hist = torch.zeros(n_pix * 16, dtype=torch.int64)
    lut = np.arange(4096) * 16 // 4096
    for (r, start, length, sec0) in shards[rank]:
        piece = streams[r][start:start + length]
        # a chunk starting after `sec0` closed seconds: prepend nothing, shift the second index
        adr = (piece >> np.uint64(56)).astype(np.int64)
        is_eos = adr == 255
        sec = sec0 + np.cumsum(is_eos) - is_eos
        ok = (sec < secs) & ~is_eos & (adr < npix)
        counts += torch.from_numpy(np.bincount(sec[ok] * n_pix + r * npix + adr[ok], minlength=secs * n_pix))
        peak = ((piece >> np.uint64(44)) & np.uint64(0xFFF)).astype(np.int64)
        hist += torch.from_numpy(np.bincount((r * npix + adr[ok]) * 16 + lut[peak[ok]], minlength=n_pix * 16))
    mdist.reduce_products([counts, hist])
    if rank == 0:
        ref = odec.packetmaster_bin(streams, npix, secs, cap)
        ok1 = np.array_equal(counts.numpy().reshape(secs, n_pix), ref['raw_counts'])
        ok2 = np.array_equal(hist.numpy().reshape(n_pix, 16), odec.pixel_field_hist(streams, npix, secs, 'peak', lut, 16))
        capped = np.minimum(counts.numpy(), cap - 1).reshape(secs, n_pix)
        ok3 = np.array_equal(capped, ref['counts'])
        q.put((ok1, ok2, ok3, len(shards[0]), len(shards[1])))
    dist.destroy_process_group()


def test_two_rank_chunk_sharding_and_reduce():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] and res[1] and res[2], res
    assert res[3] > 0 and res[4] > 0


def _lut_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mkids_sdr_b200 import dist as mdist
    from oracle import lut as olut
    n_boards, N, T, FS = 5, 2 ** 10, 8, 512e6

    def board_tables(b):                     # the oracle stands in for mkid_comb_lut: one LUT set per board
        rng = np.random.default_rng(100 + b)
        k = np.sort(rng.choice(np.arange(-N // 2 + 1, N // 2), T, replace=False))
        I, Q, sc, _ = olut.freq_comb_lut('yes', list((k % N) * FS / N), FS, FS / N, list(rng.uniform(0.1, 1.0, T)))
        return np.concatenate([I, Q]).astype(np.int64)
    sums = torch.zeros(n_boards * 2, dtype=torch.int64)       # per board: sum and a position-weighted sum of its tables
    for b in mdist.assign_boards(n_boards, world, rank):
        t = board_tables(b)
        sums[2 * b] = int(t.sum()); sums[2 * b + 1] = int((t * np.arange(1, t.size + 1)).sum())
    mdist.reduce_products([sums])            # LUT synthesis has no data-path collective; this only gathers the check
    if rank == 0:
        ref = []
        for b in range(n_boards):
            t = board_tables(b)
            ref += [int(t.sum()), int((t * np.arange(1, t.size + 1)).sum())]
        q.put(sums.tolist() == ref)
    dist.destroy_process_group()


def test_two_rank_lut_board_sharding():
    """LUT sets shard by board with no collective; every board is synthesised by exactly one rank."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_lut_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_assign_boards():
    from mkids_sdr_b200 import dist as mdist
    for total, world in ((8, 1), (8, 2), (8, 8), (80, 8), (10, 4)):
        got = [mdist.assign_boards(total, world, r) for r in range(world)]
        assert sorted(sum(got, [])) == list(range(total))
        assert max(len(g) for g in got) - min(len(g) for g in got) <= 1
