#!/usr/bin/env python
"""Benchmark of the B200-native MKID readout hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the full chain (channelize -> phase -> detect -> photon words ->
decode/bin/histogram [-> NCCL sum of the per-pixel histograms when N > 1]) over one batch of
synthetic ADC samples: 8 boards x 256 channels per GPU (BASELINE config "full ARCONS chain ...
8 boards x 256 channels", weak scaling: every GPU runs its own 8 boards).

Rank 0 prints ONE JSON line (see the contract in the task statement / DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 512e6
N_LUT = 2 ** 19
BOARDS_PER_GPU = 8
N_ACTIVE = 253
METRIC = 'channelized_adc_samples_per_s_full_chain'
UNIT = 'MS/s'
BYTES_PER_SAMPLE = 4.0          # algorithmic: one complex int16 ADC sample read


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons while the timed region runs: NVML every 5 ms (nvidia-smi, ~10 Hz, as fallback)."""

    REASONS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown'}

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
            phys = index
            if vis and all(v.strip().isdigit() for v in vis.split(',')):
                ids = [int(v) for v in vis.split(',')]
                phys = ids[index] if index < len(ids) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is not None:
            nv = self.nvml
            while not self.stop_flag:
                try:
                    sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                    try:
                        mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    except Exception:
                        mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                    self.rows.append((sm, self.max_sm, mask))
                except Exception:
                    pass
                time.sleep(0.005)
            return
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown')
        bits = [0x4, 0x8, 0x20, 0x40]
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                r = [x.strip() for x in out.strip().split(',')]
                mask = sum(b for b, v in zip(bits, r[2:6]) if v.lower().startswith('active'))
                self.rows.append((float(r[0]), float(r[1]), mask))
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = max([r[1] for r in self.rows], default=0)
        reasons = set()
        for r in self.rows:
            for bit, name in self.REASONS.items():
                if r[2] & bit:
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons),
                'samples': len(sm), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


# ------------------------------------------------------------------ NUMA placement of the rank
def numa_bind(torch, local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are allocated
    (first touch places them on that node): the end-to-end leg of 8 ranks then uploads from both sockets' memory."""
    try:
        prop = torch.cuda.get_device_properties(local_rank)
        bus = '%04x:%02x:%02x.0' % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read().strip())
        if node < 0:
            return {'node': None}
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {'node': node, 'cpus': len(allowed)}
    except Exception as e:          # placement is an optimisation only
        return {'node': None, 'error': str(e)[:80]}


# ------------------------------------------------------------------ one chain leg (device resident)
class ChainLeg:
    """The full chain over `B` boards of this rank: products (counts [exptime][n_pix] + hist [n_pix][bins]) in ONE device
    tensor, summed over the ranks by ONE NCCL all-reduce queued on the context stream (mkid_hist_allreduce)."""

    def __init__(self, torch, ctx, reducer, B, n, n_lut, n_active, npix_per_roach, n_roaches_total, roach0, seed0, synth_seed,
                 hist_bins, exptime, want_merged=True, pipelined=False):
        from mkids_sdr_b200.chain import ReadoutChain
        from mkids_sdr_b200.channelizer import synth_adc
        self.torch, self.ctx, self.reducer, self.B, self.n = torch, ctx, reducer, B, n
        self.n_pix = n_roaches_total * npix_per_roach
        self.n_counts, self.n_hist = exptime * self.n_pix, self.n_pix * hist_bins
        self.products = torch.zeros(self.n_counts + self.n_hist, dtype=torch.int32, device='cuda')
        torch.cuda.synchronize()
        self.chain, self.boards = ReadoutChain.synthetic(
            B, n_lut, n_active, seed0=seed0, ctx=ctx, exptime=exptime, n_roaches_total=n_roaches_total, roach0=roach0,
            n_bins=hist_bins, npix_per_roach=npix_per_roach, counts_buf=self.products[:self.n_counts],
            hist_buf=self.products[self.n_counts:], want_merged=want_merged, pipelined=pipelined)
        self.thr = self.chain.derive_thresholds(self.boards)
        tone_bins = np.stack([bd['tone_bins'] for bd in self.boards])
        self.iq = torch.empty((B, n, 2), dtype=torch.int16, device='cuda')
        synth_adc(B, n, tone_bins, n_lut=n_lut, pulse_rate=1000.0, seed=synth_seed, out=self.iq, ctx=ctx)
        ctx.sync()

    def run(self, k):
        """k batches queued back to back (no host round trip: word counts and carried seconds stay on the device), then
        the one reduce of the products, all on the context stream."""
        for _ in range(k):
            self.chain.process_async(self.iq, n=self.n)
        self.chain.join()                # (pipelined chain: the products are written on its second stream)
        self.ctx.record(4)
        self.reducer.allreduce(self.products, self.n_counts + self.n_hist)
        self.ctx.record(5)

    def timed(self, steps, warmup, barrier, world, dist):
        torch, ctx = self.torch, self.ctx
        self.run(max(warmup, 1))
        self.chain.sync_state()
        barrier()
        l0 = self.chain.launches
        ctx.record(0)
        t0 = time.time()
        self.run(steps)
        ctx.record(1)
        self.chain.sync_state()
        barrier()
        wall = (time.time() - t0) * 1e3
        kk = max(1, min(steps, 64))
        k4_ms = self.chain.chan.kernel_ms_sum(kk) / kk
        el = torch.tensor([max(ctx.elapsed_ms(0, 1), 0.0), wall, ctx.elapsed_ms(4, 5)], dtype=torch.float64, device='cuda')
        if dist is not None:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, red_ms = float(el[0]), float(el[1]), float(el[2])
        step_ms = max(dev_ms, wall_ms) / steps           # device events; the wall clock is the cross-check
        return dict(step_ms=step_ms, value=world * self.B * self.n / (step_ms * 1e-3) / 1e6, k4_ms=k4_ms,
                    launches=int(self.chain.launches - l0), reduce_ms=red_ms, reduce_bytes=int((self.n_counts + self.n_hist) * 4),
                    dev_ms=dev_ms, wall_ms=wall_ms)

    def free(self):
        self.iq = self.products = None
        self.chain = None
        self.torch.cuda.empty_cache()


def reference_arm(args, cores, config):
    """`--impl reference`: the oracle port of the chain on all host cores, CPU only (no GPU, no libmkidgpu.so).  A step =
    every core channelizes, detects and bins its resident 2^k-sample board stream (a bounded sample of the workload)."""
    from oracle import cpu_arm
    log2_each = 21
    while log2_each > 18 and 1.1 * 2.0 ** (log2_each - 21) * (args.warmup + args.steps + 1) > 150.0:
        log2_each -= 1
    arm = cpu_arm.CpuArm(cores, 1 << log2_each, cpu_arm.fir_int_default())
    for _ in range(args.warmup):
        arm.step()
    secs = [arm.step() for _ in range(args.steps)]
    arm.close()
    dt = float(np.mean(secs))
    v = cores * (1 << log2_each) / dt / 1e6
    desc = arm.describe()
    return {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': max(args.gpus, 1), 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': config,
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc,
                             'setup_seconds_untimed': arm.setup_seconds},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}


def workload_config(n, hist_bins):
    return {'workload': 'full ARCONS chain with matched-filter pulse detection: 8 boards x 256 channels per GPU (253 driven), '
                        'channelize->phase->detect->photon words->decode/bin/hist + merged time-ordered list',
            'boards_per_gpu': BOARDS_PER_GPU, 'samples_per_board_per_step': n, 'n_lut': N_LUT, 'fir': 'matched_30us',
            'pulse_rate_hz': 1000, 'hist_bins': hist_bins, 'hist_field': 'peak',
            'l2': 'inputs (%.0f MiB per GPU per step) are larger than the 126 MB L2' % (BOARDS_PER_GPU * n * 4 / 2 ** 20),
            'sharding': 'boards per GPU, no data-path collective; ONE NCCL all-reduce of the per-pixel products per job',
            'streams': 'two contexts per GPU: detection / decode / merged list of batch k under the channelizer kernel of batch k + 1'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--log2-samples', type=int, default=25, help='ADC samples per board per step (log2)')
    ap.add_argument('--hist-bins', type=int, default=4096)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the strong-scaling / stress / decode / LUT side measurements')
    ap.add_argument('--single-stream', action='store_true', help='headline leg on ONE stream (no overlap of the detection tail with the next channelizer kernel)')
    ap.add_argument('--stress', action='store_true', help='run the config-5 stress leg at any GPU count (default: 8 GPUs only)')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    n_gpus = max(args.gpus, 1)
    if args.impl == 'reference' and rank != 0:
        return 0

    # exactly ONE line goes to stdout (the JSON): libraries that write there (NCCL's version banner) land on stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + '\n').encode())

    t_start = time.time()

    def stage(msg):
        if os.environ.get('MKID_BENCH_VERBOSE'):
            sys.stderr.write('[bench rank %d +%.1fs] %s\n' % (rank, time.time() - t_start, msg)); sys.stderr.flush()

    n = 1 << args.log2_samples
    B = BOARDS_PER_GPU
    config = workload_config(n, args.hist_bins)
    cores = len(os.sched_getaffinity(0))
    if args.impl == 'reference':
        emit(reference_arm(args, cores, config))
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    numa = numa_bind(torch, local_rank) if world > 1 else {'node': None}
    stage('process group up, numa %s' % numa)
    from mkids_sdr_b200 import _lib
    from mkids_sdr_b200.dist import ProductReducer
    ctx = _lib.default_context(local_rank)
    reducer = ProductReducer(ctx, rank, world)           # the library's own NCCL communicator (C ABI)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    exptime = 16
    # ---------------------------------------------------------------- headline: weak scaling, 8 boards per GPU
    # two contexts (streams) per GPU: resolve / emit / decode / merged list of batch k run under the channelizer kernel of
    # batch k + 1 (measured on one GPU: 1.42 -> 1.31 ms per step; the kernel itself stretches from 1.266 to 1.293 ms because
    # the two compete for SMs when both become ready)
    leg = ChainLeg(torch, ctx, reducer, B, n, N_LUT, N_ACTIVE, 253, B * world, B * rank, 42 + 8 * rank, 1000 + rank,
                   args.hist_bins, exptime, pipelined=not args.single_stream)
    stage('chain configured, thresholds derived, input synthesised')
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = leg.timed(args.steps, args.warmup, barrier, world, dist)
    stage('device-resident steps done')
    chain = leg.chain
    n_pix = leg.n_pix

    # ---------------------------------------------------------------- end to end with host buffers
    # Two host formats of the same samples: the 12-bit packed stream (3 bytes per complex sample; the ADC is 12 bit and the
    # synthetic input is clipped to it, so the photon words are identical: tests/test_chain_gpu.py) is the headline `e2e`,
    # the int16 pairs (4 bytes per sample) are reported next to it as `e2e_int16`.
    cap = chain.chan.words_capacity(n)
    pin_words = torch.empty((B, cap), dtype=torch.int64).pin_memory()
    pin_counts = torch.empty(exptime * n_pix, dtype=torch.int32).pin_memory()
    words_np = pin_words.numpy().view(np.uint64)
    k_e2e = max(min(args.steps // 2, 25), 1)

    def e2e_run(pin_iq, fmt):
        """k batches from pinned host memory through ReadoutChain.process_stream: H2D of every batch (double buffered
        on a second stream, overlapping the previous batch's kernels), D2H of the photon words, word counts and
        per-pixel counts after every batch."""
        def steps(k):
            nw = 0
            for nwk in chain.process_stream((pin_iq for _ in range(k)), n, words_host=words_np, counts_host=pin_counts,
                                            adc_format=fmt):
                nw += int(nwk.sum())
            return nw
        steps(1)
        barrier()
        t0 = time.time()
        nw = steps(k_e2e)
        reducer.allreduce(leg.products, leg.n_counts + leg.n_hist)
        barrier()
        el = torch.tensor([(time.time() - t0) * 1e3], dtype=torch.float64, device='cuda')
        if dist is not None:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        step_ms = float(el[0]) / k_e2e
        return step_ms, world * B * n / (step_ms * 1e-3) / 1e6, nw / k_e2e

    pin_iq = torch.empty((B, n, 2), dtype=torch.int16).pin_memory()
    pin_iq.copy_(leg.iq.cpu())
    e2e16_step_ms, e2e16_value, words_per_step = e2e_run(pin_iq, 'i16')
    del pin_iq
    packed_dev = ctx.alloc(B * n * 3)
    n_clipped = ctx.adc_pack12(leg.iq, B * n, packed_dev)
    pin_p12 = torch.empty((B, 3 * n), dtype=torch.uint8).pin_memory()
    ctx._check(ctx.lib.mkid_memcpy(ctx.h, _lib.ptr(pin_p12), _lib.ptr(packed_dev), B * n * 3))
    ctx.sync()
    packed_dev.free()
    e2e_step_ms, e2e_value, words_p12 = e2e_run(pin_p12, 'p12')
    assert n_clipped == 0, n_clipped        # the packed stream holds the same samples
    del pin_p12
    if rank == 0:
        sampler.stop_flag = True
    del pin_words, pin_counts
    fir_int = np.array(chain.fir_int)
    # the channelizer kernel by itself (same launch: K4 + fused candidate mask, nothing else on the GPU): inside the
    # two-stream step its duration includes the moments it shares SMs with the detection tail of the previous batch
    k4_alone_ms = None
    try:
        chain.join(); ctx.sync()
        for _ in range(12):
            ctx._check(ctx.lib.mkid_chan_process(ctx.h, chain.chan.h, _lib.ptr(leg.iq), int(n), 2, None, 0, None, None))
        ctx.sync()
        k4_alone_ms = chain.chan.kernel_ms_sum(10) / 10
    except Exception:
        k4_alone_ms = None
    leg.free()
    stage('end-to-end done')

    extras = {}
    if not args.no_extras:
        # ------------------------------------------------------------ config 4 as written: the SAME 8 boards split over the GPUs
        if dist is not None and 8 % world == 0:
            try:
                Bs = 8 // world
                # two contexts per GPU here: detection / decode / merged list of batch k run under the channelizer kernel of
                # batch k + 1 (the latency-bound tail is what limits the step once a GPU has only a board or two)
                sl = ChainLeg(torch, ctx, reducer, Bs, n, N_LUT, N_ACTIVE, 253, 8, Bs * rank, 42 + Bs * rank, 1000 + rank,
                              args.hist_bins, exptime, pipelined=True)
                # >= 100 batches per job (6.5 s of data per board): the one reduce and the barriers are per job, not per batch
                r = sl.timed(max(args.steps, 100), 3, barrier, 1, dist)       # world = 1: total work is the 8 boards
                extras['strong_scaling'] = {'value': 8 * n / (r['step_ms'] * 1e-3) / 1e6, 'unit': UNIT, 'ms_per_step': r['step_ms'],
                                            'boards_total': 8, 'boards_per_gpu': Bs, 'n_gpus': world, 'k4_ms_per_launch': r['k4_ms'],
                                            'reduce_ms': r['reduce_ms'], 'scaling': 'strong', 'streams': 'pipelined over two contexts per GPU',
                                            'workload': 'BASELINE config 4: 8 boards sharded %d per GPU over %d GPUs' % (Bs, world)}
                sl.free()
            except Exception as e:      # side measurement only
                extras['strong_scaling'] = {'error': str(e)[:200]}
        # ------------------------------------------------------------ config 5: 20 000 resonators, NCCL-reduced [20000][4096]
        if (world == 8 or args.stress) and 80 % world == 0:
            try:
                extras['stress_config5'] = stress_leg(torch, ctx, reducer, dist, rank, world, barrier, measured_peaks()[0])
            except Exception as e:
                extras['stress_config5'] = {'error': str(e)[:200]}
        if dist is not None:
            for name, fn in (('decode_sharded', lambda: decode_sharded_bench(ctx, torch, dist, rank, world, measured_peaks()[0], reducer)),
                             ('lut_sharded', lambda: lut_sharded_bench(ctx, torch, dist, rank, world))):
                try:
                    extras[name] = fn()
                except Exception as e:
                    extras[name] = {'error': str(e)[:200]}
    stage('side legs done')
    if rank != 0:
        reducer.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peaks()
    achieved = BYTES_PER_SAMPLE * B * n / (res['k4_ms'] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'k4_traffic.json')
    if os.path.exists(tp):
        try:
            traffic = float(json.load(open(tp))['dram_bytes_per_sample']) * B * n
        except Exception:
            traffic = None
    line = {'metric': METRIC, 'value': res['value'], 'unit': UNIT, 'n_gpus': n_gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': res['step_ms'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': config,
            'roofline': {'bound': 'hbm', 'kernel': 'channelize_ws_kernel', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                         'kernel_ms_per_launch': res['k4_ms'], 'kernel_share_of_step': res['k4_ms'] / res['step_ms'],
                         'kernel_ms_alone': k4_alone_ms,
                         'frac_alone': (BYTES_PER_SAMPLE * B * n / (k4_alone_ms * 1e-3) / 1e9 / peak) if k4_alone_ms else None,
                         'note': 'algorithmic 4 B per complex ADC sample; the kernel is bound by the FP32 pipe / issue slots / shared '
                                 'memory together, not by HBM (111 FP32 lane operations per sample: FP32 roof = 0.20 of the HBM roof, '
                                 'see DESIGN.md)'},
            'collective': {'what': 'ONE mkid_hist_allreduce (NCCL sum, uint32) of counts [%d][%d] + hist [%d][%d] per job, on the '
                                   'context stream' % (exptime, n_pix, n_pix, args.hist_bins),
                           'bytes': res['reduce_bytes'], 'ms': res['reduce_ms'],
                           'bus_GB/s': (2.0 * (world - 1) / world * res['reduce_bytes'] / (res['reduce_ms'] * 1e-3) / 1e9) if world > 1 and res['reduce_ms'] > 0 else None,
                           'share_of_timed_region': res['reduce_ms'] / max(res['dev_ms'], 1e-9)},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(B * n * 3),
                    'd2h_bytes_per_step': int(B * cap * 8 + exptime * n_pix * 4 + B * 4), 'ms_per_step': e2e_step_ms, 'steps': k_e2e,
                    'adc_format': 'p12: 12-bit packed I/Q, 3 bytes per complex sample on the host link (the ADC is 12 bit; lossless, '
                                  'photon words identical to the int16 format), expanded by adc_unpack12_kernel in front of K4',
                    'api': "ReadoutChain.process_stream(adc_format='p12') (upload of batch k+1 overlaps the kernels of batch k)", 'numa': numa},
            'e2e_int16': {'value': e2e16_value, 'unit': UNIT, 'h2d_bytes_per_step': int(B * n * 4),
                          'd2h_bytes_per_step': int(B * cap * 8 + exptime * n_pix * 4 + B * 4), 'ms_per_step': e2e16_step_ms, 'steps': k_e2e,
                          'adc_format': 'i16: int16 I/Q pairs, 4 bytes per complex sample', 'api': 'ReadoutChain.process_stream'},
            'gpu_launches': res['launches'],
            'photon_words_per_step': words_per_step,
            'clocks': sampler.summary()}
    line.update(extras)
    if not args.no_extras:
        for name, fn in (('decode', lambda: decode_side_bench(ctx, peak)), ('lut', lambda: lut_side_bench(ctx))):
            try:
                line[name] = fn()
            except Exception as e:      # side measurement only
                line[name] = {'error': str(e)[:200]}
    if not args.no_cpu_baseline:
        try:
            from oracle import cpu_arm
            arm = cpu_arm.CpuArm(cores, 1 << 21, fir_int)
            secs = [arm.step() for _ in range(3)]
            arm.close()
            dt = float(np.mean(secs))
            line['cpu_baseline'] = {'value': cores * (1 << 21) / dt / 1e6, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                    'sample': arm.describe(), 'seconds_per_pass': dt, 'passes': 3,
                                    'setup_seconds_untimed': arm.setup_seconds}
        except Exception as e:
            line['cpu_baseline'] = {'error': str(e)[:200]}
    emit(line)
    reducer.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def stress_leg(torch, ctx, reducer, dist, rank, world, barrier, peak):
    """BASELINE config 5 / SURVEY 8d config 5: 20 000 resonators = 80 board streams of 250 active channels (10 feedlines x 8
    sub-boards), 80 / world per GPU, channelize + detect + per-pixel 4096-bin peak histograms [20000][4096] u32 (328 MB)
    summed over the GPUs by ONE NCCL all-reduce (mkid_hist_allreduce) per job."""
    Bs = 80 // world
    n = 1 << 23
    steps = 32                       # batches of 16 ms per job: the reduce is once per job (0.5 s of data)
    leg = ChainLeg(torch, ctx, reducer, Bs, n, N_LUT, 250, 250, 80, Bs * rank, 100 + Bs * rank, 2000 + rank, 4096, 4,
                   want_merged=False, pipelined=True)
    r = leg.timed(steps, 3, barrier, world, dist)
    # checksum across ranks: the reduced histogram must be identical everywhere and hold every binned word
    cs = torch.stack([leg.products[leg.n_counts:].sum(dtype=torch.int64), leg.products[:leg.n_counts].sum(dtype=torch.int64)])
    allcs = [torch.zeros_like(cs) for _ in range(world)]
    if dist is not None:
        dist.all_gather(allcs, cs)
    else:
        allcs = [cs]
    same = all(bool((c == allcs[0]).all()) for c in allcs)
    leg.free()
    bus = (2.0 * (world - 1) / world * r['reduce_bytes'] / (r['reduce_ms'] * 1e-3) / 1e9) if world > 1 and r['reduce_ms'] > 0 else None
    return {'value': r['value'], 'unit': UNIT, 'ms_per_step': r['step_ms'], 'steps': steps, 'boards_total': 80, 'boards_per_gpu': Bs,
            'resonators': 80 * 250, 'samples_per_board_per_step': n, 'k4_ms_per_launch': r['k4_ms'],
            'hist_shape': [20000, 4096], 'reduce_bytes': r['reduce_bytes'], 'reduce_ms': r['reduce_ms'], 'reduce_bus_GB/s': bus,
            'reduce_share_of_job': r['reduce_ms'] / max(r['dev_ms'], 1e-9), 'checksum_identical_on_all_ranks': same,
            'hist_sum': int(allcs[0][0]), 'counts_sum': int(allcs[0][1]),
            'frac_hbm_k4': 4.0 * Bs * n / (r['k4_ms'] * 1e-3) / 1e9 / peak}


def decode_sharded_bench(ctx, torch, dist, rank, world, peak, reducer):
    """The second sharding mode of SURVEY 8e: a photon file set is split by packet-file chunk across the GPUs (every
    rank decodes its own replicas of the 8-roach file, 1.28 GB per GPU resident in HBM: weak scaling), then ONE NCCL
    sum-reduce of the per-pixel products (counts [10][2024] + 10-bin spectra [2024][10]).  Whole-job words/s, device
    events, max over ranks."""
    from mkids_sdr_b200 import synth
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, secs, reps = 8, 253, 10, 16
    streams, _ = synth.photon_streams(10 ** 7, R, npix, secs, seed=1234 + rank)
    lens = [len(s) for s in streams]
    words = np.tile(np.concatenate(streams), reps)
    offs = np.concatenate([[0], np.cumsum(lens * reps)]).astype(np.int64)
    roach = np.tile(np.arange(R), reps)
    dw = ctx.to_device(words)
    # counts and spectra share one tensor: ONE all-reduce per pass.  The collective is enqueued on the context's own stream
    # (torch.cuda.ExternalStream), so a pass needs no host synchronisation: reset -> decode -> reduce queue up.
    n_counts = secs * R * npix
    buf_t = torch.zeros(n_counts + R * npix * 10, dtype=torch.int32, device='cuda')
    counts_t, hist_t = buf_t[:n_counts], buf_t[n_counts:]
    lut = np.arange(4096) * 10 // 4096
    dec = PhotonDecoder(R, npix, secs, 2500, 'p1', 10, lut, ctx=ctx, counts_buf=counts_t, hist_buf=hist_t)
    torch.cuda.synchronize()                     # buf_t was zeroed on torch's stream

    def one_pass():
        dec.reset()                              # every pass is one complete job: partial products, then the reduce
        dec.decode_words(dw, offs, roach, want_stats=False, want_sec=False)
        reducer.allreduce(buf_t, buf_t.numel())      # mkid_hist_allreduce on the context stream
    for _ in range(3):
        one_pass()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    k = 10
    t0 = time.time()
    for _ in range(k):
        one_pass()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    el = torch.tensor([(time.time() - t0) * 1e3 / k], dtype=torch.float64, device='cuda')
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
    ms = float(el[0])
    total = words.size * world
    total_hist = int(hist_t.sum().item())
    dw.free()
    gbs = total * 8 / ms / 1e6
    return {'words_per_s': total / ms * 1e3, 'GB/s': gbs, 'frac_hbm_per_gpu': gbs / world / peak, 'ms_per_pass': ms, 'n_gpus': world,
            'collective': 'ONE mkid_hist_allreduce (NCCL sum) of counts [10][2024] + spectra [2024][10] per pass, queued on the context stream (no host sync inside a pass)',
            'checksum_spectra': total_hist, 'checksum_expected': int(10 ** 7 * reps * world),
            'checksum_ok': total_hist == int(10 ** 7 * reps * world),
            'workload': '16 x 1e7 photon words per GPU (different seeds), decode + per-pixel counts + 10-bin spectra'}


def lut_sharded_bench(ctx, torch, dist, rank, world):
    """SURVEY 8e, first row: the LUT sets of the boards are independent, boards -> GPUs, no collective on the data
    path.  Every rank synthesises the whole LUT sets (comb + 256 DDS tables + DRAM image) of 64 boards, left in HBM
    (weak scaling); whole-job sets/s, max over ranks.  The boards of all ranks use the same tones, so the images must be
    identical on every rank: one all-gather of a checksum (outside the timed region) checks that."""
    from mkids_sdr_b200 import _lib, lut
    N, T, batch = 2 ** 19, 256, 64
    ctx2 = _lib.Context(ctx.device)                  # the DDS tables of the batch run on a second stream under the comb
    k = np.sort(np.random.default_rng(0).choice(np.arange(-N // 2 + 1, N // 2), T, replace=False))
    f = (k % N) * FS / N
    amps = 10 ** (-(np.random.default_rng(1).integers(0, 20, T)) / 20.)
    res = FS / N
    resid = np.rint((f - np.rint(f * 512 / FS) * FS / 512) / res) * res      # select_bins (ROACH_Setup.py:534-550)
    ff = np.tile(f, (batch, 1)); aa = np.tile(amps, (batch, 1)); rr = np.tile(resid, (batch, 1))
    bufs = [ctx.alloc(batch * N * 2) for _ in range(4)]
    img = ctx.alloc(batch * N * 8)

    zz = np.zeros_like(rr)

    def one():
        lut.dds_lut(rr, zz, FS, N, ctx=ctx2, out_I=bufs[2], out_Q=bufs[3], want_scales=False)
        ctx2.record(50)
        lut.comb_lut(ff, FS, N, aa, ctx=ctx, out_I=bufs[0], out_Q=bufs[1])
        ctx.wait_event(ctx2, 50)
        lut.pack_dram(bufs[0], bufs[1], bufs[2], bufs[3], ctx=ctx, n=batch * N, out=img)
    for _ in range(3):
        one()
    ctx.sync(); dist.barrier(); torch.cuda.synchronize()
    reps = 10
    t0 = time.time()
    for _ in range(reps):
        one()
    ctx.sync(); dist.barrier(); torch.cuda.synchronize()
    el = torch.tensor([(time.time() - t0) * 1e3 / reps], dtype=torch.float64, device='cuda')
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
    ms = float(el[0])
    first = img.download(np.uint8, count=8 * N)                       # the image of the first board of this rank
    cs = torch.tensor([int(first.astype(np.int64).sum())], dtype=torch.int64, device='cuda')
    allcs = [torch.zeros_like(cs) for _ in range(world)]
    dist.all_gather(allcs, cs)
    for v in bufs + [img]:
        v.free()
    ctx2.close()
    return {'luts_per_s': world * batch / ms * 1e3, 'ms_per_call': ms, 'sets_per_gpu_per_call': batch, 'n_gpus': world,
            'GB/s_written': world * batch * N * 16 / ms / 1e6, 'collective': 'none on the data path',
            'checksum_ok': all(int(c[0]) == int(cs[0]) for c in allcs),
            'workload': 'whole LUT sets (comb + 256 DDS tables on a second stream + DRAM image), 64 boards per GPU'}


def decode_side_bench(ctx, peak):
    """BASELINE config 0 on the GPU: decode + per-pixel binning of 1e7-word synthetic photon files
    (replicated 16x = 1.28 GB so the input exceeds L2), device resident."""
    from mkids_sdr_b200 import synth
    from mkids_sdr_b200.decode import PhotonDecoder
    from mkids_sdr_b200._lib import ptr as _lib_ptr
    R, npix, secs = 8, 253, 10
    streams, _ = synth.photon_streams(10 ** 7, R, npix, secs, seed=1234)
    lens = [len(s) for s in streams]
    reps = 16
    words = np.tile(np.concatenate(streams), reps)
    offs = np.concatenate([[0], np.cumsum(lens * reps)]).astype(np.int64)
    roach = np.tile(np.arange(R), reps)
    dw = ctx.to_device(words)
    out = {}
    for name, field, nb in (('counts_only', None, 1), ('counts_hist10', 'p1', 10), ('counts_hist4096', 'peak', 4096)):
        lut = (np.arange(4096) * 10 // 4096) if nb == 10 else None
        dec = PhotonDecoder(R, npix, secs, 2500, field, nb, lut, ctx=ctx)
        for _ in range(3):
            dec.decode_words(dw, offs, roach, want_stats=False, want_sec=False)
        ctx.sync(); ctx.record(2)
        k = 10
        for _ in range(k):
            dec.decode_words(dw, offs, roach, want_stats=False, want_sec=False)
        ctx.record(3)
        ms = ctx.elapsed_ms(2, 3) / k
        gbs = words.size * 8 / ms / 1e6
        out[name] = {'words_per_s': words.size / ms * 1e3, 'GB/s': gbs, 'frac_hbm': gbs / peak, 'ms': ms}
    # the same file in the wire format PacketMaster receives (PulseServer bundles: 32 KiB of big-endian low halves,
    # then 32 KiB of big-endian high halves), SURVEY 8d config 1
    try:
        wire = np.concatenate(synth.streams_to_wire(streams))
        nb = [len(s) // 8192 for s in streams]
        wire_big = np.tile(wire, reps)
        woffs = np.concatenate([[0], np.cumsum(nb * reps)]).astype(np.int64)
        dwire = ctx.to_device(wire_big)
        for name, field, nbins in (('wire_counts_only', None, 1), ('wire_counts_hist10', 'p1', 10)):
            lut = (np.arange(4096) * 10 // 4096) if nbins == 10 else None
            dec = PhotonDecoder(R, npix, secs, 2500, field, nbins, lut, ctx=ctx)
            for _ in range(3):
                dec.decode_wire(dwire, woffs, roach, want_stats=False, want_sec=False)
            ctx.sync(); ctx.record(2)
            k = 10
            for _ in range(k):
                dec.decode_wire(dwire, woffs, roach, want_stats=False, want_sec=False)
            ctx.record(3)
            ms = ctx.elapsed_ms(2, 3) / k
            gbs = wire_big.size / ms / 1e6
            # counts and the peak/p1 spectra only need the high halves: the 32 KiB low-half blocks are never touched
            out[name] = {'words_per_s': wire_big.size / 8 / ms * 1e3, 'GB/s': gbs, 'frac_hbm': gbs / peak, 'ms': ms,
                         'touched_bytes_per_word': 4, 'GB/s_touched': gbs / 2, 'frac_hbm_touched': gbs / 2 / peak}
        dwire.free()
    except Exception as e:
        out['wire_counts_only'] = {'error': str(e)}
    # PacketMaster's per-(second, pixel) photon lists (16 B/word algorithmic: read + sorted write); every replica of the
    # file continues the seconds of the previous one so that all keys are distinct
    try:
        dec = PhotonDecoder(R, npix, secs * reps, 2500, None, 1, None, ctx=ctx)
        sec0 = np.repeat(np.arange(reps) * secs, R).astype(np.int32)
        lw_dev = ctx.alloc(words.size * 8)
        lo_dev = ctx.alloc((secs * reps * R * npix + 1) * 8)
        import ctypes as _ct
        sec_out = np.zeros(roach.size, dtype=np.int32)
        roach32 = roach.astype(np.int32)

        def run_lists():
            ctx._check(ctx.lib.mkid_decode_lists(ctx.h, _lib_ptr(dw), words.size, _lib_ptr(offs), _lib_ptr(roach32),
                                                 _lib_ptr(sec0), _lib_ptr(sec_out), roach.size, _ct.byref(dec.cfg),
                                                 _lib_ptr(dec.counts_dev), _lib_ptr(lw_dev), words.size, _lib_ptr(lo_dev), None))
        run_lists()
        ctx.sync(); ctx.record(2)
        k = 3
        for _ in range(k):
            run_lists()
        ctx.record(3)
        ms = ctx.elapsed_ms(2, 3) / k
        gbs = words.size * 16 / ms / 1e6
        out['photon_lists'] = {'words_per_s': words.size / ms * 1e3, 'GB/s': gbs, 'frac_hbm': gbs / peak, 'ms': ms,
                               'algorithmic_bytes_per_word': 16}
        # the time-ordered merged list of config 4 (key = (second, roach)): same 16 B/word, contiguous writes
        def run_merged():
            ctx._check(ctx.lib.mkid_decode_merged(ctx.h, _lib_ptr(dw), words.size, _lib_ptr(offs), _lib_ptr(roach32),
                                                  _lib_ptr(sec0), _lib_ptr(sec_out), roach.size, _ct.byref(dec.cfg),
                                                  _lib_ptr(dec.counts_dev), _lib_ptr(lw_dev), words.size, _lib_ptr(lo_dev), None))
        run_merged()
        ctx.sync(); ctx.record(2)
        for _ in range(k):
            run_merged()
        ctx.record(3)
        ms = ctx.elapsed_ms(2, 3) / k
        gbs = words.size * 16 / ms / 1e6
        out['merged_list'] = {'words_per_s': words.size / ms * 1e3, 'GB/s': gbs, 'frac_hbm': gbs / peak, 'ms': ms,
                              'algorithmic_bytes_per_word': 16}
        lw_dev.free(); lo_dev.free()
    except Exception as e:
        out['photon_lists'] = {'error': str(e)}
    tp = os.path.join(ROOT, 'profiles', 'k6_traffic.json')
    if os.path.exists(tp):
        try:
            out['dram_bytes_per_word_ncu'] = float(json.load(open(tp))['dram_bytes_per_word'])
        except Exception:
            pass
    out['algorithmic_bytes_per_word'] = 8
    out['workload'] = 'Utils/bin.py-style decode + per-pixel binning, 2024 pixels, 16 x 1e7 photon words resident in HBM'
    return out


def lut_side_bench(ctx):
    """BASELINE config 1: 256 tones, 2^19-sample int16 I/Q comb + DDS LUT + DRAM image."""
    from mkids_sdr_b200 import _lib, lut
    N, T = 2 ** 19, 256
    k = np.sort(np.random.default_rng(0).choice(np.arange(-N // 2 + 1, N // 2), T, replace=False))
    f = (k % N) * FS / N
    amps = 10 ** (-(np.random.default_rng(1).integers(0, 20, T)) / 20.)
    out = {}
    for batch in (1, 8, 64, 512):                  # batch sizes 1, 8, 64, 512
        ff = np.tile(f, (batch, 1)); aa = np.tile(amps, (batch, 1))
        for where in ('host', 'device'):
            kw = {}
            if where == 'device':       # tables stay in HBM (the usual consumer, mkid_pack_dram / the channelizer, is on the GPU)
                kw = dict(out_I=ctx.alloc(batch * N * 2), out_Q=ctx.alloc(batch * N * 2))
            elif batch > 8:
                continue
            lut.comb_lut(ff, FS, N, aa, ctx=ctx, **kw)
            ctx.sync()
            t0 = time.time()
            reps = 5
            for _ in range(reps):
                lut.comb_lut(ff, FS, N, aa, ctx=ctx, **kw)
            ctx.sync()
            dt = (time.time() - t0) / reps
            out['comb_batch%d_%s' % (batch, where)] = {'luts_per_s': batch / dt, 'ms_per_call': dt * 1e3,
                                                       'GB/s_written': batch * N * 4 / dt / 1e9}
            for v in kw.values():
                v.free()
    # the whole LUT set of one board: comb + the 256 DDS tables + the 8*N-byte DRAM image (define_LUTs + write_LUTs),
    # everything left in HBM
    res = FS / N
    resid = np.rint((f - np.rint(f * 512 / FS) * FS / 512) / res) * res      # select_bins (ROACH_Setup.py:534-550)
    ctx2 = _lib.Context(ctx.device)                  # second stream: the DDS tables do not depend on the comb
    for batch in (1, 64):
        ff = np.tile(f, (batch, 1)); aa = np.tile(amps, (batch, 1)); rr = np.tile(resid, (batch, 1))
        zz = np.zeros_like(rr)
        bufs = [ctx.alloc(batch * N * 2) for _ in range(4)]
        img = ctx.alloc(batch * N * 8)

        def one():
            lut.comb_lut(ff, FS, N, aa, ctx=ctx, out_I=bufs[0], out_Q=bufs[1])
            lut.dds_lut(rr, zz, FS, N, ctx=ctx, out_I=bufs[2], out_Q=bufs[3])
            lut.pack_dram(bufs[0], bufs[1], bufs[2], bufs[3], ctx=ctx, n=batch * N, out=img)

        def one_two_streams():
            # the 256 DDS tables of every set on the second context while the first one synthesises the comb; the DRAM
            # image waits for both
            lut.dds_lut(rr, zz, FS, N, ctx=ctx2, out_I=bufs[2], out_Q=bufs[3], want_scales=False)
            ctx2.record(50)
            lut.comb_lut(ff, FS, N, aa, ctx=ctx, out_I=bufs[0], out_Q=bufs[1])
            ctx.wait_event(ctx2, 50)
            lut.pack_dram(bufs[0], bufs[1], bufs[2], bufs[3], ctx=ctx, n=batch * N, out=img)
        for name, fn in (('full_set_batch%d_device' % batch, one), ('full_set_batch%d_device_two_streams' % batch, one_two_streams)):
            fn()
            ctx.sync(); ctx2.sync()
            t0 = time.time()
            reps = 5
            for _ in range(reps):
                fn()
            ctx.sync(); ctx2.sync()
            dt = (time.time() - t0) / reps
            out[name] = {'luts_per_s': batch / dt, 'ms_per_call': dt * 1e3, 'GB/s_written': batch * N * 16 / dt / 1e9}
        for v in bufs + [img]:
            v.free()
    ctx2.close()
    out['workload'] = ('256 tones, 2^19-sample int16 I/Q comb (freqCombLUT incl. seed-1000 phases, scale, exact quantisation); '
                       'full_set = comb + 256 DDS tables + DRAM image (16 B per sample written)')
    return out


if __name__ == '__main__':
    try:
        rc = main()
    except BaseException:
        import traceback
        sys.stderr.write('[bench rank %s] failed:\n%s\n' % (os.environ.get('RANK', '0'), traceback.format_exc()))
        sys.stderr.flush()
        raise
    sys.exit(rc)
