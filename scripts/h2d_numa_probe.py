"""Aggregate pinned host -> device bandwidth of N ranks under different CPU / memory placements (diagnostic for the
end-to-end leg of bench.py).   torchrun --nproc-per-node N scripts/h2d_numa_probe.py"""
import glob
import os
import time

import torch
import torch.distributed as dist

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
nodes = sorted(int(p.rsplit('node', 1)[1]) for p in glob.glob('/sys/devices/system/node/node[0-9]*'))
prop = torch.cuda.get_device_properties(lr)
bus = '%04x:%02x:%02x.0' % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
try:
    gnode = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read())
except Exception as e:
    gnode = 'err %s' % e
all_cpus = sorted(os.sched_getaffinity(0))
if rank == 0:
    print('nodes', nodes, 'cpus allowed', len(all_cpus), all_cpus[:4], '...', all_cpus[-4:], flush=True)
    for n in nodes:
        print(' node', n, open('/sys/devices/system/node/node%d/cpulist' % n).read().strip(), flush=True)
print('rank', rank, 'gpu', bus, 'numa_node', gnode, flush=True)


def cpus_of(node):
    out = set()
    for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
        lo, _, hi = part.partition('-')
        out.update(range(int(lo), int(hi or lo) + 1))
    return out


def run(label, node):
    os.sched_setaffinity(0, all_cpus)
    if node is not None:
        c = cpus_of(node) & set(all_cpus)
        if c:
            os.sched_setaffinity(0, c)
    host = torch.empty(1 << 28, dtype=torch.int32).pin_memory()      # 1 GiB
    host.fill_(rank)
    dev = torch.empty(1 << 28, dtype=torch.int32, device='cuda')
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(8):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        print('%-28s aggregate %.1f GB/s (%.1f per GPU)' % (label, world * 8 * 1.0737 / dt, 8 * 1.0737 / dt), flush=True)
    del host, dev


run('default placement', None)
if len(nodes) > 1:
    run('spread rank*nodes//world', nodes[rank * len(nodes) // world])
    run('spread rank % nodes', nodes[rank % len(nodes)])
    run('all on node 0', nodes[0])
dist.destroy_process_group()
