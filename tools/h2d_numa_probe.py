"""Where the end-to-end leg of bench.py loses bandwidth on N GPUs of one box: concurrent host-to-device copies of one 300 MB
pinned shard per rank (1.25e7 x 6 float32 candidates), with the pinned pages placed (a) wherever the launcher left the
process, (b) on the NUMA node of the rank's GPU, (c) on the other node, (d) interleaved over all nodes.  Run under
torch.distributed.run; prints the topology and per-rank / aggregate GB/s of each placement.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_numa_probe.py
"""
import ctypes
import glob
import os

import torch
import torch.distributed as td


def cpulist(text):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def nodes():
    res = {}
    for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        res[int(p.rsplit("node", 1)[1])] = cpulist(open(p + "/cpulist").read())
    return res


def set_mempolicy(mode, nodemask_bits, maxnode=64):
    """mode 0 default, 2 bind, 3 interleave (linux/mempolicy.h); syscall 238 on x86_64."""
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(nodemask_bits)
    rc = libc.syscall(238, ctypes.c_int(mode), ctypes.byref(mask), ctypes.c_ulong(maxnode))
    return rc, ctypes.get_errno()


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    nd = nodes()
    bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        if isinstance(bus, bytes):
            bus = bus.decode()
    except Exception as e:                                        # noqa: BLE001
        bus = bus or f"? ({e})"
    gpu_node = -1
    for cand in (str(bus).lower(), str(bus).lower()[4:]):
        p = f"/sys/bus/pci/devices/{cand}/numa_node"
        if os.path.exists(p):
            gpu_node = int(open(p).read())
            break
    aff0 = sorted(os.sched_getaffinity(0))
    print(f"[rank {rank}] gpu {local} bus {bus} numa_node {gpu_node}; nodes {{{', '.join(f'{k}: {len(v)} cpus' for k, v in nd.items())}}}; "
          f"affinity {len(aff0)} cpus ({aff0[0]}..{aff0[-1]}); OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')}", flush=True)
    n = 12_500_000 * 6
    dev = torch.empty(n, dtype=torch.float32, device="cuda")
    src = torch.rand(n, dtype=torch.float32)

    def measure(tag):
        host = src.pin_memory()                                   # cudaHostAlloc + copy in this thread: pages land under the current policy
        for _ in range(2):
            dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(6):
            dev.copy_(host, non_blocking=True)
        ev[1].record()
        torch.cuda.synchronize()
        gbs = torch.tensor([6 * n * 4 / ev[0].elapsed_time(ev[1]) / 1e6], device="cuda")
        if world > 1:
            allg = [torch.zeros_like(gbs) for _ in range(world)]
            td.all_gather(allg, gbs)
        else:
            allg = [gbs]
        vals = [float(v) for v in allg]
        if rank == 0:
            print(f"{tag:32s} per rank {' '.join(f'{v:5.1f}' for v in vals)}  sum {sum(vals):6.1f} GB/s", flush=True)
        del host

    measure("as launched")

    def measure_wc(tag):
        """The same with write-combined pinned memory (cudaHostAlloc, cudaHostAllocWriteCombined): no cache snooping on the
        DMA reads."""
        rt = ctypes.CDLL("libcudart.so.12")
        ptr = ctypes.c_void_p()
        rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n * 4), ctypes.c_uint(4))
        if rc != 0:
            if rank == 0:
                print(f"{tag}: cudaHostAlloc failed ({rc})", flush=True)
            return
        buf = (ctypes.c_float * n).from_address(ptr.value)
        host = torch.frombuffer(buf, dtype=torch.float32)
        host.copy_(src)
        if rank == 0:
            print(f"   write-combined buffer is_pinned={host.is_pinned()}", flush=True)
        for _ in range(2):
            dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(6):
            dev.copy_(host, non_blocking=True)
        ev[1].record()
        torch.cuda.synchronize()
        gbs = torch.tensor([6 * n * 4 / ev[0].elapsed_time(ev[1]) / 1e6], device="cuda")
        allg = [torch.zeros_like(gbs) for _ in range(world)]
        if world > 1:
            td.all_gather(allg, gbs)
        else:
            allg = [gbs]
        vals = [float(v) for v in allg]
        if rank == 0:
            print(f"{tag:32s} per rank {' '.join(f'{v:5.1f}' for v in vals)}  sum {sum(vals):6.1f} GB/s", flush=True)
        del host, buf
        rt.cudaFreeHost(ptr)

    measure_wc("write-combined pinned")
    all_nodes = sorted(nd)
    if gpu_node >= 0 and gpu_node in nd:
        os.sched_setaffinity(0, nd[gpu_node])
        rc = set_mempolicy(2, 1 << gpu_node)
        measure(f"bound to the GPU's node (rc {rc[0]})")
        others = [k for k in all_nodes if k != gpu_node]
        if others:
            o = others[(local) % len(others)]
            os.sched_setaffinity(0, nd[o])
            rc = set_mempolicy(2, 1 << o)
            measure(f"bound to another node (rc {rc[0]})")
    os.sched_setaffinity(0, aff0)
    bits = 0
    for k in all_nodes:
        bits |= 1 << k
    rc = set_mempolicy(3, bits)
    measure(f"interleaved over {len(all_nodes)} nodes (rc {rc[0]})")
    set_mempolicy(0, 0)
    # spread: rank r on node r % n_nodes regardless of the GPU's node
    o = all_nodes[local % len(all_nodes)]
    os.sched_setaffinity(0, nd[o])
    rc = set_mempolicy(2, 1 << o)
    measure(f"rank r on node r % {len(all_nodes)} (rc {rc[0]})")
    os.sched_setaffinity(0, aff0)
    set_mempolicy(0, 0)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
