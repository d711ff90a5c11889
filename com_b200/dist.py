"""Multi-GPU plumbing of the hot path: frames are independent units, so inference
shards frames over ranks with NO data-path collective (reference: DistributedSampler(shuffle=False),
pcdet/datasets/__init__.py:27-47).  torch.distributed is used only to rendezvous, to barrier and to
reduce timings / counters (max or sum over ranks)."""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    """Initialise the default process group when launched by torchrun; no-op for a single process."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def frame_shard(n_frames, rank, world):
    """Indices of the frames rank `rank` owns: round-robin like DistributedSampler(shuffle=False)
    WITHOUT its padding duplicates (every frame is processed exactly once)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, int(n_frames), world))


def _reduce(value, op, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(value, device=None):
    return _reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value, device=None):
    return _reduce(value, dist.ReduceOp.SUM, device)


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


# ------------------------------------------------------------------------------------------- gradient synchronisation
class FlatGradSync:
    """Data-parallel gradient averaging for a step whose backward hands over ALL parameter gradients at once (the fused
    train step, com_b200/train.py): ONE all-reduce of one flat fp32 buffer instead of torch DDP's reducer (bucket hooks
    that fire as autograd produces gradients — there is nothing to overlap with when the whole backward is two CUDA
    graphs, and on 2 B200 the reducer cost +1.55 ms on a 4.7 ms step, efficiency 0.75).  Same result as
    `DistributedDataParallel` (tools/train.py:166 of the reference): the mean of the ranks' gradients.  NVSwitch carries
    the 10.8 MB of the backbone in one NCCL ring / NVLS pass.

        sync = FlatGradSync(net.parameters()); ... loss.backward(); sync(); opt.step()
    """

    def __init__(self, params, group=None, flat_provider=None):
        """flat_provider: optional callable returning the ONE tensor that already holds every gradient (the fused train
        step's backward graph writes them into one buffer and autograd hands out views of it): the all-reduce then runs
        in place on that buffer and nothing is gathered or copied back."""
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.sizes = [p.numel() for p in self.params]
        self.flat_provider = flat_provider
        self.flat = None
        self.in_place = False

    def _views_of(self, flat):
        """True when every p.grad is a dense view into `flat` (so reducing `flat` reduces the gradients)."""
        if flat is None or flat.dtype != torch.float32 or flat.numel() != sum(self.sizes):
            return False
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        return all(p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous()
                   and lo <= p.grad.data_ptr() and p.grad.data_ptr() + p.grad.numel() * 4 <= hi for p in self.params)

    def __call__(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        flat = self.flat_provider() if self.flat_provider is not None else None
        self.in_place = self._views_of(flat)
        if self.in_place:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(dist.get_world_size(self.group))
            self.flat = flat
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        flat = torch.cat([g.reshape(-1).float() for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(dist.get_world_size(self.group))
        outs = [c.view_as(p).to(p.dtype) for c, p in zip(flat.split(self.sizes), self.params)]
        for p, g, o in zip(self.params, grads, outs):
            if p.grad is None:
                p.grad = o
        live = [(g, o) for p, g, o in zip(self.params, grads, outs) if p.grad is g]
        if live:
            torch._foreach_copy_([g for g, _ in live], [o for _, o in live])
        self.flat = flat


# ---------------------------------------------------------------------------------------------------- NUMA placement
def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(index):
    """NUMA node of GPU `index` from sysfs (None when the platform does not say)."""
    try:
        pr = torch.cuda.get_device_properties(index)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = "%04x:%02x:%02x.0" % (int(pr.pci_domain_id), int(pr.pci_bus_id), int(pr.pci_device_id))
        else:
            import pynvml
            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else str(bus)).lower()
            if len(bus.split(":")[0]) == 8:              # NVML prints an 8-digit PCI domain, sysfs uses 4
                bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa(local_rank):
    """Run this process — and, by first touch, allocate its pinned staging buffers — on the NUMA node the rank's GPU hangs
    off.  One process per GPU on a two-socket box otherwise lands wherever the launcher was started: r1's 8-GPU run
    had all ranks on node 0 (`topology`), so the H2D/D2H traffic of GPUs 4-7 (4 x ~20 GB/s) crossed the socket
    interconnect and end-to-end scaling bent to 0.70 while the device-resident number scaled at 0.99.
    Best effort: CPU affinity to the node's cores that the cpuset allows, and a preferred-node memory policy
    (set_mempolicy) so pinned pages come from that node even when the cpuset keeps the threads elsewhere.
    Call BEFORE allocating pinned memory.  Returns a dict describing what was done (goes into the bench line)."""
    import ctypes
    info = {"gpu": int(local_rank), "node": None, "cpus_bound": 0, "mempolicy": False}
    if os.environ.get("COMB_NUMA_BIND", "1") == "0":
        info["disabled"] = True
        return info
    node = gpu_numa_node(local_rank)
    info["node"] = node
    if node is None:
        return info
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            node_cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        want = node_cpus & allowed
        if want:
            os.sched_setaffinity(0, want)
            info["cpus_bound"] = len(want)
    except Exception as e:
        info["affinity_error"] = "%s: %s" % (type(e).__name__, e)
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238       # x86_64
        nbits = 64 * ((node // 64) + 1)
        mask = (ctypes.c_ulong * (nbits // 64))()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), nbits + 1)
        info["mempolicy"] = rc == 0
    except Exception as e:
        info["mempolicy_error"] = "%s: %s" % (type(e).__name__, e)
    return info
