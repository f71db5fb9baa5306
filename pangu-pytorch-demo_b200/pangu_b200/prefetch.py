"""Input pipeline for the train / test loops: batches reach the device one step ahead.

Mirror of the reference's `DataPrefetcher` (era5_data/utils_data.py:20-57: `DataPrefetcher(loader)`, `.next()` ->
`(input, input_surface, target, target_surface, periods)` on the GPU, `len()`, wraps around at the end of the loader),
which the reference defines but never wires in -- its loops call `.to(device)` on pageable tensors inside the step
(models/pangu_sample.py:152-155), a synchronous 2 x 287 MB copy per sample.

Differences, on purpose:
  * every batch is staged through PINNED host buffers owned by the prefetcher (two slots), so the H2D copies are truly
    asynchronous whether or not the DataLoader pins its output (`torch.load` of a .pt file never does,
    era5_data/utils_data.py:344-362);
  * `.next()` returns the batch whose copy was started one call earlier and only then starts the copy of the following
    one (the reference's version overwrites the batch it is about to return);
  * the device tensors are handed to the caller's stream with `record_stream`, so the allocator cannot recycle them while
    the step still reads them.
No CPU path: constructing it without CUDA raises.
"""
import torch

from .abi import PanguError


def bind_host_thread_to_gpu(device=None):
    """Pin the calling thread to the CPUs of the GPU's NUMA node (NVML's ideal-CPU set for the device) so that the pinned
    staging buffers allocated AFTERWARDS are local to the GPU's PCIe root.  With one process per GPU this matters as soon as
    several ranks copy at once: un-bound ranks measured 22 GB/s device->host at 4 GPUs against 55 GB/s alone (bench.py
    `e2e.transfers_alone`).  Returns the CPU list, or None when NVML / the affinity call is unavailable (nothing is changed)."""
    import os
    try:
        import pynvml
        idx = torch.cuda.current_device() if device is None else torch.device(device).index
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:                                   # torch's ordinal -> the physical device NVML enumerates
            ent = visible.split(",")[idx].strip()
            h = pynvml.nvmlDeviceGetHandleByUUID(ent) if ent.startswith("GPU-") else pynvml.nvmlDeviceGetHandleByIndex(int(ent))
        else:
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:                                 # noqa: BLE001 -- an optimisation only: never fail the caller
        return None


class DataPrefetcher:
    def __init__(self, loader, device=None, slots=2):
        if not torch.cuda.is_available():
            raise PanguError("pangu_b200.prefetch.DataPrefetcher needs a CUDA device (no CPU fallback)")
        self.loader = loader
        self.length = len(loader)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.slots = slots
        self._pinned = [None] * slots            # per slot: list of pinned staging tensors (allocated on first use)
        self._copied = [torch.cuda.Event() for _ in range(slots)]
        self._slot = 0
        self.dataiter = iter(loader)
        self._ahead = None
        self._preload()

    def _stage(self, k, i, t):
        """Pinned staging tensor i of slot k, shaped like t."""
        bufs = self._pinned[k]
        if i >= len(bufs):
            bufs.extend([None] * (i + 1 - len(bufs)))
        b = bufs[i]
        if b is None or b.shape != t.shape or b.dtype != t.dtype:
            b = bufs[i] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        return b

    def _preload(self):
        try:
            batch = next(self.dataiter)
        except StopIteration:                    # era5_data/utils_data.py:39-42: start over
            self.dataiter = iter(self.loader)
            batch = next(self.dataiter)
        k = self._slot
        self._slot = (k + 1) % self.slots
        if self._pinned[k] is None:
            self._pinned[k] = []
        else:
            self._copied[k].synchronize()        # the slot's previous H2D copies have left the pinned buffers
        dev = []
        with torch.cuda.stream(self.stream):
            for i, t in enumerate(batch):
                if not torch.is_tensor(t):
                    t = torch.as_tensor(t)
                if t.is_cuda:
                    dev.append(t.to(self.device, non_blocking=True))
                    continue
                src = t if t.is_pinned() else self._stage(k, i, t).copy_(t)
                dev.append(src.to(self.device, non_blocking=True))
            self._copied[k].record(self.stream)
        self._ahead = tuple(dev)

    def next(self):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream)
        batch = self._ahead
        for t in batch:
            t.record_stream(cur)
        self._preload()
        return batch

    __next__ = next

    def __iter__(self):
        return self

    def __len__(self):
        """Number of batches of the wrapped loader (era5_data/utils_data.py:54-56)."""
        return self.length
