#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: batched PDQ hashing throughput
(BASELINE.json configs[1]: 1024x768 RGB8 images) plus, in the same JSON line, the all-pairs Hamming
grouping of 500k hashes (configs[2]) with its strong scaling over the GPUs of the box, the worst-case
inputs of the two-stage search, and the 1M-file pipeline (configs[4]: 1M hashes x 8 dihedral variants).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the CPU path on the host cores

A "step" is one pass of the hashing hot path over one batch of `--batch` synthetic images.
`value` is device-resident throughput (inputs in HBM before the timed region); `e2e` is the same
metric through the public API with pinned HOST buffers (H2D of the pixels and D2H of the hashes
inside the timed region).  One process per GPU; for N > 1 launch under torchrun.  The Hamming search
is measured on two routes: one process per GPU + torch.distributed (rh_hamming_group_shard), and ONE
process driving all N GPUs through the C ABI alone (rh_group: in-library NCCL, tiles stolen over
NVLink) -- rank 0 runs the second route while the other ranks wait on the host.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_H, IMG_W, IMG_C = 768, 1024, 3            # BASELINE.json configs[1]
ALGO_BYTES_PER_IMAGE = IMG_H * IMG_W * IMG_C + 36   # pixels read + 32 B hash + 4 B quality
HAMMING_N = 500_000                            # BASELINE.json configs[2]
HAMMING_T = 31
POPC_PER_PAIR = 8                              # algorithmic POPC.32 per 256-bit pair
# POPC.32 the tile kernel executes per pair in its hot loop, by variant (hamming.cu)
EXEC_POPC = {1: 1, 2: 1, 3: 2, 4: 3, 5: 2, 6: 2, 7: 3, 0: 4}
EXEC_LOP3 = {1: 4, 2: 2, 3: 5, 4: 6, 5: 5, 6: 6, 7: 8, 0: 16}
VARIANT_NAME = {0: "full 256-bit distance for every pair (16 LOP3 + 4 POPC, carry-save)",
                1: "two-stage: OR lower bound over the first 128 bits (4 LOP3 + 1 POPC) + exact refine of survivors",
                2: "two-stage: OR lower bound over the first 64 bits (2 LOP3 + 1 POPC) + exact refine of survivors",
                3: "two-stage: exact 96-bit prefix distance (2 POPC) + exact refine of survivors",
                4: "two-stage: exact 128-bit prefix distance (3 POPC) + exact refine of survivors",
                5: "two-stage: OR lower bound over the first 160 bits (5 LOP3 + 2 POPC) + exact refine of survivors",
                6: "two-stage: OR lower bound over the first 192 bits (6 LOP3 + 2 POPC) + exact refine of survivors",
                7: "two-stage: OR lower bound over all 256 bits in three groups (8 LOP3 + 3 POPC) + exact refine of survivors"}

# identical in both arms (the driver compares the two `config` objects)
CONFIG = {"workload": "configs[1]: batched PDQ hashing of synthetic 1024x768 RGB8 images",
          "image": [IMG_H, IMG_W, IMG_C]}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="images per step (device-resident arm)")
    ap.add_argument("--e2e-batch", type=int, default=1024, help="images per step (pinned-host arm)")
    ap.add_argument("--host-pool", type=int, default=256, help="distinct images in the pinned host pool")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--hamming-n", type=int, default=HAMMING_N)
    ap.add_argument("--config5-n", type=int, default=1_000_000)
    ap.add_argument("--worstcase-n", type=int, default=200_000)
    ap.add_argument("--skip-hamming", action="store_true")
    ap.add_argument("--skip-worstcase", action="store_true")
    ap.add_argument("--skip-config5", action="store_true")
    ap.add_argument("--skip-config4", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-peaks", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md).  NVML is polled from a
    thread every few milliseconds (the timed region of a short run lasts tens of ms, too short for
    `nvidia-smi -lms`); nvidia-smi is the fallback when pynvml is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self.mask = 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(physical_index(gpu_index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _poll(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            except Exception:
                pass
            self._stop.wait(0.004)

    def start(self):
        if self._h is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()

    def stop(self) -> dict:
        if self._h is not None:
            self._stop.set()
            self._thread.join()
            reasons = sorted(name for name, bit in self.REASONS if self.mask & bit)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.samples), "source": "nvml"}
        # fallback: one nvidia-smi query
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                 capture_output=True, text=True, timeout=20).stdout.strip().split(",")
            f = [x.strip() for x in out]
            reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                         f[2:6]) if v.lower().startswith("active")]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": reasons, "samples": 1,
                    "source": "nvidia-smi (after the timed region)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


def physical_index(gpu_index: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    return int(vis.split(",")[gpu_index]) if vis and vis.split(",")[0].isdigit() else gpu_index


def pin_to_gpu_numa_node(gpu_index: int) -> dict:
    """Bind this process (and, by first touch, the pinned pools it allocates afterwards) to the NUMA node
    the GPU's PCIe root hangs off, so that every rank's H2D traffic stays on its own socket."""
    info = {"node": None, "cpus": None}
    global _ALL_CPUS
    _ALL_CPUS = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_index(gpu_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        info["pci"] = bus
        try:
            node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        except OSError:
            node = -1
        if node < 0:
            # no NUMA node in sysfs (virtualised topology): ask NVML which CPUs are closest to this GPU
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = {64 * wi + b for wi, wd in enumerate(mask) for b in range(64) if (int(wd) >> b) & 1}
            allowed = cpus & os.sched_getaffinity(0)
            if allowed and len(allowed) < len(os.sched_getaffinity(0)):
                os.sched_setaffinity(0, allowed)
                info.update(node="nvml-affinity", cpus=len(allowed))
            else:
                info["nvml_affinity_cpus"] = len(allowed)
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(node=node, cpus=len(allowed))
    except Exception as e:   # not fatal: the numbers are then simply unpinned
        info["error"] = str(e)[:80]
    return info


_ALL_CPUS = None


def unpin_cpus():
    """The CPU baselines use every host core, as the reference's rayon pool would."""
    if _ALL_CPUS:
        try:
            os.sched_setaffinity(0, _ALL_CPUS)
        except OSError:
            pass


def synth_pool_device(torch, count, seed):
    """(count, 768, 1024, 3) uint8 on the current device: blocky low-frequency field + pixel noise
    + per-channel offset, generated in slabs to bound scratch."""
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((count, IMG_H, IMG_W, IMG_C), dtype=torch.uint8, device=dev)
    ys = (torch.arange(IMG_H, device=dev) * 24 // IMG_H)
    xs = (torch.arange(IMG_W, device=dev) * 32 // IMG_W)
    slab = 32
    for s in range(0, count, slab):
        m = min(slab, count - s)
        field = torch.randn((m, 24, 32), generator=g, device=dev) * 50.0
        base = field[:, ys][:, :, xs].unsqueeze(-1)
        noise = torch.randn((m, IMG_H, IMG_W, IMG_C), generator=g, device=dev) * 20.0
        offs = torch.randn((m, 1, 1, IMG_C), generator=g, device=dev) * 10.0
        out[s:s + m] = (base + noise + offs + 128.0).clamp_(0, 255).to(torch.uint8)
    return out


def run_reference(args, emit):
    """--impl reference: the reference's CPU algorithm (oracle port; the Rust crate cannot be
    built here) on all host threads, same metric / config; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from rupphash_b200.synth import synth_images
    oracle.build()
    cores = os.cpu_count() or 1
    # a bounded sample per step: 16 images per host thread, so that thread start-up does not weigh on
    # the rate (64 distinct synthetic images, cycled)
    per_step = max(cores * 16, 128)
    pool = synth_images(64, IMG_H, IMG_W, seed=0xB200)
    imgs = np.ascontiguousarray(pool[np.arange(per_step) % len(pool)])
    for _ in range(args.warmup):
        oracle.pdq_batch(imgs[: max(cores, 8)], threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.pdq_batch(imgs, threads=cores)
    dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "pdq_images_per_sec", "value": v, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(CONFIG),
        "run": {"images_per_step": per_step, "how": "CPU oracle port of pdqhash.rs, one image per task over all host threads"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} images/step x {args.steps} steps"},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------- datasets ----

def dihedral_coeff_dataset(torch, ctx, n, seed):
    """configs[3]/[4] grouping input with REAL dihedral variants: n coefficient blocks (16 x 16 f32, spread
    falling with frequency like a photo's DCT) -> rh_pdq_dihedral_from_coeffs -> 8 variants per file
    (scanner.rs:1615-1628), hash = variant 0.  Planted structure: ~2 % of the files are near-copies of another
    file's coefficients *after one of the 8 dihedral transforms* (sign flips by frequency parity and / or a
    transpose, pdqhash.rs:71-87), so they match one of the source's variants, not its plain hash; 2 % of the
    files are low-confidence.  Returns device tensors (hashes n x 32, variants n x 8 x 32, low_conf n)."""
    from rupphash_b200 import _lib
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    fr = torch.arange(1, 17, device=dev, dtype=torch.float32)
    scale = 40.0 / torch.sqrt(fr[:, None] ** 2 + fr[None, :] ** 2)
    coeffs = torch.randn((n, 16, 16), generator=g, device=dev) * scale
    k = max(2, n // 50)
    perm = torch.randperm(n, generator=g, device=dev)
    src, dst = perm[:k], perm[k:2 * k]
    t = torch.randint(0, 8, (k,), generator=g, device=dev)
    c = coeffs[src].clone()
    odd = (torch.arange(16, device=dev) % 2 == 0).float() * -2.0 + 1.0   # -1 where the frequency r+1 / c+1 is odd
    neg_rows = ((t == 1) | (t == 2) | (t == 5) | (t == 7))[:, None, None]
    neg_cols = ((t == 2) | (t == 3) | (t == 4) | (t == 7))[:, None, None]
    c = torch.where(neg_rows, c * odd[None, :, None], c)
    c = torch.where(neg_cols, c * odd[None, None, :], c)
    tr = ((t == 1) | (t == 3) | (t == 6) | (t == 7))[:, None, None]
    c = torch.where(tr, c.transpose(1, 2), c)
    coeffs[dst] = c + torch.randn(c.shape, generator=g, device=dev) * 0.05 * scale
    coeffs = coeffs.reshape(n, 256).contiguous()
    variants = torch.empty((n, 8, 32), dtype=torch.uint8, device=dev)
    ctx.check(_lib.lib().rh_pdq_dihedral_from_coeffs(ctx.handle, coeffs.data_ptr(), n, variants.data_ptr()))
    hashes = variants[:, 0, :].contiguous()
    low_conf = (torch.rand((n,), generator=g, device=dev) < 0.02).to(torch.uint8)
    del coeffs
    return hashes, variants, low_conf


def correlated_pdq_hashes(torch, ctx, n, seed):
    """Hashes with the bit correlations of real PDQ output: smooth random 64 x 64 buffers (a low-frequency
    field + a little noise) through the real quality / DCT / median tail (rh_pdq_from_buffer64)."""
    from rupphash_b200 import _lib
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    slab = 16384
    for s in range(0, n, slab):
        m = min(slab, n - s)
        low = torch.randn((m, 1, 6, 6), generator=g, device=dev) * 40.0
        buf = torch.nn.functional.interpolate(low, size=(64, 64), mode="bilinear", align_corners=False)[:, 0]
        buf = (buf + torch.randn((m, 64, 64), generator=g, device=dev) * 2.0 + 128.0).contiguous()
        ctx.check(_lib.lib().rh_pdq_from_buffer64(ctx.handle, buf.data_ptr(), m, out[s:s + m].data_ptr(), None, None, None))
    return out


def main():
    args = parse()
    # stdout carries exactly one JSON line: anything a library prints there meanwhile (NCCL's
    # version banner, torchrun notices) is diverted to stderr until the result is ready.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)

    if args.impl == "reference":
        return run_reference(args, emit)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    numa = pin_to_gpu_numa_node(local_rank)   # before any pinned allocation (first touch)

    import torch
    import torch.distributed as dist

    from rupphash_b200 import _lib, pdqhash, scanner
    from rupphash_b200.synth import planted_hashes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; rupphash_b200 has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # host-side rendezvous for the phases in which rank 0 alone drives every GPU: the waiting ranks
        # must not park a spinning NCCL kernel on a GPU rank 0 is using
        host_group = dist.new_group(backend="gloo")
    ctx = _lib.Context(local_rank)
    # a dedicated non-default stream shared by torch, NCCL and the library, so that the CUDA
    # events below are recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_group)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hbm_gbs, peak_src = peaks()
    traffic_per_image = None   # dram read + write bytes per image of the dominant kernel, from the ncu capture
    try:
        traffic_per_image = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[
            "pdq_fused_dram_bytes_per_image"]
    except Exception:
        pass
    B, K, W = args.batch, args.steps, args.warmup

    # ------------------------------------------------------------- PDQ, device resident ----
    pool = synth_pool_device(torch, B, seed=0xB200 + rank)        # B x 2.36 MB >> 126 MB L2
    out_hash = torch.empty((B, 32), dtype=torch.uint8, device="cuda")
    out_q = torch.empty((B,), dtype=torch.float32, device="cuda")
    L = _lib.lib()

    def step_device():
        ctx.check(L.rh_pdq_hash_batch(ctx.handle, pool.data_ptr(), _lib.LAYOUT_RGB8, B, IMG_W, IMG_H, 0, 0,
                                      out_hash.data_ptr(), out_q.data_ptr(), None, None, None))

    for _ in range(W):
        step_device()
    barrier()
    launches0 = ctx.kernel_launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = 0.0
    ev0.record(stream)
    for _ in range(K):
        step_device()
        kernel_ms += ctx.last_kernel_time()[0]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.kernel_launches - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / K
    value = world * B * K / (ms_total * 1e-3)
    kernel_ms_per_step = max_over_ranks(kernel_ms / K)
    achieved_gbs = B * ALGO_BYTES_PER_IMAGE / (kernel_ms_per_step * 1e-3) / 1e9

    # ------------------------------------------------------------- PDQ, other shapes -------
    # the same device-resident bytes viewed as other image shapes (throughput only; the pixels are the pool's):
    # portrait planes take the float-chain fused kernel, 512 x 512 has no pre-downsample (4x fewer source
    # bytes per plane pixel: bound by the chains, not by HBM)
    shapes = {}
    for name, (h, w, kernel) in {"portrait_1024x768": (1024, 768, "pdq_float_kernel"),
                                 "square_512x512": (512, 512, "pdq_fused_kernel"),
                                 "portrait_512x384": (512, 384, "pdq_float_kernel"),
                                 "small_256x256": (256, 256, "pdq_float_kernel")}.items():
        m = pool.numel() // (h * w * 3)
        view = pool.reshape(-1)[: m * h * w * 3].reshape(m, h, w, 3)
        oh = torch.empty((m, 32), dtype=torch.uint8, device="cuda")
        oq = torch.empty((m,), dtype=torch.float32, device="cuda")
        times = []
        for rep in range(4):
            ctx.check(L.rh_pdq_hash_batch(ctx.handle, view.data_ptr(), _lib.LAYOUT_RGB8, m, w, h, 0, 0, oh.data_ptr(),
                                          oq.data_ptr(), None, None, None))
            if rep:
                times.append(ctx.last_kernel_time()[0])
        ms = max_over_ranks(float(np.median(times)))
        rate = m / (ms * 1e-3)
        shapes[name] = {"images_per_s_per_gpu": rate, "images": m, "kernel": kernel,
                        "hbm_roofline_frac": rate * (3 * h * w + 36) / 1e9 / hbm_gbs}
        del oh, oq

    # ------------------------------------------------------------- configs[3]: 512 x 512, 8 variants + pHash ----
    config4 = None
    if rank == 0 and not args.skip_config4:
        config4 = bench_config4(torch, _lib, pdqhash, scanner, ctx, pool, hbm_gbs, args)

    # ------------------------------------------------------------- PDQ, end to end ---------
    Be = min(args.e2e_batch, B)
    hp = min(args.host_pool, Be)
    host_pool = torch.empty((Be, IMG_H, IMG_W, IMG_C), dtype=torch.uint8, pin_memory=True)
    for s in range(0, Be, hp):
        m = min(hp, Be - s)
        host_pool[s:s + m].copy_(pool[:m])
    torch.cuda.synchronize()
    host_np = host_pool.numpy()
    e2e_steps = max(2, min(K, 5))
    pdqhash.hash_batch(host_np, ctx=ctx)      # warm-up (allocates staging)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = pdqhash.hash_batch(host_np, ctx=ctx)   # H2D pixels + kernels + D2H hash/quality/valid
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * Be * e2e_steps / e2e_wall
    h2d_bytes = Be * IMG_H * IMG_W * IMG_C
    d2h_bytes = Be * (32 + 4 + 1)
    pool_idx = np.concatenate([np.arange(min(hp, Be - s)) for s in range(0, Be, hp)])
    same = bool(np.array_equal(res["hash"], out_hash.cpu().numpy()[pool_idx]))

    line = {
        "metric": "pdq_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(CONFIG),
        "run": {"images_per_step_per_gpu": B, "pool": "device-resident pool cycled per step",
                "l2_policy": f"inputs larger than L2: {B * IMG_H * IMG_W * IMG_C / 1e9:.2f} GB read per step",
                "parallelism": f"independent batches per GPU x{world} (no collective)", "numa": numa},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "images_per_step_per_gpu": Be, "steps": e2e_steps,
                "matches_device_resident_hashes": same},
        "gpu_launches": int(launches),
        "pdq_shapes": shapes,
        "config4": config4,
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_gbs, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_gbs,
                     "traffic": traffic_per_image * B if traffic_per_image else None,
                     "traffic_source": "profiles/ncu_traffic.json (ncu dram read + write bytes per image x images "
                                       "per launch)" if traffic_per_image else None,
                     "peak_source": peak_src,
                     "kernel_ms_per_step": kernel_ms_per_step,
                     "algorithmic_bytes_per_image": ALGO_BYTES_PER_IMAGE},
    }

    # ------------------------------------------------------------- PDQ through the scanner-style feeder ----
    # the whole host side of the reference's scan loop restructured into batches (scanner.rs:1202-1521): a pool of
    # "decode" workers -- here a memcpy out of a PAGEABLE pool, i.e. the cost of handing a decoded image over --
    # copies into page-locked staging batches, one submitter keeps two batches in flight (rh_pdq_hash_batch_async)
    if rank == 0:
        n_feed = 2048
        pageable = [np.array(host_np[k % Be]) for k in range(64)]      # 64 distinct pageable images, cycled
        workers = max(2, min(8, ((os.cpu_count() or 4) // 2) // max(1, world)))   # more copy threads than that fight the DMA reads
        fdr = scanner.Feeder()
        scanner.hash_files_batched(range(n_feed), decode=lambda k: pageable[k % 64], workers=workers, batch_size=256,
                                   want_coeffs=False, ctx=ctx, feeder=fdr)   # warm-up: page-locks the staging buffers
        t0 = time.perf_counter()
        fed = scanner.hash_files_batched(range(n_feed), decode=lambda k: pageable[k % 64], workers=workers, batch_size=256,
                                         want_coeffs=False, ctx=ctx, feeder=fdr)
        dt = time.perf_counter() - t0
        fdr.close()
        ok = all(np.array_equal(fed[k]["hash"], res["hash"][k % 64 % Be]) for k in range(0, n_feed, 97))
        line["e2e"]["feeder"] = {"value": n_feed / dt, "unit": "images/s", "images": n_feed, "decode_workers": workers,
                                 "what": "scanner.hash_files_batched: memcpy 'decode' from pageable memory -> pinned staging "
                                         "-> async H2D + kernels + D2H, results per file (1 GPU, rank 0)",
                                 "matches_hash_batch": bool(ok)}
        del pageable, fed

    # ------------------------------------------------------------- integer / copy peaks -----
    # every rank measures at the same time: the pinned-H2D figure is then the CONCURRENT rate each GPU gets
    # while its neighbours copy too -- the denominator of the N-GPU e2e number (a solo figure cannot tell a
    # shared-uplink limit from a software problem)
    pk = None
    if not args.skip_peaks:
        barrier()
        pk = ctx.measure_peaks()
        # sustained figure: after a barrier every rank copies its own 1 GiB of pinned memory four times
        hb = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
        db = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
        db.copy_(hb, non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(4):
            db.copy_(hb, non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize()
        sustained = 4 * (1 << 30) / (e0.elapsed_time(e1) * 1e-3) / 1e9
        h2d_all = [sustained]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, sustained, group=host_group)
            h2d_all = [float(x) for x in gathered]
        del hb, db
        line["measured_peaks"] = dict(pk, h2d_gbs_concurrent_per_rank=h2d_all, h2d_gbs_concurrent_sum=sum(h2d_all))
        if sum(h2d_all) > 0:
            line["e2e"]["h2d_roofline_frac"] = e2e_value * IMG_H * IMG_W * IMG_C / 1e9 / sum(h2d_all)
            line["e2e"]["h2d_roofline_note"] = ("whole-job pixel bytes/s over the sum of the pinned H2D rates all ranks "
                                                "measured at the same time")

    # ------------------------------------------------------------- Hamming grouping --------
    labels = None
    if not args.skip_hamming:
        n = args.hamming_n
        hashes, low_conf = planted_hashes(n, seed=0xB200, n_clusters=5000, identical_block=1000,
                                          threshold=HAMMING_T)
        d_hashes = torch.from_numpy(hashes).cuda()
        d_lc = torch.from_numpy(low_conf).cuda()
        pairs = n * (n - 1) // 2

        def group_device():
            if world == 1:
                lab, cnt = scanner.group_labels(d_hashes, HAMMING_T, low_conf=d_lc, ctx=ctx)
                return lab, cnt, ctx.last_kernel_time()[0]
            parent, cnt = scanner.group_shard(d_hashes, HAMMING_T, rank, world, low_conf=d_lc, ctx=ctx)
            kms = ctx.last_kernel_time()[0]
            gathered = torch.empty((world, n), dtype=parent.dtype, device="cuda")
            dist.all_gather_into_tensor(gathered, parent)
            total = torch.tensor([cnt], dtype=torch.int64, device="cuda")
            dist.all_reduce(total)
            lab = scanner.merge_forests(gathered, ctx)
            return lab, int(total.item()), kms

        # route A: one process per GPU (this process is one rank), static tile ownership
        group_device()
        barrier()
        reps = 3
        kms_sum = 0.0
        t0 = time.perf_counter()
        for _ in range(reps):
            labels, edges, kms = group_device()
            kms_sum += kms
        barrier()
        wall_dev = max_over_ranks((time.perf_counter() - t0) / reps)
        tile_ms = max_over_ranks(kms_sum / reps)
        pf_used = ctx.hamming_last_variant()      # chosen on the device from the sampled selectivity
        # the full-distance variant (8 words per pair, no prefix filter) for reference
        ctx.set_option("hamming.prefilter", 0)
        group_device()
        _, _, full_ms = group_device()
        full_ms = max_over_ranks(full_ms)
        ctx.set_option("hamming.prefilter", -1)
        ham = {"metric": "hamming_pairs_per_sec", "n_hashes": n, "threshold": HAMMING_T, "pairs": pairs,
               "value": pairs / wall_dev, "unit": "pairs/s", "tile_kernel_ms": tile_ms,
               "tile_kernel_pairs_per_s": pairs / (tile_ms * 1e-3) if tile_ms > 0 else None,
               "group_wall_ms_device_resident": wall_dev * 1e3, "edges": int(edges), "n_gpus": world,
               "route": "one process per GPU (rh_hamming_group / rh_hamming_group_shard + torch.distributed)",
               "kernel_variant": VARIANT_NAME[pf_used], "kernel_variant_id": pf_used,
               "full_distance_variant": {"tile_kernel_ms": full_ms,
                                         "pairs_per_s": pairs / (full_ms * 1e-3) if full_ms > 0 else None,
                                         "note": "every pair gets all 8 words: 16 LOP3 + 4 POPC (carry-save)"},
               "exchange": "none (1 GPU)" if world == 1 else "NCCL all-gather of n x u32 forests + all-reduce of edge counts"}
        if pk:
            per_gpu = pairs / world / (tile_ms * 1e-3)
            ham["roofline"] = {
                "bound": "int-pipe (POPC.32)", "unit": "POPC/s", "peak": pk["popc_per_s"],
                "achieved": per_gpu * EXEC_POPC[pf_used], "frac": per_gpu * EXEC_POPC[pf_used] / pk["popc_per_s"],
                "definition": f"POPC.32 the kernel EXECUTES in its hot loop ({EXEC_POPC[pf_used]} per pair for the chosen variant) over "
                              "the measured POPC issue rate, per GPU; ncu: sm__inst_executed_pipe_xu "
                              "(profiles/ncu_hamming_tiles_*)",
                "executed_lop3_frac": per_gpu * EXEC_LOP3[pf_used] / pk["lop3_per_s"],
                "algorithmic_speedup": POPC_PER_PAIR / EXEC_POPC[pf_used],
                "algorithmic_popc_rate_over_peak": per_gpu * POPC_PER_PAIR / pk["popc_per_s"],
                "algorithmic_note": "SURVEY 8d counts 8 POPC per pair; the exact two-stage search skips most of them for "
                                    "> 99.9 % of the pairs, so the algorithmic rate exceeds the pipe peak -- that ratio is a "
                                    "speed-up of the algorithm, not a roofline fraction",
                "full_distance_variant_frac": (pairs / world / (full_ms * 1e-3)) * EXEC_POPC[0] / pk["popc_per_s"]
                if full_ms > 0 else None}
        line["hamming"] = ham

    # ------- rank 0 alone drives the GPUs from here on (in-library multi-GPU); the others wait on the host ----
    n_dev = min(world, torch.cuda.device_count())
    host_barrier()
    if rank == 0 and not args.skip_hamming:
        line["hamming"]["in_library"] = bench_group_route(torch, _lib, scanner, hashes, low_conf, HAMMING_T, n_dev,
                                                          labels, pairs)
        line["hamming_strong_scaling_efficiency"] = line["hamming"]["in_library"].get("strong_scaling_efficiency")
    if rank == 0 and not args.skip_worstcase and not args.skip_hamming:
        line["hamming"]["worst_case"] = bench_worst_case(torch, _lib, scanner, ctx, pk, args.worstcase_n)
    if rank == 0 and not args.skip_config5:
        line["config5"] = bench_config5(torch, _lib, scanner, ctx, args, n_dev, value, e2e_value, world)

    # ------------------------------------------------------------- CPU baseline (rank 0) ---
    if rank == 0 and not args.skip_cpu:
        import oracle
        oracle.build()
        unpin_cpus()
        cores = os.cpu_count() or 1
        ns = args.cpu_sample or max(2 * cores, min(1024, Be))
        ns = min(ns, Be)
        t0 = time.perf_counter()
        ref = oracle.pdq_batch(host_np[:ns], threads=cores)
        dt = time.perf_counter() - t0
        ident = int((ref["hash"] == res["hash"][:ns]).all(axis=1).sum())
        line["cpu_baseline"] = {"value": ns / dt, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"first {ns} images of the step's batch, oracle/ (C port of pdqhash.rs), "
                                          f"{cores} threads, {dt:.2f} s",
                                "parity_identical_hashes": ident, "parity_checked": ns}
        if not args.skip_hamming:
            t0 = time.perf_counter()
            ref_labels, ref_cnt, _ = oracle.group_generic(hashes, HAMMING_T, low_conf=low_conf, threads=cores)
            dt = time.perf_counter() - t0
            lab = labels.cpu().numpy().view(np.uint32) if torch.is_tensor(labels) else np.asarray(labels)
            line["hamming"]["cpu_baseline"] = {
                "group_wall_ms": dt * 1e3, "pairs_equivalent_per_s": pairs / dt, "cores": cores, "kind": "port",
                "sample": f"full {n} hashes, oracle MIH probing + sequential union-find (scanner.rs:1673-1807)",
                "labels_identical": bool(np.array_equal(lab, ref_labels)), "edges_identical": bool(ref_cnt == edges),
                "in_library_labels_identical": bool(np.array_equal(line["hamming"]["in_library"]["_labels"], ref_labels))}
    if rank == 0 and "hamming" in line:
        line["hamming"].get("in_library", {}).pop("_labels", None)
    host_barrier()
    if rank == 0:
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def bench_config4(torch, _lib, pdqhash, scanner, ctx, pool, hbm_gbs, args):
    """configs[3]: PDQ with all 8 dihedral variants (+ coefficients) and the 64-bit pHash over 512 x 512 images, then
    grouping across the variants.  The images are the device-resident pool viewed as 512 x 512 (throughput); parity
    samples go through the CPU oracle.  The 1M x 8 PDQ grouping itself is the `config5` section (same input shape);
    here: u64 grouping of the pHashes with phash::generate_dihedral_hashes variants -- offered for symmetry with
    `impl HammingHash for u64` (hamminghash.rs:23-41); NO reference caller groups u64 hashes."""
    from rupphash_b200 import phash
    L = _lib.lib()
    h = w = 512
    m = min(pool.numel() // (h * w * 3), 8192)
    view = pool.reshape(-1)[: m * h * w * 3].reshape(m, h, w, 3)
    oh = torch.empty((m, 32), dtype=torch.uint8, device="cuda")
    oq = torch.empty((m,), dtype=torch.float32, device="cuda")
    oc = torch.empty((m, 256), dtype=torch.float32, device="cuda")
    od = torch.empty((m, 8, 32), dtype=torch.uint8, device="cuda")
    ms = []
    for rep in range(4):
        ctx.check(L.rh_pdq_hash_batch(ctx.handle, view.data_ptr(), _lib.LAYOUT_RGB8, m, w, h, 0, 0, oh.data_ptr(), oq.data_ptr(),
                                      oc.data_ptr(), od.data_ptr(), None))
        ms.append(ctx.last_kernel_time()[0])
    t_pdq = float(np.median(ms[1:])) * 1e-3
    ph = torch.empty((m,), dtype=torch.int64, device="cuda")
    pd8 = torch.empty((m, 8), dtype=torch.int64, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tp = []
    for rep in range(4):
        ev0.record(torch.cuda.current_stream())
        ctx.check(L.rh_phash_batch(ctx.handle, view.data_ptr(), _lib.LAYOUT_RGB8, m, w, h, 0, 0, ph.data_ptr(), pd8.data_ptr()))
        ev1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        tp.append(ev0.elapsed_time(ev1))
    t_ph = float(np.median(tp[1:])) * 1e-3
    out = {"image": [h, w, 3], "images": m,
           "pdq_8_variants_coeffs": {"images_per_s": m / t_pdq,
                                     "hbm_roofline_frac": (m / t_pdq) * (3 * h * w + 36 + 1024 + 224) / 1e9 / hbm_gbs,
                                     "outputs": "hash + quality + 256 coefficients + 8 dihedral hashes per image"},
           "phash_8_variants": {"images_per_s": m / t_ph, "hbm_roofline_frac": (m / t_ph) * (3 * h * w + 72) / 1e9 / hbm_gbs,
                                "note": "phash.rs:48-83 on the device; parity with the `image` / `rustdct` crates unpinned"}}
    if not args.skip_cpu:
        import oracle
        oracle.build()
        k = 32
        host = view[:k].cpu().numpy()
        ref = oracle.pdq_batch(host, threads=min(k, os.cpu_count() or 1), want_coeffs=True, want_dihedral=True)
        out["pdq_8_variants_coeffs"]["parity_vs_oracle"] = {
            "images": k, "hash": bool(np.array_equal(oh[:k].cpu().numpy(), ref["hash"])),
            "dihedral": bool(np.array_equal(od[:k].cpu().numpy(), ref["dihedral"])),
            "coefficient_bits": bool(np.array_equal(oc[:k].cpu().numpy().view(np.uint32), ref["coeffs"].view(np.uint32)))}
        want = np.array([oracle.phash_image(host[i])[0] for i in range(k)], np.uint64)
        got = ph[:k].cpu().numpy().view(np.uint64)
        out["phash_8_variants"]["parity_vs_oracle"] = {
            "images": k, "hash": bool(np.array_equal(got, want)),
            "dihedral": bool(all(list(pd8[i].cpu().numpy().view(np.uint64)) == oracle.phash_dihedral(int(want[i])) for i in range(k)))}
    # u64 grouping across the 8 pHash variants: 1M synthetic 64-bit hashes with planted near-duplicates
    n = 1_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(64)
    base = torch.randint(-(1 << 62), 1 << 62, (n,), generator=g, device="cuda", dtype=torch.int64)
    src = torch.randint(0, n, (n // 50,), generator=g, device="cuda")
    dst = torch.randint(0, n, (n // 50,), generator=g, device="cuda")
    flips = torch.randint(0, 64, (n // 50, 3), generator=g, device="cuda")
    mask = (torch.ones_like(flips) << flips).sum(dim=1)     # up to 3 bits flipped (a repeated position counts twice: fine)
    base[dst] = base[src] ^ mask
    hv = base.cpu().numpy().view(np.uint64)
    var = np.empty((n, 8), np.uint64)
    # the 8 bit-level variants of every hash on the host (pure bit permutations, phash.rs:242-255), vectorised
    bits = ((hv[:, None] >> np.arange(63, -1, -1, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8).reshape(n, 8, 8)
    xs, ys = np.meshgrid(np.arange(8), np.arange(8))

    def pack(bm):
        return (bm.reshape(n, 64).astype(np.uint64) << np.arange(63, -1, -1, dtype=np.uint64)[None, :]).sum(axis=1, dtype=np.uint64)
    odd_x, odd_y, odd_xy = (xs % 2 == 1), (ys % 2 == 1), ((xs + ys) % 2 == 1)

    def r90(bm):
        return np.where(odd_x[None], 1 - bm.transpose(0, 2, 1), bm.transpose(0, 2, 1))

    def r180(bm):
        return np.where(odd_xy[None], 1 - bm, bm)

    def r270(bm):
        return np.where(odd_y[None], 1 - bm.transpose(0, 2, 1), bm.transpose(0, 2, 1))
    fl = np.where(odd_x[None], 1 - bits, bits)
    for kk, bm in enumerate((bits, r90(bits), r180(bits), r270(bits), fl, r90(fl), r180(fl), r270(fl))):
        var[:, kk] = pack(bm)
    ok_variants = all(list(var[i]) == phash.generate_dihedral_hashes(int(hv[i])) for i in range(0, n, n // 16))
    d_h = torch.from_numpy(hv.view(np.int64)).cuda()
    d_v = torch.from_numpy(var.view(np.int64)).cuda()
    lab = torch.empty(n, dtype=torch.int32, device="cuda")
    import ctypes as C
    cnt = C.c_uint64()
    walls = []
    ctx.check(L.rh_hamming_group_u64(ctx.handle, d_h.data_ptr(), None, d_v.data_ptr(), None, None, 20000, 10, lab.data_ptr(),
                                     C.byref(cnt)))    # warm-up on a prefix
    for rep in range(1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.check(L.rh_hamming_group_u64(ctx.handle, d_h.data_ptr(), None, d_v.data_ptr(), None, None, n, 10, lab.data_ptr(),
                                         C.byref(cnt)))
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
    labn = lab.cpu().numpy()
    out["phash_grouping_u64"] = {"n_hashes": n, "variants": 8, "similarity": 10, "pairs": 8 * n * (n - 1) // 2,
                                 "group_wall_ms": min(walls) * 1e3, "pairs_per_s": 8 * n * (n - 1) / 2 / min(walls),
                                 "edges": int(cnt.value), "groups": int((np.bincount(labn, minlength=n) > 1).sum()),
                                 "variant_rule_matches_rh_phash_dihedral": bool(ok_variants),
                                 "note": "no reference caller groups u64 hashes; parity of rh_hamming_group_u64 vs brute force "
                                         "is tests/test_gpu_hamming.py::test_u64_group_and_find_groups_kats"}
    return out


def pinned(torch, arr):
    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t, t.numpy()


def time_group(torch, _lib, scanner, n_dev, flags, hashes, similarity, variants=None, low_conf=None, reps=3):
    """rh_hamming_group_multi from pinned host buffers to labels in pinned host memory, on n_dev GPUs."""
    g = _lib.Group(n_dev=n_dev, flags=flags)
    try:
        out_t = torch.empty(len(hashes), dtype=torch.int32).pin_memory()
        out = out_t.numpy().view(np.uint32)
        scanner.group_labels_multi(g, hashes, similarity, variants=variants, low_conf=low_conf, out=out)   # warm-up
        walls, tiles, sums = [], [], []
        for _ in range(reps):
            t0 = time.perf_counter()
            _, cnt = scanner.group_labels_multi(g, hashes, similarity, variants=variants, low_conf=low_conf, out=out)
            walls.append((time.perf_counter() - t0) * 1e3)
            t = g.last_times()
            tiles.append(t["tile_ms_max"])
            sums.append(t["tile_ms_sum"])
        info = g.info()
        info["gpu0_timeline"] = {k: round(v, 4) for k, v in t.items() if k.startswith("gpu0_")}
        return {"wall_ms": float(np.median(walls)), "tile_ms_max": float(np.median(tiles)),
                "tile_ms_sum": float(np.median(sums)), "edges": int(cnt), "info": info, "labels": out.copy()}
    finally:
        g.close()


def bench_group_route(torch, _lib, scanner, hashes, low_conf, similarity, n_dev, route_a_labels, pairs):
    """configs[2] through ONE process and the C ABI alone: pinned host hashes -> labels in host memory on
    1 and on n_dev GPUs; strong-scaling efficiency = t(1) / (n_dev * t(n_dev))."""
    _, h = pinned(torch, hashes)
    _, lc = pinned(torch, low_conf)
    one = time_group(torch, _lib, scanner, 1, 0, h, similarity, low_conf=lc)
    res = {"route": "one process, rh_group + rh_hamming_group_multi (in-library NCCL over NVLink; tile t on GPU t mod N); "
                    "pinned host hashes in, labels in host memory out",
           "n_gpus": n_dev, "group_wall_ms_1gpu": one["wall_ms"], "tile_kernel_ms_1gpu": one["tile_ms_max"],
           "edges": one["edges"], "_labels": one["labels"]}
    if n_dev > 1:
        multi = time_group(torch, _lib, scanner, n_dev, 0, h, similarity, low_conf=lc)
        steal = time_group(torch, _lib, scanner, n_dev, _lib.GROUP_STEAL_TILES, h, similarity, low_conf=lc)
        p2p = time_group(torch, _lib, scanner, n_dev, _lib.GROUP_NO_NCCL, h, similarity, low_conf=lc)
        res.update({
            "group_wall_ms": multi["wall_ms"], "tile_kernel_ms_slowest_gpu": multi["tile_ms_max"],
            "tile_kernel_gpu_ms_sum": multi["tile_ms_sum"],
            "pairs_per_s": pairs / (multi["wall_ms"] * 1e-3),
            "strong_scaling_efficiency": one["wall_ms"] / (n_dev * multi["wall_ms"]),
            "strong_scaling_efficiency_tile_kernel": one["tile_ms_max"] / (n_dev * multi["tile_ms_max"]),
            "nccl_version": multi["info"]["nccl_version"], "work_stealing": multi["info"]["work_stealing"],
            "gpu0_timeline_ms": multi["info"].get("gpu0_timeline"), "gpu0_timeline_ms_1gpu": one["info"].get("gpu0_timeline"),
            "labels_identical_to_1gpu": bool(np.array_equal(multi["labels"], one["labels"])),
            "edges_identical_to_1gpu": multi["edges"] == one["edges"],
            "work_stealing_tiles": {"group_wall_ms": steal["wall_ms"], "tile_kernel_ms_slowest_gpu": steal["tile_ms_max"],
                                    "labels_identical": bool(np.array_equal(steal["labels"], one["labels"]))},
            "peer_copy_exchange": {"group_wall_ms": p2p["wall_ms"],
                                   "labels_identical": bool(np.array_equal(p2p["labels"], one["labels"]))},
        })
    else:
        res.update({"group_wall_ms": one["wall_ms"], "pairs_per_s": pairs / (one["wall_ms"] * 1e-3),
                    "strong_scaling_efficiency": 1.0})
    if route_a_labels is not None:
        a = route_a_labels.cpu().numpy().view(np.uint32) if torch.is_tensor(route_a_labels) else np.asarray(route_a_labels)
        res["labels_identical_to_per_process_route"] = bool(np.array_equal(a, one["labels"]))
    return res


def bench_worst_case(torch, _lib, scanner, ctx, pk, n):
    """The prefix filter's selectivity is a property of the input.  Three inputs, each searched with the
    two-stage kernels (PF = 3 .. 7) and the full-distance kernel (PF = 0) at similarity 31, plus the reference's
    default setting (similarity 40, 8 variants per file): tile-kernel time and executed-POPC fraction."""
    from rupphash_b200.synth import planted_hashes, random_variants
    rng = np.random.default_rng(0xADE)
    uni, _ = planted_hashes(n, seed=0xB201, n_clusters=2000, identical_block=200, threshold=HAMMING_T)
    adv = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    adv[:, :12] = adv[0, :12]          # every hash shares its first 96 bits: the prefix stage rejects nothing
    sets = {"uniform_planted": torch.from_numpy(uni).cuda(),
            "pdq_of_smooth_images": correlated_pdq_hashes(torch, ctx, n, seed=0xC0FE),
            "adversarial_shared_96bit_prefix": torch.from_numpy(adv).cuda()}
    pairs = n * (n - 1) // 2
    out = {"n_hashes": n, "pairs": pairs, "similarity": HAMMING_T}
    for name, d in sets.items():
        row = {}
        ref = None
        for pf in (3, 5, 6, 7, 4, 0):
            ctx.set_option("hamming.prefilter", pf)
            scanner.group_labels(d, HAMMING_T, ctx=ctx)
            lab, cnt = scanner.group_labels(d, HAMMING_T, ctx=ctx)
            ms = ctx.last_kernel_time()[0]
            lab = lab.cpu().numpy()
            if ref is None:
                ref = (lab, cnt)
            row[f"pf{pf}"] = {"tile_ms": ms, "pairs_per_s": pairs / (ms * 1e-3),
                              "executed_popc_frac_nominal": (pairs / (ms * 1e-3)) * EXEC_POPC[pf] / pk["popc_per_s"] if pk else None,
                              "same_result_as_pf3": bool(cnt == ref[1] and np.array_equal(lab, ref[0]))}
        ctx.set_option("hamming.prefilter", -1)
        scanner.group_labels(d, HAMMING_T, ctx=ctx)
        lab, cnt = scanner.group_labels(d, HAMMING_T, ctx=ctx)
        row["chosen_on_device"] = {"kernel_variant_id": ctx.hamming_last_variant(), "tile_ms": ctx.last_kernel_time()[0],
                                   "same_result_as_pf3": bool(cnt == ref[1] and np.array_equal(lab.cpu().numpy(), ref[0]))}
        row["edges"] = int(ref[1])
        out[name] = row
    # the reference's default: --similarity 40, 8 dihedral variants per file (README.md:13)
    m = n // 2
    var = torch.from_numpy(random_variants(uni[:m], seed=5)).cuda()
    d = sets["uniform_planted"][:m].contiguous()
    for sim in (31, 40, 63):
        scanner.group_labels(d, sim, variants=var, ctx=ctx)
        scanner.group_labels(d, sim, variants=var, ctx=ctx)
        ms = ctx.last_kernel_time()[0]
        pr = 8 * m * (m - 1) // 2
        pf = ctx.hamming_last_variant()
        out[f"variants8_similarity{sim}"] = {"n_hashes": m, "pairs": pr, "tile_ms": ms, "pairs_per_s": pr / (ms * 1e-3),
                                             "kernel_variant": f"pf{pf}",
                                             "executed_popc_frac": (pr / (ms * 1e-3)) * EXEC_POPC[pf] / pk["popc_per_s"] if pk else None}
    return out


def bench_config5(torch, _lib, scanner, ctx, args, n_dev, pdq_value, pdq_e2e_value, world):
    """configs[4]: the 1M-file pipeline in the reference's two phases (scanner.rs:1542-1559 prints hash-phase and
    group-phase wall time).  Hash phase: the measured PDQ rates applied to 1M images (1M distinct 1024x768 images
    are 2.36 TB, more than HBM or host RAM hold; the pool is cycled).  Group phase: 1M files x 8 REAL dihedral
    variants (from coefficients, scanner.rs:1615-1628) at similarity 31 and 40 (the reference's default), one
    process driving all GPUs through the C ABI, pinned host buffers in, labels in host memory out; the CPU MIH
    port beside it on a bounded sample of the query files."""
    n = args.config5_n
    d_h, d_v, d_l = dihedral_coeff_dataset(torch, ctx, n, seed=0x5EED)
    th, h = pinned(torch, d_h.cpu().numpy())
    tv, v = pinned(torch, d_v.cpu().numpy())
    tl, lc = pinned(torch, d_l.cpu().numpy())
    del d_h, d_v, d_l
    torch.cuda.empty_cache()
    pairs = 8 * n * (n - 1) // 2
    out = {"n_files": n, "query_rows": 8 * n, "pairs": pairs, "n_gpus": n_dev,
           "hash_phase": {"images": n, "device_resident_s": n / pdq_value, "pinned_host_s": n / pdq_e2e_value,
                          "note": f"1M images at the rates measured above on {world} GPU(s) (value / e2e of this line)"},
           "route": "one process, rh_group + rh_hamming_group_multi"}
    cores = os.cpu_count() or 1
    for sim in (31, 40):
        one = time_group(torch, _lib, scanner, 1, 0, h, sim, variants=v, low_conf=lc, reps=1)
        row = {"group_wall_ms_1gpu": one["wall_ms"], "tile_kernel_ms_1gpu": one["tile_ms_max"], "edges": one["edges"],
               "pairs_per_s_1gpu": pairs / (one["wall_ms"] * 1e-3)}
        lab = one["labels"]
        row["groups"] = int((np.bincount(lab, minlength=n) > 1).sum())
        if n_dev > 1:
            multi = time_group(torch, _lib, scanner, n_dev, 0, h, sim, variants=v, low_conf=lc, reps=2)
            row.update({"group_wall_ms": multi["wall_ms"], "tile_kernel_ms_slowest_gpu": multi["tile_ms_max"],
                        "pairs_per_s": pairs / (multi["wall_ms"] * 1e-3),
                        "strong_scaling_efficiency": one["wall_ms"] / (n_dev * multi["wall_ms"]),
                        "labels_identical_to_1gpu": bool(np.array_equal(multi["labels"], lab)),
                        "edges_identical_to_1gpu": multi["edges"] == one["edges"]})
        else:
            row.update({"group_wall_ms": one["wall_ms"], "pairs_per_s": pairs / (one["wall_ms"] * 1e-3)})
        if not args.skip_cpu:
            import oracle
            oracle.build()
            unpin_cpus()
            # bounded sample: the first m files of EVERY 2000-file chunk of query files against the full index
            # (many small work units keep every host thread busy; whole chunks at a stride left most idle)
            m = 200 if sim <= 31 else 16
            t0 = time.perf_counter()
            oracle.group_generic_sampled(h, sim, 1, variants=v, low_conf=lc, threads=cores, sample_files=m)
            dt = time.perf_counter() - t0
            t1 = time.perf_counter()
            oracle.MIHIndex(h)                       # index build alone (not scaled)
            t_index = time.perf_counter() - t1
            est = t_index + (dt - t_index) * 2000.0 / m
            row["cpu_baseline"] = {"estimated_group_wall_s": est, "cores": cores, "kind": "port",
                                   "sample": f"the first {m} of every 2000 query files probed against the full MIH index "
                                             f"({dt:.1f} s measured, index build {t_index:.2f} s not scaled; "
                                             f"tools/config5_check.py runs the full search)",
                                   "speedup_vs_cpu": est / (row["group_wall_ms"] * 1e-3)}
        out[f"similarity{sim}"] = row
    out["total_s_device_resident_hash_plus_group40"] = out["hash_phase"]["device_resident_s"] + \
        out["similarity40"]["group_wall_ms"] * 1e-3
    return out


if __name__ == "__main__":
    main()
