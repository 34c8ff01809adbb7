#!/usr/bin/env python
"""bench.py — forecast-windows/sec of the StemGNN hot path on B200 (BASELINE.json metric).

A "step" is ONE pass of the hot path (Model.forward in eval(), no_grad — SURVEY.md §8(d)) over one
batch of synthetic windows at the north-star shape (B,N,W,H)=(32,358,12,3) per GPU (weak scaling:
every rank owns its own batch; the eval path has no collective).

    python bench.py [--gpus N --steps K --warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                      # the reference's CPU path (oracle port)

Prints ONE JSON line.  Keys (see the task contract):
  value / ms_per_step   device-timed (CUDA events, L2 flushed between steps), inputs resident in HBM
  e2e                   same metric through the public API with HOST buffers: pinned x -> H2D ->
                        Model.forward -> D2H of the forecast, all inside the timed region
  roofline              the dominant kernel (GRU recurrence), timed live with CUDA events via the
                        library's stemgnn_profile_gru hook; flops = SURVEY.md §8(d) GRU terms
  cpu_baseline          oracle/torch_port.py (the reference's ATen op sequence) on the host cores
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, N, W, H, MULTI = 32, 358, 12, 3, 5
WORKLOAD = "synthetic N=358 W=12 H=3 batch=32 fp32 eval forward (BASELINE.json configs[1])"
METRIC = "forecast-windows/sec (B,N,W)=(32,358,12)"
FORWARD_FLOPS = 28140101688      # SURVEY.md §8(d) dead-work-free count at (32,358,12,3); = oracle.forward_flops


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst)"
    except Exception:
        return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.  NVML is polled from a background thread
    every ~2 ms (the timed region is only tens of milliseconds long, too short for `nvidia-smi -lms`); falls back
    to one nvidia-smi query when pynvml is unavailable."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, index):
        import threading
        self.index, self.samples, self.reason_bits, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indices follow the physical order; honour CUDA_VISIBLE_DEVICES when it lists integers
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except (ValueError, IndexError):
                    phys = index
            self._nv, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        except Exception:
            self._nv = None

    def _poll(self):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            sm = sorted(self.samples)
            reasons = sorted(k for k, bit in self.REASONS.items() if self.reason_bits & bit)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml polled every 2 ms inside the timed region"}
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                  "-i", str(self.index)], capture_output=True, text=True, timeout=10).stdout
            f = [v.strip() for v in out.strip().split(",")]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": [], "samples": 1,
                    "source": "nvidia-smi, one query right after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}


def cpu_reference_forward(steps, warmup):
    """The reference's CPU path: oracle/torch_port.model_forward (same ATen ops as base_model.py) on the
    host cores.  torchrun pins OMP_NUM_THREADS=1 and "all cores" is not the fastest setting on a
    128-thread host, so the thread count is picked as the best of a short sweep (the reference gets its
    best shot).  Returns (windows_per_s, ms_per_step, threads)."""
    import torch
    from oracle import torch_port as tp
    p = tp.synthetic_params(N, W, H, MULTI, seed=0)
    x, _ = tp.synthetic_batch(B, N, W, H)
    ncpu = os.cpu_count() or 8
    cands = sorted({c for c in (8, 16, 32, 64, ncpu // 2, ncpu) if 1 <= c <= ncpu})
    best, best_t = None, None
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            tp.model_forward(x, p)
            t0 = time.perf_counter()
            tp.model_forward(x, p)
            dt = time.perf_counter() - t0
            if best_t is None or dt < best_t:
                best, best_t = c, dt
        torch.set_num_threads(best)
        for _ in range(warmup):
            tp.model_forward(x, p)
        t0 = time.perf_counter()
        for _ in range(steps):
            tp.model_forward(x, p)
        dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3, best


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 40))
    v, ms, threads = cpu_reference_forward(steps, max(1, min(args.warmup, 3)))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "windows/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "B": B, "N": N, "W": W, "H": H,
                       "note": "reference CPU path = oracle/torch_port.py (reference's ATen op sequence; "
                               "the Python reference itself cannot travel to the GPU box)"},
            "cpu_baseline": {"value": v, "unit": "windows/s", "cores": threads, "kind": "port",
                             "sample": f"{steps} eval forwards of one {B}-window batch; os.cpu_count()={os.cpu_count()}"},
            "e2e": {"value": v, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from models.base_model import Model
    from stemgnn_b200 import synthetic as tp      # seeded weights / inputs (plain data generators)
    from stemgnn_b200 import _lib

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()

    model = Model(N, 2, W, MULTI, horizon=H)
    model.load_state_dict(tp.synthetic_params(N, W, H, MULTI, seed=0))
    model = model.to(dev).eval()
    x_host, _ = tp.synthetic_batch(B, N, W, H, seed=1234 + rank)
    x_host = x_host.pin_memory()
    x_dev = x_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            model(x_dev)
        # ---- device-timed: inputs resident in HBM, L2 flushed between steps -------------------
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
        sampler = ClockSampler(local_rank) if rank == 0 else None
        barrier()
        launches0 = lib.stemgnn_launch_count()
        for i in range(args.steps):
            flush.zero_()
            ev[i][0].record()
            model(x_dev)
            ev[i][1].record()
        barrier()
        launches = lib.stemgnn_launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)

        # ---- same forward with every GEMM forced to exact fp32 FFMA2 (Model.gemm_mode = 1) ---------------
        from stemgnn_b200 import runtime as _rt
        model.gemm_mode = _rt.GEMM_FP32
        for _ in range(3):
            model(x_dev)
        ev32 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                for _ in range(min(args.steps, 20))]
        for a0, a1 in ev32:
            flush.zero_()
            a0.record()
            model(x_dev)
            a1.record()
        torch.cuda.synchronize()
        fp32_ms = sum(a0.elapsed_time(a1) for a0, a1 in ev32) / len(ev32)
        model.gemm_mode = _rt.GEMM_AUTO
        model(x_dev)

        # ---- dominant kernel (GRU recurrence) with CUDA events on the launching stream --------------
        gru_ms = None
        try:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); e1.record(); torch.cuda.synchronize()           # materialise the handles
            tot = 0.0
            reps = min(args.steps, 10)
            for _ in range(reps):
                flush.zero_()
                lib.stemgnn_profile_gru(e0.cuda_event, e1.cuda_event)
                model(x_dev)
                lib.stemgnn_profile_gru(None, None)
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            gru_ms = tot / reps
        except Exception as exc:                                        # hook is optional evidence
            gru_ms = None
            sys.stderr.write(f"[bench] GRU event hook failed: {exc}\n")

        # ---- the tcgen05 GLU layer (61 % of the flops), timed alone through the C ABI with CUDA events ----
        glu_ms = None
        try:
            R, d = B * N, 4 * MULTI * W
            ga = torch.randn(R, d, device=dev)
            gw = [torch.randn(d, d, device=dev) / d ** 0.5 for _ in range(2)]
            gb = [torch.zeros(d, device=dev) for _ in range(2)]
            go = torch.empty(R, d, device=dev)
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

            def glu_call():
                rc = lib.stemgnn_glu_gemm(R, d, d, ga.data_ptr(), d, gw[0].data_ptr(), gb[0].data_ptr(),
                                          gw[1].data_ptr(), gb[1].data_ptr(), go.data_ptr(), d, 1, st)
                if rc:
                    raise RuntimeError(lib.stemgnn_last_error().decode())
            for _ in range(3):
                glu_call()
            tot, reps = 0.0, 10
            for _ in range(reps):
                flush.zero_()
                e0.record(); glu_call(); e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            glu_ms = tot / reps
        except Exception as exc:
            sys.stderr.write(f"[bench] GLU kernel timing failed: {exc}\n")

        # ---- end to end through the public API with host buffers ----------------------------------
        out_host = torch.empty(B, H, N).pin_memory()
        for _ in range(3):
            f, _a = model(x_host.to(dev, non_blocking=True))
            out_host.copy_(f, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            xd = x_host.to(dev, non_blocking=True)                      # H2D of the step's input
            f, _a = model(xd)
            out_host.copy_(f, non_blocking=True)                        # D2H of the step's result
            torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()

    # ---- second column: training step (handler.py:160-165) = zero_grad + forward (train mode, Philox
    #      dropout) + MSE + backward (+ the single flat-gradient all-reduce at N>1) + RMSprop ----------
    from stemgnn_b200 import ddp
    ddp.attach(model)
    model.train()
    y_dev = tp.synthetic_batch(B, N, W, H, seed=1234 + rank)[1].to(dev)
    optim = torch.optim.RMSprop(model.parameters(), lr=1e-4, eps=1e-8)
    crit = torch.nn.MSELoss()

    def train_step():
        model.zero_grad()
        f, _a = model(x_dev)
        loss = crit(f, y_dev)
        loss.backward()
        optim.step()
        return loss

    t_steps = max(5, min(args.steps, 20))
    for _ in range(3):
        train_step()
    tev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(t_steps)]
    barrier()
    for i in range(t_steps):
        flush.zero_()
        tev[i][0].record()
        train_step()
        tev[i][1].record()
    barrier()
    train_ms = sum(a.elapsed_time(b) for a, b in tev)
    model.eval()

    t = torch.tensor([dev_ms, e2e_s * 1e3, train_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, train_ms = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        return

    peak_tf, peak_gbs, peak_src = _peaks()
    gru_flops = 6.0 * B * N ** 3 + 12.0 * B * N * N               # recurrence + gates
    roof = None
    if gru_ms:
        ach = gru_flops / (gru_ms * 1e-3) / 1e12
        roof = {"kernel": "gru_cluster_kernel (GRU recurrence, fp32 FFMA2, 16-CTA clusters)",
                "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach / peak_tf,
                # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
                # (profiles/r01_ncu_gru_cluster_v3_raw.csv): 50.79 MB + 0.04 MB; algorithmic: the 49.2 MB input
                # projection it streams + 1.5 MB of W_hh per cluster wave
                "traffic": 50.83e6, "traffic_unit": "bytes/launch", "peak_source": peak_src,
                "ms_per_launch": gru_ms, "share_of_step": gru_ms / (dev_ms / args.steps),
                # context: this is an fp32 FFMA2 kernel (no tensor-core form keeps fp32 parity and fits the DSMEM
                # exchange budget); against the fp32 FMA peak of 148 SMs x 128 FMA/clk x 1.965 GHz:
                "fp32_fma_peak_tflops": 74.4, "frac_of_fp32_fma_peak": ach / 74.4,
                "note": "latency-bound recurrence: 358 dependent steps; flops = 6BN^3 + 12BN^2"}
    cpu_steps = 15
    cpu_v, cpu_ms, threads = cpu_reference_forward(cpu_steps, 2)
    value = world * B * args.steps / (dev_ms * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (GLU chain + folded output map: tf32 tensor-core operands, fp32 accumulate; see exact_fp32)",
            "data": "synthetic",
            "exact_fp32": {"value": B / (fp32_ms * 1e-3) * world, "unit": "windows/s", "ms_per_step": fp32_ms,
                           "what": "same forward with every GEMM on the fp32 FFMA2 path (Model.gemm_mode = 1), rank 0"},
            "config": {"workload": WORKLOAD, "B_per_gpu": B, "N": N, "W": W, "H": H, "multi_layer": MULTI,
                       "mode": "eval forward (Model.forward, no_grad)", "parallelism": f"dp{world} replicas",
                       "l2": "flushed between timed steps (256 MiB memset outside the event pairs)",
                       "flops_per_step": FORWARD_FLOPS},
            "e2e": {"value": world * B * args.steps / (e2e_ms * 1e-3), "unit": "windows/s",
                    "h2d_bytes_per_step": B * W * N * 4, "d2h_bytes_per_step": B * H * N * 4,
                    "ms_per_step": e2e_ms / args.steps},
            "train": {"value": world * B * t_steps / (train_ms * 1e-3), "unit": "windows/s",
                      "ms_per_step": train_ms / t_steps, "steps": t_steps,
                      "what": "zero_grad + forward(train, Philox dropout) + MSE + backward"
                              + (" + flat-gradient NCCL all-reduce" if world > 1 else "") + " + RMSprop step"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "roofline_glu": None if not glu_ms else {
                "kernel": "glu_tc_kernel (tcgen05 kind::tf32, one 240->240 GLU layer over B*N=11456 rows)",
                "bound": "tensor", "achieved": 4.0 * B * N * 240 * 240 / (glu_ms * 1e-3) / 1e12,
                "peak": peak_tf, "unit": "TFLOP/s",
                "frac": 4.0 * B * N * 240 * 240 / (glu_ms * 1e-3) / 1e12 / peak_tf,
                "ms_per_launch": glu_ms, "peak_source": peak_src + " (bf16 figure; TF32 runs at half rate)",
                "note": "L2 flushed before each launch: weights and activations come from HBM"},
            "cpu_baseline": {"value": cpu_v, "unit": "windows/s", "cores": threads, "kind": "port",
                             "sample": f"{cpu_steps} eval forwards of one {B}-window batch "
                                       f"({cpu_ms:.1f} ms each); os.cpu_count()={os.cpu_count()}"},
            "clocks": clocks}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
