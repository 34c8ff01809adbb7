"""Bring-up / measurement harness for the tensor-core GRU recurrence (gru_tc.cu), run on the GPU box.
Each case runs in a SUBPROCESS so a trapped kernel (lost mbarrier signal) cannot take the other cases down.

    python tools/gru_tc_check.py            # parity of path 3 vs the host GRU for a list of shapes + timings
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [(4, 40, 12), (8, 70, 12), (3, 53, 12), (7, 140, 8), (32, 140, 12), (32, 358, 12), (5, 358, 12),
         (64, 228, 12), (33, 325, 28), (2, 448, 12)]


def _host_gru(x_seq, p):
    """ATen GRU on the host (1 layer, h0 = 0) with the model's weights: the comparison point of this tool.  (Plain torch — the
    tools do not import oracle/.)"""
    import torch
    hidden = p["GRU.weight_hh_l0"].shape[1]
    h0 = x_seq.new_zeros(1, x_seq.shape[1], hidden)
    flat = [p["GRU.weight_ih_l0"], p["GRU.weight_hh_l0"], p["GRU.bias_ih_l0"], p["GRU.bias_hh_l0"]]
    out, _ = torch._VF.gru(x_seq, h0, flat, True, 1, 0.0, False, False, False)
    return out


def one(B, N, W, path, reps):
    import torch
    from stemgnn_b200 import _lib as L, runtime, synthetic as sy
    lib = L.load()
    dev = torch.device("cuda:0")
    p = sy.synthetic_params(N, W, 3, 5, seed=N, scale_mode="trained")
    x, _ = sy.synthetic_batch(B, N, W, 3, seed=7)
    pd = {k: v.to(dev) for k, v in p.items()}
    dims = L.Dims(B, N, W, 3, 5)
    ptrs = runtime.build_ptrs({k: pd.get(k) for k in runtime.PARAM_KEYS})
    ws = runtime.alloc_workspace(dims, False, dev)
    xd = x.to(dev)
    key = torch.empty(B, N, device=dev); query = torch.empty(B, N, device=dev); out = torch.empty(N, B, N, device=dev)
    st = runtime._stream_ptr(dev)

    def call(o):
        rc = lib.stemgnn_gru_keyquery_forward(ctypes.byref(dims), ctypes.byref(ptrs), xd.data_ptr(), key.data_ptr(),
                                              query.data_ptr(), o, path, ws.data_ptr(), ws.numel(), st)
        L.check(rc, "gru")
    call(out.data_ptr())
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = _host_gru(x.permute(2, 0, 1).contiguous(), p)
    rk = torch.einsum("sbh,s->bh", ref.double(), p["weight_key"][:, 0].double()).float()
    e_out = float((out.cpu() - ref).abs().max()); e_key = float((key.cpu() - rk).abs().max())
    msg = f"B={B} N={N} W={W} path={path}: max|gru_out err|={e_out:.3e} max|key err|={e_key:.3e}"
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            call(None)
        e0.record()
        for _ in range(reps):
            call(None)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        msg += f"  | {us:.1f} us/call (incl. prep+pack) = {us * 1965 / N:.0f} cycles/step"
    print(msg, flush=True)


def stress(B, N, W, calls):
    """Race detector: `calls` independent forwards with gru_out, every result compared with the host GRU."""
    import torch
    from stemgnn_b200 import _lib as L, runtime, synthetic as sy
    lib = L.load()
    dev = torch.device("cuda:0")
    p = sy.synthetic_params(N, W, 3, 5, seed=N, scale_mode="trained")
    x, _ = sy.synthetic_batch(B, N, W, 3, seed=7)
    pd = {k: v.to(dev) for k, v in p.items()}
    dims = L.Dims(B, N, W, 3, 5)
    ptrs = runtime.build_ptrs({k: pd.get(k) for k in runtime.PARAM_KEYS})
    ws = runtime.alloc_workspace(dims, False, dev)
    xd = x.to(dev)
    key = torch.empty(B, N, device=dev); query = torch.empty(B, N, device=dev); out = torch.empty(N, B, N, device=dev)
    with torch.no_grad():
        ref_cpu = _host_gru(x.permute(2, 0, 1).contiguous(), p)
        ref64 = _host_gru(x.permute(2, 0, 1).contiguous().double(), {k: v.double() for k, v in p.items()})
        ref = ref_cpu.to(dev)
    d = (ref_cpu.double() - ref64).abs()
    rows = sorted(set(int(v) for v in (d.amax(dim=(0, 2)) > 4e-6).nonzero().flatten()))
    print(f"  host reference check: max|fp32 ATen GRU - fp64 ATen GRU| = {float(d.max()):.3e} (threads={torch.get_num_threads()}); "
          f"sequences above 4e-6: {rows}", flush=True)
    ref = ref64.float().to(dev)          # compare the GPU result with the fp64 host GRU
    errs = []
    junk = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    for i in range(calls):
        if i % 3 == 1:
            junk.zero_()                      # perturb timing / cache state
        rc = lib.stemgnn_gru_keyquery_forward(ctypes.byref(dims), ctypes.byref(ptrs), xd.data_ptr(), key.data_ptr(),
                                              query.data_ptr(), out.data_ptr() if i % 2 == 0 else None, 3, ws.data_ptr(),
                                              ws.numel(), runtime._stream_ptr(dev))
        L.check(rc, "gru")
        if i % 2 == 0:
            e = (out - ref).abs()
            errs.append(float(e.max()))
            if errs[-1] > 4e-6 and not any(x > 4e-6 for x in errs[:-1]):
                per_step = e.amax(dim=(1, 2))
                s0 = int((per_step > 4e-6).nonzero()[0])
                bad = (e[s0] > 4e-6).nonzero()
                bs = sorted(set(int(v) for v in bad[:, 0])); us = sorted(set(int(v) for v in bad[:, 1]))
                print(f"  first bad call {i}: first bad step {s0} (err {float(per_step[s0]):.2e}); sequences {bs[:12]} "
                      f"units {us[:6]}..{us[-3:]} ({len(us)} units); err at step {s0 - 1}: {float(per_step[s0 - 1]):.2e}", flush=True)
    bad = [e for e in errs if e > 4e-6]
    print(f"stress B={B} N={N}: {len(errs)} checked calls, max err {max(errs):.3e}, {len(bad)} above 4e-6", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stress":
        stress(*[int(v) for v in sys.argv[2:6]])
        sys.exit(0)
    if len(sys.argv) > 1:
        B, N, W, path, reps = [int(v) for v in sys.argv[1:6]]
        one(B, N, W, path, reps)
        sys.exit(0)
    for (B, N, W) in CASES:
        for path in (3,):
            r = subprocess.run([sys.executable, __file__, str(B), str(N), str(W), str(path), "5"], capture_output=True,
                               text=True, timeout=300)
            print(r.stdout.strip() or f"B={B} N={N} W={W} path={path}: FAILED rc={r.returncode}\n{r.stderr[-1500:]}", flush=True)
    for g in ("2", "3", "4", "5", "8"):
        env = dict(os.environ, STEMGNN_GRU_TC_G=g)
        r = subprocess.run([sys.executable, __file__, "32", "358", "12", "3", "20"], capture_output=True, text=True,
                           timeout=300, env=env)
        print(f"G={g}: " + (r.stdout.strip() or f"FAILED rc={r.returncode} {r.stderr[-800:]}"), flush=True)
    r = subprocess.run([sys.executable, __file__, "32", "358", "12", "2", "20"], capture_output=True, text=True, timeout=300)
    print("FFMA2 cluster kernel: " + (r.stdout.strip() or r.stderr[-800:]), flush=True)
