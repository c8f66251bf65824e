"""PCIe copy rates on the GPU box (pinned host memory): D2H alone, H2D alone, both at once.  The end-to-end
(host-buffer) figure of bench.py is bounded by these: amars_lw moves 676 B in and 1300 B out per solve."""
import json
import torch

dev = torch.device("cuda:0")
n_out, n_in = 681574400 // 8, 354549760 // 8  # one bench step of amars_lw
d_out = torch.empty(n_out, dtype=torch.float64, device=dev)
d_in = torch.empty(n_in, dtype=torch.float64, device=dev)
h_out = torch.empty(n_out, dtype=torch.float64).pin_memory()
h_in = torch.empty(n_in, dtype=torch.float64).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def d2h():
    with torch.cuda.stream(s1):
        h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)


def h2d():
    with torch.cuda.stream(s2):
        d_in.copy_(h_in, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


def both():
    with torch.cuda.stream(s1):
        h_out.copy_(d_out, non_blocking=True)
    with torch.cuda.stream(s2):
        d_in.copy_(h_in, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


t1, t2, t3 = timed(d2h), timed(h2d), timed(both)
print(json.dumps({"d2h_GBs": n_out * 8 / t1 / 1e9, "h2d_GBs": n_in * 8 / t2 / 1e9,
                  "both_ms": t3 * 1e3, "both_d2h_GBs": n_out * 8 / t3 / 1e9, "both_total_GBs": (n_out + n_in) * 8 / t3 / 1e9,
                  "amars_lw_e2e_ceiling_Msolves_s": 524288 / t3 / 1e6}))
