"""PCIe ceiling for the end-to-end step: pinned H2D and D2H copies of the step's byte counts,
alone and concurrently (what step_e2e overlaps).  python tools/pcie_probe.py"""
import torch

MB = 206
n = MB * 1024 * 1024
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, a, b in (("h2d only", True, False), ("d2h only", False, True), ("both", True, True)):
    run(a, b, 2)
    ms = run(a, b)
    print("%-9s %.3f ms per %d MB  -> %.1f GB/s per direction" % (name, ms, MB, n / ms / 1e6))
