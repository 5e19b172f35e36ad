"""PCIe ceilings on the box for the end-to-end leg: D2H alone, H2D alone, both at once (pinned host memory)."""
import torch, time
dev = torch.device("cuda")
GB = 1 << 30
d_out = torch.empty(3 * GB, dtype=torch.uint8, device=dev); h_out = torch.empty(3 * GB, dtype=torch.uint8).pin_memory()
d_in = torch.empty(3 * GB // 2, dtype=torch.uint8, device=dev); h_in = torch.empty(3 * GB // 2, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(d2h, h2d, chunk=None):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if d2h:
        with torch.cuda.stream(s1):
            if chunk:
                for o in range(0, h_out.numel(), chunk): h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
            else: h_out.copy_(d_out, non_blocking=True)
    if h2d:
        with torch.cuda.stream(s2): d_in.copy_(h_in, non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t0
for name, a in (("D2H alone", (True, False)), ("H2D alone", (False, True)), ("both", (True, True)), ("D2H in 256 MB chunks", (True, False, 256 << 20)), ("D2H in 64 MB chunks + H2D", (True, True, 64 << 20))):
    run(*a); t = min(run(*a) for _ in range(3))
    print(f"{name}: {t*1e3:.1f} ms  D2H {h_out.numel()/t/1e9 if a[0] else 0:.1f} GB/s  H2D {d_in.numel()/t/1e9 if a[1] else 0:.1f} GB/s")
