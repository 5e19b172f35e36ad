import os, sys, torch
sys.path.insert(0, "/root/repo")
from haptools_b200 import engine
dev = torch.device('cuda')
n, H = 200_000, 100_000
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
tiles = (n + 1023) // 1024
out = torch.empty((n, 2 * H), dtype=torch.uint8, device=dev)
bits = torch.empty((tiles, H, 32, 2), dtype=torch.int32, device=dev); bits.random_(0, 2**31-1)
ms = timeit(lambda: engine.unpack_bits_u8(bits, n, H, out.data_ptr(), 2 * H))
print("HT_EXP_UNPACK=%s: %.2f ms" % (os.environ.get("HT_EXP_UNPACK", "0"), ms))
