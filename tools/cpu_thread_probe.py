"""How many host threads should the CPU oracle use on this box?  Times a UNet-like conv + linear at several thread counts."""
import time, torch, torch.nn.functional as F, os
x = torch.randn(2, 320, 32, 32); w = torch.randn(320, 320, 3, 3)
a = torch.randn(2048, 1280); b = torch.randn(1280, 1280)
for nt in (8, 16, 32, 64, 128):
    if nt > (os.cpu_count() or 8): break
    torch.set_num_threads(nt)
    for _ in range(2): F.conv2d(x, w, padding=1); a @ b
    t0 = time.perf_counter()
    for _ in range(10): F.conv2d(x, w, padding=1); a @ b
    dt = (time.perf_counter() - t0) / 10
    fl = 2 * 2 * 320 * 320 * 9 * 32 * 32 + 2 * 2048 * 1280 * 1280
    print(nt, "threads:", f"{dt*1e3:.2f} ms", f"{fl/dt/1e12:.3f} TFLOP/s")
