"""A competing GPU process (time-slices the device with whatever runs next to it): used to shake out timing assumptions."""
import sys, time, torch
a = torch.randn(8192, 8192, device="cuda"); t0 = time.time()
while time.time() - t0 < float(sys.argv[1]):
    for _ in range(20):
        b = a @ a
    torch.cuda.synchronize()
