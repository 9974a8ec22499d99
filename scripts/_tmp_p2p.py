import torch, time
a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
b = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:1")
print("can_access_peer", torch.cuda.can_device_access_peer(0, 1))
for n in (1 << 20, 16 << 20, 256 << 20):
    for _ in range(3):
        b[:n].copy_(a[:n])
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(10):
        b[:n].copy_(a[:n])
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    dt = (time.perf_counter() - t0) / 10
    print("copy %d MB: %.1f us, %.1f GB/s" % (n >> 20, dt * 1e6, n / dt / 1e9))
