import os, time, torch, numpy as np
print('cpus', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
n = 132 * 1024 * 1024
a = torch.empty(n, dtype=torch.uint8).pin_memory(); b = torch.empty(n, dtype=torch.uint8).pin_memory(); a.fill_(1); b.fill_(2)
for _ in range(3):
    t0 = time.perf_counter(); b.copy_(a); t1 = time.perf_counter(); print('pinned->pinned copy %.1f ms %.1f GB/s' % ((t1-t0)*1e3, n/(t1-t0)/1e9))
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(a, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter(); print('H2D pinned %.1f ms %.1f GB/s' % ((t1-t0)*1e3, n/(t1-t0)/1e9))
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); a.copy_(d, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter(); print('D2H pinned %.1f ms %.1f GB/s' % ((t1-t0)*1e3, n/(t1-t0)/1e9))
x = np.ones(n, dtype=np.uint8); y = np.empty_like(x)
for _ in range(3):
    t0 = time.perf_counter(); np.copyto(y, x); t1 = time.perf_counter(); print('numpy copy %.1f ms %.1f GB/s' % ((t1-t0)*1e3, n/(t1-t0)/1e9))
