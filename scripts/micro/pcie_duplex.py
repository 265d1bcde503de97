"""PCIe: H2D alone, D2H alone, both at once on two streams (pinned memory)."""
import torch, time
n = 64 * 1024 * 1024
h1 = torch.empty(n, dtype=torch.float32).pin_memory(); h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d1 = torch.empty(n, dtype=torch.float32, device='cuda'); d2 = torch.empty(n, dtype=torch.float32, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
    return n * 4 / dt / 1e9
for _ in range(2):
    print("H2D alone %.1f GB/s | D2H alone %.1f GB/s | both: %.1f GB/s each" % (run(True, False), run(False, True), run(True, True)))
