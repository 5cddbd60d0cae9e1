import torch, time
n = 1146617856 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device="cuda"); d_out = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunks=1, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        c = n // chunks
        for i in range(chunks):
            if h2d:
                with torch.cuda.stream(s1): d_in[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return n * 4 / dt / 1e9
for _ in range(2): run(True, True)
print("H2D alone %.1f GB/s" % run(True, False)); print("D2H alone %.1f GB/s" % run(False, True))
print("both %.1f GB/s each" % run(True, True)); print("both, 6 chunks %.1f GB/s each" % run(True, True, 6))
