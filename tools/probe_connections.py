"""Which streams share a hardware queue?  A 10 ms spin kernel on stream i, then a tiny kernel on stream j: if the tiny one
only finishes with the spin, i and j alias one connection (CUDA_DEVICE_MAX_CONNECTIONS)."""
import os, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
from cuda.bindings import runtime as rt
torch.zeros(1, device="cuda")
N = 72
def make(n):
    out = []
    for _ in range(n):
        err, s = rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking)
        assert int(err) == 0
        out.append(torch.cuda.ExternalStream(int(s)))
    return out
x = torch.zeros(1024, device="cuda")
spin = int(10e-3 * 1.9e9)
def touch(st):
    with torch.cuda.stream(st): x.add_(1)
def aliased(si, sj):
    torch.cuda.synchronize()
    with torch.cuda.stream(si): torch.cuda._sleep(spin)
    ev = torch.cuda.Event()
    t0 = time.perf_counter()
    with torch.cuda.stream(sj):
        x.add_(1); ev.record()
    ev.synchronize()
    return time.perf_counter() - t0 > 4e-3
A = make(N)
for st in A: touch(st)          # first use in creation order
for i in (0, 5):
    print("set A (used in creation order): stream %d shares a queue with" % i, [j for j in range(N) if j != i and aliased(A[i], A[j])], flush=True)
B = make(N)
for st in reversed(B): touch(st)  # first use in reverse order
print("set B (created after A, first used in reverse order): stream 0 shares a queue with", [j for j in range(N) if j != 0 and aliased(B[0], B[j])], flush=True)
print("A[0] vs set B:", [j for j in range(N) if aliased(A[0], B[j])], flush=True)
