import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pharmacoforge_b200 import train_ops as T
x = torch.randn(256, 128, device="cuda", requires_grad=True)
w = torch.randn(128, 128, device="cuda", requires_grad=True)
b = torch.randn(128, device="cuda", requires_grad=True)
def bench(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
with torch.no_grad():
    print("silu fwd no_grad      us", bench(lambda: T.silu(x)))
    print("linear fwd no_grad    us", bench(lambda: T.linear(x, w, b)))
    print("torch silu no_grad    us", bench(lambda: torch.nn.functional.silu(x)))
print("silu fwd+bwd          us", bench(lambda: T.silu(x).sum().backward()))
print("linear fwd+bwd        us", bench(lambda: T.linear(x, w, b).sum().backward()))
print("torch linear fwd+bwd  us", bench(lambda: torch.nn.functional.linear(x, w, b).sum().backward()))
