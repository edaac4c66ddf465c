"""Raw pinned-memory PCIe bandwidth of the box (the ceiling of bench.py's e2e figure)."""
import json
import torch

def main():
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=4):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        s1.synchronize(); s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)
    def both():
        h2d(); d2h()
    out = {"h2d_GBps": timed(h2d), "d2h_GBps": timed(d2h)}
    b = timed(both)
    out["bidir_each_GBps"] = b
    print(json.dumps(out))

if __name__ == "__main__":
    main()
