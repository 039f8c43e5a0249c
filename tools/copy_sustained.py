"""Does the HBM bandwidth itself hold in a long run on this box?  (denominator check for the roofline fraction)
    python tools/copy_sustained.py"""
import json, subprocess, time, torch
dev = torch.device("cuda", 0)
a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev).normal_()
b = torch.empty_like(a)
def timed(fn, reps, nbytes):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return [nbytes / ev[i].elapsed_time(ev[i + 1]) / 1e6 for i in range(reps)]   # GB/s
med = lambda x: sorted(x)[len(x) // 2]
out = {}
timed(lambda: b.copy_(a), 3, 1); time.sleep(2.0)
g = timed(lambda: b.copy_(a), 600, 2 * a.numel() * 2)            # read + write bytes, ~0.4 s back to back
out["copy"] = {"first5": [round(x) for x in g[:5]], "last5": [round(x) for x in g[-5:]], "best": round(max(g)), "median": round(med(g))}
r = a.view(torch.float32)
timed(lambda: r.sum(), 3, 1); time.sleep(2.0)
g = timed(lambda: r.sum(), 600, r.numel() * 4)
out["read_only_sum"] = {"first5": [round(x) for x in g[:5]], "last5": [round(x) for x in g[-5:]], "best": round(max(g)), "median": round(med(g))}
out["after"] = subprocess.run(["nvidia-smi", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print(json.dumps(out))
