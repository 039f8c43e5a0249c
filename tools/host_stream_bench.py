"""Host-buffer scoring (ocm_simca_score_host_f32 / _f64in): pinned fp32 rows (direct copies) against pageable fp32 and
fp64 rows (pinned bounce buffers filled by host threads) and against uploading the whole matrix first.
usage: python tools/host_stream_bench.py [rows]"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from ocm_vae_simca_b200 import engine, native, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
p, A = 1024, 10
dev = torch.device("cuda")
Xd = synth.make_spectra_device(n, p, dev, seed=3)
pca = engine.fit_pca(Xd[:5000].contiguous(), None, A)
Sinv = torch.diag(torch.as_tensor(1.0 / pca.eigs_all[:A], device=dev))
X32 = Xd.cpu().numpy()
del Xd
out = {"rows": n, "p": p, "host_threads": os.environ.get("OCM_HOST_THREADS", "default"), "cores": len(os.sched_getaffinity(0))}
kw = dict(rule="alt", ct=0.05, cq=100.0, x_rms=pca.x_rms)


def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts)


t = timed(lambda: engine.score_host(X32, pca.mean, pca.P, Sinv, **kw))
out["pageable_f32"] = {"spectra_per_s": n / t, "link_GBs": 4.0 * p * n / t / 1e9}
Xp = torch.from_numpy(X32).pin_memory().numpy()
t = timed(lambda: engine.score_host(Xp, pca.mean, pca.P, Sinv, **kw))
out["pinned_f32"] = {"spectra_per_s": n / t, "link_GBs": 4.0 * p * n / t / 1e9}
del Xp
t = timed(lambda: engine.score(torch.from_numpy(X32).to(dev), pca.mean, pca.P, Sinv, **kw))
out["upload_whole_then_score_f32"] = {"spectra_per_s": n / t}
X64 = X32.astype(np.float64)
t = timed(lambda: engine.score_host(X64, pca.mean, pca.P, Sinv, **kw))
out["pageable_f64"] = {"spectra_per_s": n / t, "host_read_GBs": 8.0 * p * n / t / 1e9}
t = timed(lambda: engine.score(torch.from_numpy(X64).to(dev), pca.mean, pca.P, Sinv, **kw))
out["upload_whole_then_score_f64"] = {"spectra_per_s": n / t}
print(json.dumps(out))
# engine.to_device of a large numpy matrix (ocm_upload_rows) against torch's own pageable copy
t = timed(lambda: engine.to_device(X32))
print(json.dumps({"to_device_f32_GBs": X32.nbytes / t / 1e9, "torch_to_GBs": X32.nbytes / timed(lambda: torch.from_numpy(X32).to(dev)) / 1e9,
                  "to_device_f64_GBs": X64.nbytes / timed(lambda: engine.to_device(X64)) / 1e9}))
