import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
def bench(fn, rep=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(rep):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {}
for n in (2048, 4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    ms = bench(lambda: a @ b)
    out[f"dgemm_{n}_tflops"] = 2 * n**3 / ms * 1e-9
a = torch.randn(512, 4096, dtype=torch.float64, device="cuda"); b = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
ms = bench(lambda: a @ b); out["dgemm_512x4096x4096_tflops"] = 2*512*4096*4096/ms*1e-9
a = torch.randn(512, 256, dtype=torch.float64, device="cuda"); b = torch.randn(256, 65536, dtype=torch.float64, device="cuda")
ms = bench(lambda: a @ b); out["dgemm_512x65536x256_tflops"] = 2*512*256*65536/ms*1e-9
n = 8192
a = torch.randn(n, n, dtype=torch.float32, device="cuda"); b = torch.randn(n, n, dtype=torch.float32, device="cuda")
ms = bench(lambda: a @ b); out["sgemm_fp32_8192_tflops"] = 2*n**3/ms*1e-9
torch.backends.cuda.matmul.allow_tf32 = True
ms = bench(lambda: a @ b); out["tf32_8192_tflops"] = 2*n**3/ms*1e-9
# LAPACK-ish via cusolver for reference
a = torch.randn(512, 512, dtype=torch.float64, device="cuda")
ms = bench(lambda: torch.linalg.svd(a)); out["cusolver_svd_512_ms"] = ms
ms = bench(lambda: torch.linalg.eigh(a @ a.T)); out["cusolver_eigh_512_ms"] = ms
a = torch.randn(512, 256, dtype=torch.float64, device="cuda")
ms = bench(lambda: torch.linalg.qr(a)); out["cusolver_qr_512x256_ms"] = ms
print(json.dumps(out, indent=1))
open("gpurun_out/peaks_fp64.json", "w").write(json.dumps(out, indent=1))
