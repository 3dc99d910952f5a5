"""Bandwidth micro-benchmark of the instance-norm kernels (CUDA events; tensors >> L2).
python tools/bench_norm.py [N H W C]   -- register-staged (*8) kernels against the cp.async-pipelined range kernels at several ring depths /
grid sizes, plus a cross-check that both families produce the same tensors."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shmgan_b200 import ops
from shmgan_b200._lib import call

dims = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else [80, 256, 256, 64]
N, H, W, C = dims
x = torch.randn((N, H, W, C), device="cuda").bfloat16()
dy = torch.randn((N, H, W, C), device="cuda").bfloat16()
dyp = torch.randn((N, H // 2, W // 2, C), device="cuda").bfloat16()
addt = torch.randn((1, H, W, C), device="cuda").bfloat16()
gamma = torch.rand(C, device="cuda") + 0.5; beta = torch.randn(C, device="cuda")
db = torch.zeros(C, device="cuda")
out = torch.empty_like(x); cat = torch.empty((N, H, W, 2 * C), device="cuda", dtype=torch.bfloat16)
sums = ops.inorm_stats(x)
e = x.numel() * 2


def timeit(fn, iters=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


CASES = [
    ("inorm_stats", lambda: ops.inorm_stats(x), e),
    ("inorm_apply", lambda: ops.inorm_apply(x, sums, gamma, beta, out=out), 2 * e),
    ("inorm_apply +add", lambda: ops.inorm_apply(x, sums, gamma, beta, add=addt, out=out), 2 * e),
    ("inorm_apply +pool -> cat slice", lambda: ops.inorm_apply(x, sums, gamma, beta, out=cat[..., C:], pooled=True), 2.25 * e),
    ("inorm_apply pool only", lambda: ops.inorm_apply(x, sums, gamma, beta, pooled=True, want_out=False), 1.25 * e),
    ("inorm_bwd (stats+apply+dbias)", lambda: ops.inorm_bwd(x, sums, gamma, dy, None, dx=out, dbias=db), 5 * e),
    ("inorm_bwd dyA+dyP", lambda: ops.inorm_bwd(x, sums, gamma, dy, dyp, dx=out, dbias=db), 5.5 * e),
    ("inorm_bwd dyP", lambda: ops.inorm_bwd(x, sums, gamma, None, dyp, dx=out, dbias=db), 3.5 * e),
    ("act_bwd + dbias", lambda: ops.act_bwd(dy, x, 1, out=out, dbias=db), 3 * e),
]
if os.environ.get("BN_ONCE"):            # one launch of each case at the default configuration (for an ncu --set full capture)
    for name, fn, nbytes in CASES:
        fn()
    torch.cuda.synchronize()
    sys.exit(0)
CONFIGS = [("regs", (1, 0, 0))] + [("d%d g%d" % (d, g), (0, d, g)) for d in (1, 2, 4) for g in (0, 8, 16, 32)]
print("N=%d %dx%d C=%d bf16 (%.0f MB per tensor); GB/s of algorithmic bytes (%% of 6544)" % (N, H, W, C, e / 1e6))
print("%-32s" % "" + "".join("%9s" % c[0] for c in CONFIGS))
for name, fn, nbytes in CASES:
    row = "%-32s" % name
    for cname, cfg in CONFIGS:
        call("shm_norm_tune", *cfg)
        ms = timeit(fn)
        row += "%9.0f" % (nbytes / ms / 1e6)
    print(row, flush=True)

# ---- cross-check: both kernel families on the same inputs
def outputs():
    r = {}
    r["stats"] = ops.inorm_stats(x).clone()
    o, _ = ops.inorm_apply(x, sums, gamma, beta); r["apply"] = o.clone()
    o, _ = ops.inorm_apply(x, sums, gamma, beta, add=addt); r["apply_add"] = o.clone()
    c2 = torch.zeros_like(cat)
    _, p = ops.inorm_apply(x, sums, gamma, beta, add=addt, out=c2[..., C:], pooled=True); r["pool_out"] = c2.clone(); r["pool_p"] = p.clone()
    _, p = ops.inorm_apply(x, sums, gamma, beta, pooled=True, want_out=False); r["pool_only"] = p.clone()
    for tag, (a, b) in {"A": (dy, None), "AP": (dy, dyp), "P": (None, dyp)}.items():
        d = torch.zeros(C, device="cuda")
        r["bwd_" + tag] = ops.inorm_bwd(x, sums, gamma, a, b, dbias=d).clone(); r["dbias_" + tag] = d
    d = torch.zeros(C, device="cuda")
    r["act"] = ops.act_bwd(dy, x, 1, dbias=d).clone(); r["act_db"] = d
    return r
call("shm_norm_tune", 1, 0, 0); ref = outputs()
worst = 0.0
for cfg in [(0, 0, 0), (0, 2, 0), (0, 0, 16)]:
    call("shm_norm_tune", *cfg); got = outputs()
    for k in ref:
        a, b = ref[k].double(), got[k].double()
        err = float((a - b).abs().max() / (a.abs().max() + 1e-30))
        worst = max(worst, err)
        if err > 2e-3:
            print("MISMATCH", cfg, k, err)
call("shm_norm_tune", 0, 0, 0)
print("cross-check pipelined vs register-staged kernels: worst relative difference %.2e" % worst)
