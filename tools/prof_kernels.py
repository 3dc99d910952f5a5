"""One representative launch of every tensor-core kernel and of the instance-norm stream kernels at the training step's shapes (the 5B = 80-image
generator pass of configs[1] / the 10B = 160-image discriminator pass), for `ncu --set full` captures:

    python tools/prof_kernels.py && ncu --set full --clock-control none --import-source on \
        -k regex:'conv_|wgrad_|in_.*_p|act_bwd_p' -o gpurun_out/r02_kernels python tools/prof_kernels.py

Every kernel is launched ONCE (cold L2), after the weight re-layout kernels, so the report holds exactly one launch per label below.
Prints "label -> algorithmic FLOPs / bytes" so that tools/ncu_summary.py's durations can be turned into rates."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shmgan_b200 import ops

g = torch.Generator(device="cuda").manual_seed(1)
N = 80
ROWS = []


def conv(cin, cout, k=3, stride=1, transposed=False, act=1, bias=True):
    c = ops.Conv("t", k, k, cin, cout, stride=stride, transposed=transposed, act=act, bias=bias)
    wshape = (k, k, cout, cin) if transposed else (k, k, cin, cout)
    c.w = torch.randn(wshape, device="cuda", generator=g) * 0.03
    c.b = torch.zeros(cout, device="cuda") if bias else None
    c.dw = torch.zeros(wshape, device="cuda")
    c.db = torch.zeros(cout, device="cuda") if bias else None
    c.refresh_tc(1)
    return c


def x(n, s, c):
    return torch.randn((n, s, s, c), device="cuda", generator=g).bfloat16()


FAMILY = {"conv_big2_kernel<2, 2>": "conv_multi_kernel (big)", "conv_big2_kernel<2, 1>": "conv_multi_kernel (scatter)",
          "conv_scat_res_kernel<2>": "conv_multi_kernel (scatter)", "conv_multi_kernel<64, 2, 1>": "conv_multi_kernel (scatter)",
          "conv_tc_kernel<128>": "conv_tc_kernel", "wgrad_halo2_kernel": "wgrad_halo_kernel<1>", "wgrad_tc_kernel<128>": "wgrad_tc_kernel"}


def note(label, kernel, flops, nbytes):
    fam = FAMILY.get(kernel, "conv_halo_kernel" if kernel.startswith("conv_halo_kernel") else kernel)      # bench.py's per-kernel table keys
    ROWS.append({"label": label, "kernel": kernel, "family": fam, "algorithmic_flops": flops, "algorithmic_bytes": nbytes})


def fwd(label, kernel, c, xin, stats=False):
    n, h, w, _ = xin.shape
    c.fwd(xin, None, True, 1, want_stats=stats)
    note(label, kernel, c.flops(n, h, w), c.io_bytes(n, h, w, 2))


def dgrad(label, kernel, c, xshape):
    n, h, w, _ = xshape
    ho, wo = c.out_hw(h, w)
    dy = x(n, ho, c.cout)
    c.dgrad(dy, xshape, None, True, 1)
    note(label, kernel, c.flops(n, h, w), c.io_bytes(n, h, w, 2))


def wgrad(label, kernel, c, xin):
    n, h, w, _ = xin.shape
    ho, wo = c.out_hw(h, w)
    c.wgrad(xin, x(n, ho, c.cout), True, bias_done=True)
    note(label, kernel, c.flops(n, h, w), c.io_bytes(n, h, w, 2))


# layers first (their weight re-layout kernels are not captured: -k filters on the conv / wgrad / norm names)
L = {
    "dec2a": conv(512, 256), "enc2b": conv(128, 128), "enc1b": conv(64, 64), "dec4a": conv(128, 64), "enc2a": conv(64, 128),
    "up3T": conv(256, 128, 3, 2, True), "up4T": conv(128, 64, 3, 2, True), "d3": conv(128, 256, 3, 2, bias=False), "d5": conv(512, 1024, 3, 2, bias=False),
    "bott": conv(512, 512, 1),
}
X = {"512@64": x(N, 64, 512), "128@128": x(N, 128, 128), "64@256": x(N, 256, 64), "128@256": x(N, 256, 128), "256@64": x(N, 64, 256),
     "128@64x160": x(160, 64, 128), "512@16x160": x(160, 16, 512), "512@16": x(N, 16, 512)}
torch.cuda.synchronize()

fwd("big pair fwd 512->256 @64x64 N=80", "conv_big2_kernel<2, 2>", L["dec2a"], X["512@64"])
fwd("big pair fwd 128->128 @128x128 N=80", "conv_big2_kernel<2, 2>", L["enc2b"], X["128@128"])
fwd("halo fwd + IN stats 64->64 @256x256 N=80", "conv_halo_kernel<1, 64, 128, 1>", L["enc1b"], X["64@256"], stats=True)
fwd("halo fwd + IN stats 128->64 @256x256 N=80 (dec4a)", "conv_halo_kernel<2, 64, 128, 1>", L["dec4a"], X["128@256"], stats=True)
fwd("halo fwd 64->128 @256x256 N=80 (enc2a at level-1 size)", "conv_halo_kernel<1, 128, 128, 1>", L["enc2a"], X["64@256"])
fwd("scatter pair fwd ConvT 256->128 @64->128 N=80 (up3T)", "conv_big2_kernel<2, 1>", L["up3T"], X["256@64"])
fwd("scatter (resident weights) fwd ConvT 128->64 @128->256 N=80 (up4T)", "conv_scat_res_kernel<2>", L["up4T"], X["128@128"])
fwd("generic fwd 128->256 s2 @64->32 N=160 (d3)", "conv_tc_kernel<128>", L["d3"], X["128@64x160"])
fwd("generic fwd 1x1 512->512 @16x16 N=80 (bott)", "conv_tc_kernel<128>", L["bott"], X["512@16"])
dgrad("generic dgrad ConvT 256->128 (gather at stride 2, up3T)", "conv_tc_kernel<128>", L["up3T"], (N, 64, 64, 256))
dgrad("scatter pair dgrad 128->256 s2 (d3)", "conv_big2_kernel<2, 1>", L["d3"], (160, 64, 64, 128))
wgrad("wgrad halo MODE 0 64->64 @256x256 N=80", "wgrad_halo_kernel<0>", L["enc1b"], X["64@256"])
wgrad("wgrad halo MODE 0 128->64 @256x256 N=80 (dec4a)", "wgrad_halo_kernel<0>", L["dec4a"], X["128@256"])
wgrad("wgrad halo pair 512->256 @64x64 N=80", "wgrad_halo2_kernel", L["dec2a"], X["512@64"])
wgrad("wgrad halo MODE 1 128->128 @128x128 N=80 (single CTA: Cin % 256 != 0)", "wgrad_halo_kernel<1>", L["enc2b"], X["128@128"])
wgrad("wgrad s2 MODE 1 ConvT 256->128 (up3T)", "wgrad_s2_kernel<1>", L["up3T"], X["256@64"])
wgrad("wgrad s2 MODE 0 ConvT 128->64 (up4T)", "wgrad_s2_kernel<0>", L["up4T"], X["128@128"])
wgrad("wgrad generic 512->1024 s2 @16->8 N=160 (d5)", "wgrad_tc_kernel<128>", L["d5"], X["512@16x160"])

# instance-norm streams at the level-1 shape (80 x 256 x 256 x 64 bf16 = 671 MB per tensor)
z, dy = X["64@256"], x(N, 256, 64)
gamma, beta = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
e = z.numel() * 2
sums = ops.inorm_stats(z)
note("in_stats_p 80x256x256x64", "in_stats_p", 0, e)
ops.inorm_apply(z, sums, gamma, beta)
note("in_apply_p 80x256x256x64", "in_apply_p", 0, 2 * e)
ops.inorm_bwd(z, sums, gamma, dy, None, 1, dbias=torch.zeros(64, device="cuda"))
note("in_bwd_stats_p 80x256x256x64", "in_bwd_stats_p", 0, 2 * e)
note("in_bwd_apply_p 80x256x256x64", "in_bwd_apply_p", 0, 3 * e)
ops.act_bwd(dy, z, 1, dbias=torch.zeros(64, device="cuda"))
note("act_bwd_p 80x256x256x64", "act_bwd_p", 0, 3 * e)
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "prof_kernels_workloads.json"), "w") as fh:
    json.dump(ROWS, fh, indent=1)
for r in ROWS:
    print("%-58s %-34s %10.1f GFLOP %8.1f MB" % (r["label"], r["kernel"], r["algorithmic_flops"] / 1e9, r["algorithmic_bytes"] / 1e6))
