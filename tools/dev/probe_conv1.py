"""dev probe: the front end fused into conv1 (tcgen05) vs the unfused route
(our bf16 front end -> HBM -> cuDNN bf16 conv through torch, channels_last)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import shdr
from shdr import _native as N
shdr.require_gpu()
FLOP_PX = 2 * 49 * 93 * 64
REPS = int(os.environ.get('REPS', '20'))
def timeit(fn, reps=None):
    reps = reps or REPS
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
sh = torch.cuda.current_stream().cuda_stream
shapes = [(8, 512, 512), (32, 512, 512), (1, 2160, 3840)]
if len(sys.argv) > 1: shapes = shapes[:int(sys.argv[1])]
for (n, h, w) in shapes:
    img = torch.rand((n, h, w, 3), device="cuda")
    kern = torch.randn((7, 7, 93, 64), device="cuda") / 67.5
    bias = torch.randn(64, device="cuda") * 0.1
    packed = torch.empty(N.lib.shdr_conv1_packed_bytes() // 4, device="cuda")
    N.check(N.lib.shdr_conv1_pack_weights_f32(kern.data_ptr(), packed.data_ptr(), sh))
    oh, ow = (h + 1) // 2, (w + 1) // 2
    out = torch.empty((n, oh, ow, 64), device="cuda")
    fused = lambda: N.check(N.lib.shdr_frontend_conv1_f32(img.data_ptr(), packed.data_ptr(), None, bias.data_ptr(), 0, out.data_ptr(), n, h, w, sh))
    tf_ = timeit(fused)
    opx = n * oh * ow
    line = f"{n}x{h}x{w}: fused {tf_:.4f} ms = {opx * FLOP_PX / tf_ / 1e9:.0f} TFLOP/s"
    if "--nolib" not in sys.argv:
        feat = torch.empty((n, h, w, 93), device="cuda", dtype=torch.bfloat16)
        wt = kern.permute(3, 2, 0, 1).contiguous(memory_format=torch.channels_last).bfloat16()
        bb = bias.bfloat16()
        def unfused():
            N.check(N.lib.shdr_frontend_bf16(img.data_ptr(), feat.data_ptr(), n, h, w, sh))
            x = torch.nn.functional.pad(feat.permute(0, 3, 1, 2), (2, 3, 2, 3))
            return torch.nn.functional.conv2d(x, wt, bb, stride=2)
        def conv_only():
            x = torch.nn.functional.pad(feat.permute(0, 3, 1, 2), (2, 3, 2, 3))
            return torch.nn.functional.conv2d(x, wt, bb, stride=2)
        tu = timeit(unfused); tc = timeit(conv_only)
        ref = unfused().permute(0, 2, 3, 1).float()
        fused(); torch.cuda.synchronize()
        err = (out - ref).abs().max().item() / ref.abs().max().item()
        line += f" | bf16 front end + cuDNN bf16 conv {tu:.4f} ms (conv alone {tc:.4f}) -> fused is {tu / tf_:.2f}x; rel diff {err:.1e}"
    print(line, flush=True)
