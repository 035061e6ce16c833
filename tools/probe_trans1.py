"""cuDNN time of the 256 -> 18 / 36 transition convs, with the output channels as they are vs
zero-padded to a multiple of 16/32 (+ a slicing copy).  CUDA-graph timing, bf16 channels-last."""
import torch

B = 8
x = torch.randn(B, 256, 96, 160, device='cuda', dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)


def timed(fn, n=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / n * 1e3


for cout, stride in ((18, 1), (36, 2)):
    for pad_to in (cout, 32 if cout == 18 else 48, 64):
        w = torch.zeros(pad_to, 256, 3, 3, device='cuda', dtype=torch.bfloat16)
        w[:cout] = torch.randn(cout, 256, 3, 3, device='cuda') * 0.02
        w = w.contiguous(memory_format=torch.channels_last)
        b = torch.zeros(pad_to, device='cuda', dtype=torch.bfloat16)

        def run():
            y = torch.cudnn_convolution_relu(x, w, b, (stride, stride), (1, 1), (1, 1), 1)
            if pad_to != cout:
                y = y[:, :cout].contiguous(memory_format=torch.channels_last)
            return y
        print(f'256 -> {cout} (stride {stride}) computed as Cout = {pad_to}: {timed(run):7.1f} us', flush=True)
# bare conv (transition1[0][0] is conv only, no BN/ReLU on branch 0 in the reference quirk)
w = (torch.randn(18, 256, 3, 3, device='cuda') * 0.02).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
print(f'bare F.conv2d 256 -> 18: {timed(lambda: torch.nn.functional.conv2d(x, w, None, 1, 1)):7.1f} us')
w32 = torch.zeros(32, 256, 3, 3, device='cuda', dtype=torch.bfloat16)
w32[:18] = w
w32 = w32.contiguous(memory_format=torch.channels_last)
print(f'bare F.conv2d 256 -> 32 + slice: '
      f'{timed(lambda: torch.nn.functional.conv2d(x, w32, None, 1, 1)[:, :18].contiguous(memory_format=torch.channels_last)):7.1f} us')
