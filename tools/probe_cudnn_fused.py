"""Does this torch/cuDNN build fuse conv+bias+relu (and +residual) for bf16 channels_last?"""
import time
import torch
import torch.nn.functional as F

dev = 'cuda'
torch.manual_seed(0)
for (cin, cout, k, s, H, W) in [(64, 64, 3, 2, 192, 320), (64, 64, 1, 1, 96, 160), (64, 256, 1, 1, 96, 160),
                                (256, 64, 1, 1, 96, 160), (64, 64, 3, 1, 96, 160), (256, 18, 3, 1, 96, 160),
                                (3, 64, 3, 2, 384, 640)]:
    for dt in (torch.bfloat16, torch.float32):
        x = torch.randn(8, cin, H, W, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(cout, cin, k, k, device=dev, dtype=dt) * 0.05).contiguous(memory_format=torch.channels_last)
        b = torch.randn(cout, device=dev, dtype=dt)
        ref = F.conv2d(x, w, b, s, k // 2).relu_()
        z = torch.randn_like(ref)
        ref2 = (F.conv2d(x, w, b, s, k // 2) + z).relu_()
        res = {}
        try:
            y = torch.cudnn_convolution_relu(x, w, b, (s, s), (k // 2, k // 2), (1, 1), 1)
            res['relu_err'] = float((y.float() - ref.float()).abs().max())
            res['cl'] = y.is_contiguous(memory_format=torch.channels_last)
            y2 = torch.cudnn_convolution_add_relu(x, w, z, 1.0, b, (s, s), (k // 2, k // 2), (1, 1), 1)
            res['add_relu_err'] = float((y2.float() - ref2.float()).abs().max())
            for name, fn in (('unfused', lambda: F.conv2d(x, w, b, s, k // 2).relu_()),
                             ('fused', lambda: torch.cudnn_convolution_relu(x, w, b, (s, s), (k // 2, k // 2), (1, 1), 1)),
                             ('nobias', lambda: F.conv2d(x, w, None, s, k // 2))):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res[name + '_us'] = round(e0.elapsed_time(e1) * 100, 1)
        except Exception as ex:
            res['error'] = repr(ex)[:200]
        print((cin, cout, k, s, H, W), str(dt).split('.')[-1], res, flush=True)
