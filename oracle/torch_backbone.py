"""oracle/torch_backbone.py -- plain PyTorch fp32 executor for hyperpose_b200.models.Graph.  TEST INFRASTRUCTURE ONLY
(tests, and the CPU-baseline / --impl reference legs of bench.py, where it stands in for the conv stage on the host cores:
the reference runs its convs in TensorRT on a GPU, src/tensorrt.cpp:387-396, and has no CPU implementation of them).

The backbone is a floating-point kernel, so its checker is a torch fp32 reference of the same ops
(F.conv2d / max_pool2d / PReLU), as the reference's TensorRT FP32 engine would compute them.
`emulate_fp16=True` additionally rounds weights and stored activations to fp16 exactly where the
engine does (fp16 operands, fp32 accumulation), which isolates kernel bugs from precision."""
import numpy as np
import torch
import torch.nn.functional as F

from hyperpose_b200 import models


def run_graph(g: models.Graph, frames_u8: np.ndarray, factor=1.0 / 255, flip_rgb=True, emulate_fp16=False, device="cuda",
              upto=None):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    N, H, W, _ = frames_u8.shape
    q = (lambda t: t.half().float()) if emulate_fp16 else (lambda t: t)
    bufs = []
    for (c, d) in g.buffers:
        h, w = H, W
        for _ in range(d):
            h, w = (h + 1) // 2, (w + 1) // 2
        bufs.append(torch.zeros(N, c, h, w, device=device))
    conf = paf = None
    img, img_stride = None, 1

    def same_pad(x, k, stride):
        """TF 'SAME': out = ceil(in/stride); pad_before = total // 2"""
        pads = []
        for dim in (x.shape[3], x.shape[2]):   # F.pad order: W first, then H
            out = (dim + stride - 1) // stride
            total = max((out - 1) * stride + k - dim, 0)
            pads += [total // 2, total - total // 2]
        return F.pad(x, pads)

    for oi, op in enumerate(g.ops):
        if upto is not None and oi > upto:
            break
        if op.type == models.OP_IM2COL3:
            x = (frames_u8.astype(np.float64) * factor).astype(np.float32)          # data.cpp:48
            if flip_rgb:
                x = x[..., ::-1]
            x = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2))).to(device)
            x = x - torch.tensor(g.mean, dtype=torch.float32, device=device).view(1, 3, 1, 1)
            img, img_stride = q(x), op.stride
            if op.stride == 1:
                bufs[op.out_buf][:, :3] = img
        elif op.type == models.OP_MAXPOOL2:
            x = bufs[op.in_buf][:, op.in_ch_off:op.in_ch_off + op.cout_g]
            K = op.R if op.R else 2
            # TF SAME max-pool: pad with -inf (window clipped at the border)
            pads = []
            for dim in (x.shape[3], x.shape[2]):
                out = (dim + 1) // 2
                total = max((out - 1) * 2 + K - dim, 0)
                pads += [total // 2, total - total // 2]
            xp = F.pad(x, pads, value=float("-inf"))
            bufs[op.out_buf][:, op.out_ch_off:op.out_ch_off + op.cout_g] = F.max_pool2d(xp, K, 2)
        elif op.type == models.OP_CONV:
            G, co, ci, R, S = op.weight.shape
            w = q(torch.from_numpy(op.weight.reshape(G * co, ci, R, S)).to(device))
            if op.im2col_input:
                y = F.conv2d(same_pad(img, R, img_stride), w, torch.from_numpy(op.bias).to(device), stride=img_stride)
            else:
                x = bufs[op.in_buf][:, op.in_ch_off:op.in_ch_off + G * ci]
                y = F.conv2d(x, w, torch.from_numpy(op.bias).to(device), padding=(R // 2, S // 2), groups=G)
            a = torch.from_numpy(op.alpha).to(device).view(1, -1, 1, 1)
            res = bufs[op.res_buf][:, op.res_ch_off:op.res_ch_off + G * co] if op.res_mode else None
            if op.res_mode == 1:
                y = y + res
            y = torch.where(y > 0, y, y * a)
            if op.res_mode == 2:
                y = y + res
            if op.out_mode == models.OUT_F32_NCHW_SPLIT:
                conf, paf = y[:, :op.split].contiguous(), y[:, op.split:].contiguous()
            else:
                bufs[op.out_buf][:, op.out_ch_off:op.out_ch_off + G * co] = q(y)
        elif op.type == models.OP_PIFPAF_HEAD:
            def head(raw, fields, comps, is_paf):
                x = raw[:, :fields * comps * 4]
                b, c, h, w = x.shape
                x = x.reshape(b, c // 4, 2, 2, h, w).permute(0, 1, 4, 2, 5, 3).reshape(b, c // 4, 2 * h, 2 * w)   # pifpaf/utils.py:371-379
                x = x[:, :, :2 * h - 1, :2 * w - 1].reshape(b, fields, comps, 2 * h - 1, 2 * w - 1).clone()
                gy, gx = torch.meshgrid(torch.arange(2 * h - 1, device=device), torch.arange(2 * w - 1, device=device), indexing="ij")
                x[:, :, 0] = torch.sigmoid(x[:, :, 0])
                for cx in ((1, 3) if is_paf else (1,)):
                    x[:, :, cx] += gx
                for cy in ((2, 4) if is_paf else (2,)):
                    x[:, :, cy] += gy
                for cs in ((7, 8) if is_paf else (4,)):
                    x[:, :, cs] = F.softplus(x[:, :, cs])
                return x
            conf = head(bufs[op.in_buf], 17, 5, False)
            paf = head(bufs[op.res_buf], 19, 9, True)
        elif op.type == models.OP_DWCONV:
            C, K, _ = op.weight.shape
            x = bufs[op.in_buf][:, op.in_ch_off:op.in_ch_off + C]
            w = torch.from_numpy(op.weight.reshape(C, 1, K, K)).to(device)      # depthwise weights stay fp32 in the engine
            y = F.conv2d(same_pad(x, K, op.stride), w, torch.from_numpy(op.bias).to(device), stride=op.stride, groups=C)
            a = torch.from_numpy(op.alpha).to(device).view(1, -1, 1, 1)
            y = torch.where(y > 0, y, y * a)
            bufs[op.out_buf][:, op.out_ch_off:op.out_ch_off + C] = q(y)
    return conf, paf, bufs
