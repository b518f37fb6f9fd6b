"""Drop-in for the reference's model_fuse.py: ``model_confuse(path, name='')`` (model_fuse.py:271-350) plus
the device-resident entry ``fuse_device`` used by predict.predict().  The work is done by bd_fuse /
bd_mask_cleanup (csrc/post.cu); this file is only the file-system contract and pointer plumbing."""
from __future__ import annotations

import glob

import numpy as np

from . import runtime as R


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise R.NativeError("building_detection_b200 needs a CUDA device (B200); there is no CPU path")
    return torch


def fuse_device(masks5):
    """masks5: (5,H,W) u8 {0,255} cuda tensor -> fused (H,W) u8 cuda tensor (stream-ordered, no host sync)."""
    t = _torch()
    assert masks5.is_cuda and masks5.dtype == t.uint8 and masks5.dim() == 3 and masks5.shape[0] == 5
    masks5 = masks5.contiguous()
    _, h, w = masks5.shape
    out = t.empty((h, w), dtype=t.uint8, device=masks5.device)
    stream = t.cuda.current_stream(masks5.device).cuda_stream
    R.check(R.lib().bd_fuse(R.context(masks5.device.index), masks5.data_ptr(), h, w, out.data_ptr(), stream))
    return out


def fuse_cleaned_device(cleaned5):
    """Vote >= 3 of five ALREADY cleaned masks, then the final clean-up (model_fuse.py:315-346): the second half of
    bd_fuse, for the multi-GPU path where the five clean-ups ran on different ranks."""
    t = _torch()
    cleaned5 = cleaned5.contiguous()
    _, h, w = cleaned5.shape
    out = t.empty((h, w), dtype=t.uint8, device=cleaned5.device)
    stream = t.cuda.current_stream(cleaned5.device).cuda_stream
    R.check(R.lib().bd_fuse_cleaned(R.context(cleaned5.device.index), cleaned5.data_ptr(), h, w, out.data_ptr(), stream))
    return out


def cleanup_device(mask):
    """One clean-up pass (fill_and_delete + eroede_dilate_process + only_plt) of an (H,W) u8 cuda tensor."""
    t = _torch()
    mask = mask.contiguous()
    h, w = mask.shape
    out = t.empty((h, w), dtype=t.uint8, device=mask.device)
    stream = t.cuda.current_stream(mask.device).cuda_stream
    R.check(R.lib().bd_mask_cleanup(R.context(mask.device.index), mask.data_ptr(), h, w, out.data_ptr(), stream))
    return out


def fuse(masks5):
    """Host arrays in, host array out: five (H,W) u8 masks -> fused (H,W) u8."""
    t = _torch()
    m = np.ascontiguousarray(np.stack([np.asarray(a, np.uint8) for a in masks5]))
    return fuse_device(t.from_numpy(m).cuda()).cpu().numpy()


def model_confuse(path, name=''):
    """model_fuse.py:271-350: needs exactly five ``*.png`` in ``path`` (otherwise prints and returns, :281-283);
    writes the fused mask to ``path + '\\' + name + '_result.png'`` (a literal backslash, :350)."""
    import cv2 as cv
    all_path = glob.glob(path + '/' + '*.png')
    print(all_path)
    if len(all_path) != 5:
        print('no five images')
        return
    masks = [cv.imread(p)[:, :, 0] for p in all_path]  # fill_and_delete reads channel 0 (:10)
    cv.imwrite(path + r'\{}_result.png'.format(name), fuse(masks))
