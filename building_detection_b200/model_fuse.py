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


def plane_words(w):
    """words per row of a bit-packed mask plane (csrc/rle.cuh)"""
    return int(R.lib().bd_plane_words_per_row(int(w)))


def pack_device(mask):
    """(H,W) u8 cuda tensor -> (H, plane_words(W)) int32 cuda tensor, 1 bit per pixel (nonzero = set).  Rows are
    independent, so a band of rows can be packed, shipped and OR-ed on its own (multi-GPU gather)."""
    t = _torch()
    mask = mask.contiguous()
    h, w = mask.shape
    out = t.empty((h, plane_words(w)), dtype=t.int32, device=mask.device)
    stream = t.cuda.current_stream(mask.device).cuda_stream
    R.check(R.lib().bd_mask_pack(R.context(mask.device.index), mask.data_ptr(), h, w, out.data_ptr(), stream))
    return out


def unpack_device(plane, w):
    t = _torch()
    plane = plane.contiguous()
    h = plane.shape[0]
    out = t.empty((h, w), dtype=t.uint8, device=plane.device)
    stream = t.cuda.current_stream(plane.device).cuda_stream
    R.check(R.lib().bd_mask_unpack(R.context(plane.device.index), plane.data_ptr(), h, w, out.data_ptr(), stream))
    return out


def fuse_planes_device(planes5, w, cleaned=False):
    """planes5: (5, H, plane_words(W)) int32 cuda tensor of packed masks -> fused (H,W) u8 cuda tensor."""
    t = _torch()
    planes5 = planes5.contiguous()
    assert planes5.dim() == 3 and planes5.shape[0] == 5 and planes5.shape[2] == plane_words(w)
    h = planes5.shape[1]
    out = t.empty((h, w), dtype=t.uint8, device=planes5.device)
    stream = t.cuda.current_stream(planes5.device).cuda_stream
    R.check(R.lib().bd_fuse_planes(R.context(planes5.device.index), planes5.data_ptr(), 1 if cleaned else 0, h, w,
                                   out.data_ptr(), None, stream))
    return out


def debug_stage(mask, stage):
    """test hook: an intermediate plane of one clean-up pass (bd_debug_cleanup_stage) of a host (H,W) u8 mask"""
    t = _torch()
    m = t.from_numpy(np.ascontiguousarray(mask, np.uint8)).cuda()
    out = t.empty_like(m)
    R.check(R.lib().bd_debug_cleanup_stage(R.context(m.device.index), m.data_ptr(), m.shape[0], m.shape[1], stage,
                                           out.data_ptr(), t.cuda.current_stream().cuda_stream))
    return out.cpu().numpy()


def debug_labels(mask, fg=1, conn8=1):
    """test hook: run-based component labels as per-pixel labels (bd_debug_labels)"""
    t = _torch()
    m = t.from_numpy(np.ascontiguousarray(mask, np.uint8)).cuda()
    out = t.empty(m.shape, dtype=t.int32, device=m.device)
    R.check(R.lib().bd_debug_labels(R.context(m.device.index), m.data_ptr(), m.shape[0], m.shape[1], fg, conn8,
                                    out.data_ptr(), t.cuda.current_stream().cuda_stream))
    return out.cpu().numpy()


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
    from . import png0
    masks = []
    for p in all_path:  # fill_and_delete reads channel 0 of cv.imread's BGR image (:10); a grey PNG has one
        with open(p, 'rb') as f:
            m = png0.decode_gray(f.read())
        masks.append(m if m.ndim == 2 else cv.imread(p)[:, :, 0])
    with open(path + r'\{}_result.png'.format(name), 'wb') as f:  # level-0 PNG like cv.imwrite(..., 0) (:350)
        f.write(png0.encode_gray(fuse(masks)))
