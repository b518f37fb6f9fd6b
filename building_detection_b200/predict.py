"""Drop-in for the reference's predict.py call surface (load_model / run_model / detection /
write_points, predict.py:17-132) plus the in-memory ``predict(image)`` entry the north star names.

The file-based functions keep the reference's contract -- PNG files in ``user_path`` are the hand-off
between stages -- so that buildAPI-style callers work unchanged; ``predict(image)`` runs the same
stages without touching the disk, with every intermediate resident in HBM.
"""
from __future__ import annotations

import os

import numpy as np

from . import scene as S
from .predict_model.bam import Xception_DeepLabV3_Plus_bam
from .predict_model.hrnet import HRNet
from .predict_model.res34 import ResNetFamily
from .predict_model.scse import UNet
from .predict_model.v3plus import Xception_DeepLabV3_Plus

# module-level models, like predict.py:138 (filled by load_model())
res_model = hr_model = v3_model = unet_model = bam_model = None
MODEL_PREFIXES = ('res34_', 'hrnet_', 'v3plus_', 'scse_', 'bam_')  # predict.py:76

_runner = None


def load_model(weight_paths=None):
    """predict.py:17-54: build the five models and try to load their checkpoints; a missing file is
    reported and the model keeps its random initialisation (the reference catches OSError only).
    ``weight_paths``: optional dict name -> path (the reference hard-codes D:\\ paths)."""
    global res_model, hr_model, v3_model, unet_model, bam_model, _runner
    ctors = (("res34", lambda: ResNetFamily().run_model('res34')), ("hrnet", HRNet),
             ("v3plus", Xception_DeepLabV3_Plus), ("scse", lambda: UNet(2)), ("bam", Xception_DeepLabV3_Plus_bam))
    models = []
    for i, (name, ctor) in enumerate(ctors):
        m = ctor()
        try:
            if weight_paths and name in weight_paths:
                m.load_weights(weight_paths[name])
                print('load weights {} {}/5'.format(name, i + 1))
        except OSError as e:
            print('error while loading weights: {}'.format(e))
        models.append(m)
    res_model, hr_model, v3_model, unet_model, bam_model = models
    _runner = None
    return tuple(models)


def _models():
    if res_model is None:
        load_model()
    return [res_model, hr_model, v3_model, unet_model, bam_model]


def runner(batch=None):
    global _runner
    if _runner is None or _runner.batch != batch or _runner.models != _models():
        _runner = S.SceneRunner(_models(), batch=batch)
    return _runner


def detection(img_path, user_path, model, save_name='model', bug_compatible=True):
    """predict.py:90-116: tile the image, predict every tile, OR-stitch the argmax masks and write
    ``user_path/save_name.png`` (0/255, PNG compression 0)."""
    import cv2 as cv
    img = cv.imread(img_path)
    if img is None:
        raise TypeError("cv.imread returned None for {!r}".format(img_path))  # cvtColor(None) raises in the reference
    r = S.SceneRunner([model])
    mask = r.run(r.upload(img), bug_compatible=bug_compatible)[0]
    _write_png0(user_path + '/{}.png'.format(save_name), mask.cpu().numpy())


def run_model(img_path, user_path, name='', bug_compatible=True):
    """predict.py:75-87: the five detections with the reference's file names.  The image is read and
    uploaded once; all five scene masks are produced on the device."""
    import cv2 as cv
    img = cv.imread(img_path)
    if img is None:
        raise TypeError("cv.imread returned None for {!r}".format(img_path))
    r = runner()
    masks = r.run(r.upload(img), bug_compatible=bug_compatible).cpu().numpy()
    for prefix, mask in zip(MODEL_PREFIXES, masks):
        _write_png0(user_path + '/{}.png'.format(prefix + name), mask)


def _write_png0(path, mask):
    """cv.imwrite(path, mask, [IMWRITE_PNG_COMPRESSION, 0]) (predict.py:115) through the multi-threaded level-0 encoder"""
    from . import png0
    with open(path, 'wb') as f:
        f.write(png0.encode_gray(mask))


def write_points(points, path):
    """predict.py:119-132: one polygon per line, 'x,y ' per vertex."""
    with open(path, 'w', encoding='utf-8') as f:
        for xs, ys in points:
            f.write(''.join('{},{} '.format(x, y) for x, y in zip(xs, ys)))
            f.write('\n')


def predict(image, bug_compatible=True, fusion="vote"):
    """In-memory hot path: image (H,W,3) u8 BGR -> (mask (H,W) u8 {0,255}, points [[xs, ys], ...]).
    = run_model -> model_confuse -> _detection (predict.py:148-152) without the PNG hand-offs.
    ``fusion="average"`` (opt-in, not the reference's behaviour): the five probability maps are averaged per pixel and
    thresholded at 0.5 instead of cleaning and voting on the five argmax masks; the final clean-up stays."""
    from . import edge_3, model_fuse
    r = runner()
    if fusion == "average":
        fused = model_fuse.cleanup_device(r.run_average(r.upload(image), bug_compatible=bug_compatible))
        points, _h = edge_3.contours_device(fused)
        return fused.cpu().numpy(), points
    if fusion != "vote":
        raise ValueError("fusion must be 'vote' (the reference's 3-of-5) or 'average'")
    masks = r.run(r.upload(image), bug_compatible=bug_compatible)
    fused = model_fuse.fuse_device(masks)
    points, _h = edge_3.contours_device(fused)
    return fused.cpu().numpy(), points
