"""Drop-in for the POST contract of the reference's buildAPI.py (``POST /photo``, buildAPI.py:82-149) without the
Flask dependency: ``handle_photo`` is get_frame() as a function -- image bytes in, the JSON document out -- and
``build_response`` is the part that assembles the document (buildAPI.py:121-147).  ``app`` is a WSGI callable
(standard library only) that serves it; the reference's own server is not functional as written (SURVEY App. D
#3, #4: it reads result.png while the fuse stage writes '\\_result.png', passes ``encoding=`` to json.dumps and puts
``bytes`` into the document), so the document here is what its client expects to parse (CLient/Client.py:47-66):

    {"status": "success", "data": <base64 of the result-mask PNG>, "points": {"0": "x,y x,y ... ", ...}, "error": "None"}
    {"status": "NG", "data": null, "points": {}, "error": <message>}
"""
from __future__ import annotations

import base64
import json

import numpy as np

from . import png0


def error_response(err):
    """buildAPI.py:100-102,134-137,148-149: the NG document."""
    return {"status": "NG", "data": None, "points": {}, "error": str(err)}


def build_response(mask_png, points):
    """buildAPI.py:121-147.  mask_png: the bytes of the result PNG (what the reference reads back from disk, :122);
    points: what edge_3._detection returned, [[xs, ys], ...].  The base64 payload is returned as ``str`` so that the
    document serialises (the reference leaves it as ``bytes``, which json.dumps rejects)."""
    data = {"status": "success", "data": base64.b64encode(bytes(mask_png)).decode("ascii"), "points": {}}
    for i in range(len(points)):  # TypeError when points is None (the reference's except turns that into NG)
        point_x, point_y = points[i][0], points[i][1]
        if len(point_x) != len(point_y):
            return error_response('轮廓优化时出现错误，请检查服务端 edge_3.py文件')  # :134-137
        data["points"]["{}".format(i)] = "".join("{},{} ".format(x, y) for x, y in zip(point_x, point_y))  # :138-143
    data["error"] = "None"
    return data


def handle_photo(image, bug_compatible=True):
    """get_frame() minus the transport: ``image`` is the uploaded file's bytes (any format cv.imdecode reads) or an
    (H,W,3) u8 BGR array.  Runs run_model -> model_confuse -> _detection in memory (predict.predict) and returns the
    response dict.  Any exception becomes the NG document, as in the reference (:148-149); a failing contour stage
    alone yields points = None there (:116-119), which its loop then turns into NG as well."""
    try:
        from . import predict
        if not isinstance(image, np.ndarray):
            import cv2 as cv
            image = cv.imdecode(np.frombuffer(bytes(image), np.uint8), cv.IMREAD_COLOR)
            if image is None:
                return error_response('传入的图片错误')  # :100
        mask, points = predict.predict(image, bug_compatible=bug_compatible)
        return build_response(png0.encode_gray(mask), points)
    except Exception as e:  # noqa: BLE001 -- the contract is "never raise, answer NG"
        return error_response(e)


def dumps(doc):
    """json.dumps(data, ensure_ascii=False) (:147); numpy scalars in the point strings were formatted already."""
    return json.dumps(doc, ensure_ascii=False)


def app(environ, start_response):
    """WSGI: POST /photo with a multipart field ``file`` (buildAPI.py:82-99); header ``clientID`` is accepted and
    ignored (the reference uses it as a scratch directory name).  ``wsgiref.simple_server.make_server('', 5000, app)``."""
    if environ.get("REQUEST_METHOD") != "POST" or environ.get("PATH_INFO", "").rstrip("/") != "/photo":
        start_response("404 Not Found", [("Content-Type", "text/plain")])
        return [b"POST /photo"]
    try:
        n = int(environ.get("CONTENT_LENGTH") or 0)
        body = environ["wsgi.input"].read(n)
        upload = _multipart_file(environ.get("CONTENT_TYPE", ""), body)
        doc = error_response('传入的图片错误') if upload is None else handle_photo(upload)
    except Exception as e:  # noqa: BLE001
        doc = error_response(e)
    out = dumps(doc).encode("utf-8")
    start_response("200 OK", [("Content-Type", "application/json; charset=utf-8"), ("Content-Length", str(len(out)))])
    return [out]


def _multipart_file(content_type, body):
    """bytes of the part named ``file`` of a multipart/form-data body (None if absent)."""
    if "boundary=" not in content_type:
        return None
    boundary = content_type.split("boundary=", 1)[1].split(";")[0].strip().strip('"').encode()
    for part in body.split(b"--" + boundary):
        head, sep, data = part.partition(b"\r\n\r\n")
        if sep and b'name="file"' in head:
            return data[:-2] if data.endswith(b"\r\n") else data
    return None
