"""Host-side pieces either side of the hot path (SURVEY section 8b HTTP row, 8f item 2): the POST /photo document of
buildAPI.py:121-149 and the level-0 PNG codec of the stage hand-offs.  CPU only."""
import base64
import io
import json

import cv2 as cv
import numpy as np
import pytest

from building_detection_b200 import buildAPI, png0


@pytest.mark.parametrize("shape", [(1, 1), (7, 5), (300, 217), (256, 256), (255, 256), (1, 65534), (65535, 1), (1000, 1311)])
def test_png0_round_trip_and_opencv_interop(shape):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    m = (rng.random(shape) < 0.4).astype(np.uint8) * 255
    data = png0.encode_gray(m)
    assert data == png0.encode_gray_py(m) == png0.encode_gray(m, threads=3)  # native (all cores / 3 threads) == numpy twin
    np.testing.assert_array_equal(png0.decode_gray(data), m)
    np.testing.assert_array_equal(cv.imdecode(np.frombuffer(data, np.uint8), cv.IMREAD_UNCHANGED), m)   # libpng reads ours
    ok, theirs = cv.imencode(".png", m, [int(cv.IMWRITE_PNG_COMPRESSION), 0])                            # what the reference writes
    assert ok
    np.testing.assert_array_equal(png0.decode_gray(theirs.tobytes()), m)                                # we read libpng's
    ok, packed = cv.imencode(".png", m, [int(cv.IMWRITE_PNG_COMPRESSION), 6])                            # compressed: fallback path
    np.testing.assert_array_equal(png0.decode_gray(packed.tobytes()), m)
    assert len(data) < m.size + m.shape[0] + 5 * (m.size // 65535 + 2) + 80


def test_png0_rejects_garbage():
    with pytest.raises(ValueError):
        png0.decode_gray(b"not a png at all")
    good = bytearray(png0.encode_gray(np.full((4, 4), 255, np.uint8)))
    good[60] ^= 0xFF  # flip a pixel: the stream checksum no longer matches
    with pytest.raises(ValueError):
        png0.decode_gray(bytes(good))


def test_build_response_literal_document():
    """buildAPI.py:121-147 on a fixed mask and two polygons (int32 coordinates as edge_3 returns them)."""
    mask = np.zeros((2, 3), np.uint8)
    mask[0, 1] = 255
    png = png0.encode_gray(mask)
    points = [[[np.int32(1), np.int32(4), np.int32(1)], [np.int32(2), np.int32(5), np.int32(2)]],
              [[np.float32(1.5), np.float32(2.0)], [np.float32(0.25), np.float32(3.0)]]]
    doc = buildAPI.build_response(png, points)
    want = {"status": "success", "data": base64.b64encode(png).decode(), "error": "None",
            "points": {"0": "1,2 4,5 1,2 ", "1": "{},{} {},{} ".format(np.float32(1.5), np.float32(0.25), np.float32(2.0), np.float32(3.0))}}
    assert doc == want
    assert json.loads(buildAPI.dumps(doc)) == want
    # the client's view (CLient/Client.py:47-66): status, points, data -> image bytes
    back = json.loads(buildAPI.dumps(doc))
    np.testing.assert_array_equal(png0.decode_gray(base64.b64decode(back["data"])), mask)
    # mismatching coordinate lists -> NG document (:134-137)
    bad = buildAPI.build_response(png, [[[1, 2], [3]]])
    assert bad["status"] == "NG" and bad["data"] is None and bad["points"] == {}
    # points = None (contour stage failed, :116-119): the reference's loop raises, its except answers NG
    with pytest.raises(TypeError):
        buildAPI.build_response(png, None)
    assert buildAPI.error_response(ValueError("x")) == {"status": "NG", "data": None, "points": {}, "error": "x"}


def test_wsgi_app_contract_without_gpu(monkeypatch):
    """POST /photo multipart plumbing; the pipeline itself is replaced (no GPU here) to check the transport only."""
    mask = np.zeros((5, 6), np.uint8)
    mask[1:3, 2:5] = 255
    seen = {}

    def fake_predict(image, bug_compatible=True):
        seen["shape"] = image.shape
        return mask, [[[2, 4, 2], [1, 2, 1]]]
    from building_detection_b200 import predict
    monkeypatch.setattr(predict, "predict", fake_predict)
    img = np.random.default_rng(0).integers(0, 255, (5, 6, 3), dtype=np.uint8)
    ok, enc = cv.imencode(".png", img)
    body = (b"--XYZ\r\nContent-Disposition: form-data; name=\"file\"; filename=\"a.png\"\r\nContent-Type: image/png\r\n\r\n"
            + enc.tobytes() + b"\r\n--XYZ--\r\n")
    status = {}
    env = {"REQUEST_METHOD": "POST", "PATH_INFO": "/photo", "CONTENT_TYPE": "multipart/form-data; boundary=XYZ",
           "CONTENT_LENGTH": str(len(body)), "wsgi.input": io.BytesIO(body), "HTTP_CLIENTID": "1.2.3.4"}
    out = b"".join(buildAPI.app(env, lambda s, h: status.update(s=s, h=dict(h))))
    doc = json.loads(out.decode("utf-8"))
    assert status["s"].startswith("200") and seen["shape"] == (5, 6, 3)
    assert doc["status"] == "success" and doc["points"] == {"0": "2,1 4,2 2,1 "} and doc["error"] == "None"
    np.testing.assert_array_equal(png0.decode_gray(base64.b64decode(doc["data"])), mask)
    # undecodable upload -> NG, missing field -> NG, pipeline exception -> NG (never an HTTP error)
    for b in (b"--XYZ\r\nContent-Disposition: form-data; name=\"file\"\r\n\r\njunk\r\n--XYZ--\r\n",
              b"--XYZ\r\nContent-Disposition: form-data; name=\"other\"\r\n\r\njunk\r\n--XYZ--\r\n"):
        env.update({"CONTENT_LENGTH": str(len(b)), "wsgi.input": io.BytesIO(b)})
        doc = json.loads(b"".join(buildAPI.app(env, lambda s, h: None)).decode())
        assert doc == {"status": "NG", "data": None, "points": {}, "error": doc["error"]} and doc["error"]
    monkeypatch.setattr(predict, "predict", lambda *a, **k: (_ for _ in ()).throw(RuntimeError("boom")))
    assert buildAPI.handle_photo(img) == {"status": "NG", "data": None, "points": {}, "error": "boom"}
    env.update({"REQUEST_METHOD": "GET"})
    assert buildAPI.app(env, lambda s, h: status.update(s=s)) == [b"POST /photo"] and status["s"].startswith("404")


def test_constants_dataclass_matches_reference_literals():
    from building_detection_b200.constants import DEFAULT
    assert (DEFAULT.fuse_min_area, DEFAULT.fuse_min_fragment, DEFAULT.fuse_split_width, DEFAULT.fuse_votes) == (1000, 500, 21, 3)
    assert (DEFAULT.edge_min_area, DEFAULT.edge_min_fragment, DEFAULT.edge_split_width) == (100, 50, 7)
    assert (DEFAULT.tier_small, DEFAULT.tier_mid, DEFAULT.tier_big0, DEFAULT.tier_big1, DEFAULT.tier_big2) == (150, 300, 3000, 8000, 15000)
