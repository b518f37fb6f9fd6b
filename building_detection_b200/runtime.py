"""ctypes binding of the C ABI (include/bd_b200.h) and the native plan object.

The shared library is built in-tree (``csrc/build.sh`` -> ``libbd_b200.so``).  If it is missing, or
there is no B200, every entry point raises: the product has no CPU or eager fallback."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import graph as G

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbd_b200.so")
MAX_TAPS = 18


class TRef(C.Structure):
    _fields_ = [("buf", C.c_int32), ("c0", C.c_int32), ("c", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("x", TRef), ("y", TRef), ("res", TRef), ("ntaps", C.c_int32),
                ("dy", C.c_int32 * MAX_TAPS), ("dx", C.c_int32 * MAX_TAPS),
                ("stride", C.c_int32), ("ho", C.c_int32), ("wo", C.c_int32),
                ("act_pre", C.c_int32), ("act_post", C.c_int32),
                ("out_scale", C.c_int32), ("out_oy", C.c_int32), ("out_ox", C.c_int32),
                ("path", C.c_int32), ("w_host", C.c_void_p), ("bias_host", C.c_void_p),
                ("dw_w_host", C.c_void_p), ("dw_relu_in", C.c_int32)]


class PostConstants(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("fuse_min_area", "fuse_min_fragment", "fuse_split_width", "fuse_votes",
                                         "edge_min_area", "edge_min_fragment", "edge_split_width")] + \
               [(n, C.c_double) for n in ("edge_iou", "edge_min_moment", "tier_small", "tier_mid", "tier_big0", "tier_big1",
                                          "tier_big2", "eps_default", "eps_mid_mult", "eps_big0", "eps_big1", "eps_big2")]


class Polys(C.Structure):
    _fields_ = [("n_polys", C.c_int32), ("n_points", C.c_int32), ("offsets", C.POINTER(C.c_int32)),
                ("xs", C.POINTER(C.c_float)), ("ys", C.POINTER(C.c_float)), ("is_float", C.POINTER(C.c_uint8))]


# name -> (restype, argtypes); mirrors include/bd_b200.h one to one
_SIGS = {
    "bd_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "bd_destroy": (None, [C.c_void_p]),
    "bd_last_error": (C.c_char_p, []),
    "bd_version": (C.c_char_p, []),
    "bd_launch_count": (C.c_int64, [C.c_void_p]),
    "bd_debug_read_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "bd_plan_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "bd_plan_destroy": (None, [C.c_void_p]),
    "bd_plan_add_buffer": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "bd_plan_add_conv": (C.c_int, [C.c_void_p, C.POINTER(ConvDesc)]),
    "bd_plan_add_dwconv": (C.c_int, [C.c_void_p, TRef, TRef, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "bd_plan_add_maxpool": (C.c_int, [C.c_void_p, TRef, TRef, C.c_int, C.c_int, C.c_int, C.c_int]),
    "bd_plan_add_addn": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(TRef), C.POINTER(C.c_int32), TRef, C.c_int]),
    "bd_plan_add_gap": (C.c_int, [C.c_void_p, TRef, C.c_int]),
    "bd_plan_add_dense": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p]),
    "bd_plan_add_gate": (C.c_int, [C.c_void_p, C.c_int, TRef, TRef, C.c_int, TRef, C.c_void_p, C.c_float]),
    "bd_plan_add_skfuse": (C.c_int, [C.c_void_p, C.POINTER(TRef), C.c_int, C.POINTER(C.c_int32), TRef, C.c_void_p,
                                     C.c_void_p]),
    "bd_plan_add_bcast": (C.c_int, [C.c_void_p, C.c_int, TRef]),
    "bd_plan_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "bd_plan_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "bd_plan_load": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "bd_plan_input_stride": (C.c_int, [C.c_void_p]),
    "bd_plan_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bd_plan_run_head": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bd_plan_run_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bd_plan_buffer_ptr": (C.c_void_p, [C.c_void_p, C.c_int]),
    "bd_plan_buffer_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "bd_plan_read_buffer": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "bd_plan_write_buffer": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "bd_plan_arena_bytes": (C.c_size_t, [C.c_void_p]),
    "bd_plan_arena_bytes_flat": (C.c_size_t, [C.c_void_p]),
    "bd_plan_set_arena_reuse": (C.c_int, [C.c_void_p, C.c_int]),
    "bd_plan_buffer_lifetime": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bd_plan_num_launches": (C.c_int, [C.c_void_p]),
    "bd_plan_num_ops": (C.c_int, [C.c_void_p]),
    "bd_plan_time_ops": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "bd_plan_op_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "bd_tiles_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_int, C.c_void_p]),
    "bd_stitch_or": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                               C.c_void_p]),
    "bd_tiles_set_origins": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "bd_tiles_gather_at": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p]),
    "bd_stitch_or_at": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "bd_scene_run": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_int, C.c_void_p, C.c_void_p]),
    "bd_scene_run_average": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "bd_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "bd_plan_uses_graph": (C.c_int, [C.c_void_p]),
    "bd_fuse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bd_fuse_cleaned": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bd_mask_cleanup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bd_post_constants": (C.c_int, [C.c_void_p, C.POINTER(PostConstants)]),
    "bd_plane_words_per_row": (C.c_size_t, [C.c_int]),
    "bd_mask_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bd_mask_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bd_fuse_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bd_debug_cleanup_stage": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bd_debug_labels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bd_contours": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Polys), C.c_void_p]),
    "bd_polys_free": (None, [C.POINTER(Polys)]),
    "bd_png0_size": (C.c_size_t, [C.c_int, C.c_int]),
    "bd_png0_encode": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]),
    "bd_host_contour_area": (C.c_double, [C.c_void_p, C.c_int]),
    "bd_host_arc_length": (C.c_double, [C.c_void_p, C.c_int]),
    "bd_host_approx_poly": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p]),
    "bd_host_simplify": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Load libbd_b200.so (once).  Fails loudly when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} is missing: build it with building_detection_b200/csrc/build.sh "
                              "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise NativeError(lib().bd_last_error().decode(errors="replace"))


_ctx = {}


def resolve_device(device=None):
    """None -> BD_DEVICE, else LOCAL_RANK (torchrun: one process per GPU), else 0."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if "BD_DEVICE" not in os.environ else int(os.environ["BD_DEVICE"])
    return int(device)


def context(device=None):
    """One bd_ctx per GPU (per process)."""
    device = resolve_device(device)
    if device not in _ctx:
        h = C.c_void_p()
        check(lib().bd_create(device, C.byref(h)))
        _ctx[device] = h
    return _ctx[device]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _tref(r):
    return TRef(*r) if r is not None else TRef(-1, 0, 0)


class NativePlan:
    """A graph.Plan uploaded through the C ABI: arena + weights + launch list on one GPU."""

    def __init__(self, plan: G.Plan, device=None, reuse=False):
        """``reuse``: map buffers with disjoint lifetimes share arena ranges (bd_plan_set_arena_reuse; what
        Model.native_plan asks for).  Off here by default: tests and tools read intermediates back."""
        L = lib()
        self.plan = plan
        self.ctx = context(device)
        self.h = C.c_void_p()
        check(L.bd_plan_create(self.ctx, plan.batch, C.byref(self.h)))
        check(L.bd_plan_set_arena_reuse(self.h, 1 if reuse else 0))
        self.reuse = bool(reuse)
        try:
            self._upload(L, plan)
        except Exception:
            self.close()
            raise

    def _upload(self, L, plan):
        keep = []  # keep numpy temporaries alive until finalize
        for b in plan.bufs:
            bid = L.bd_plan_add_buffer(self.h, b.H, b.W, b.C, 1 if b.dtype == "f32" else 0, 1 if b.kind == "vec" else 0)
            if bid != b.id:
                raise NativeError("buffer id mismatch: " + L.bd_last_error().decode())
        fuse_sep = os.environ.get("BD_FUSE_SEPCONV", "1") != "0"
        self.native_to_plan = []  # native op index -> plan op index (fused depthwise stages have no native op)
        for i_op, op in enumerate(plan.ops):
            kind = op["op"]
            if not (kind == G.OP_DWCONV and op.get("fuse") and fuse_sep):
                self.native_to_plan.append(i_op)
            if kind == G.OP_CONV:
                d = ConvDesc()
                d.x, d.y, d.res = _tref(op["x"]), _tref(op["y"]), _tref(op["res"])
                d.ntaps = len(op["taps"])
                for i, (dy, dx) in enumerate(op["taps"]):
                    d.dy[i], d.dx[i] = dy, dx
                d.stride, d.ho, d.wo = op["stride"], op["Ho"], op["Wo"]
                d.act_pre, d.act_post = op["act_pre"], op["act_post"]
                d.out_scale, d.out_oy, d.out_ox = op["out_scale"], op["out_oy"], op["out_ox"]
                d.path = {"direct": 0, "umma": 1, "small": 2}[op["path"]]
                w = np.ascontiguousarray(op["w"], np.uint16)
                b = np.ascontiguousarray(op["b"], np.float32)
                keep += [w, b]
                d.w_host, d.bias_host = w.ctypes.data, b.ctypes.data
                d.dw_w_host, d.dw_relu_in = None, 0
                if op.get("fused_dw") is not None and fuse_sep:
                    # SeparableConv2D in one kernel: x becomes the depthwise input (whole padded buffer when the
                    # pointwise stage reads the padded intermediate), depthwise weights padded to the same width
                    dwop = plan.ops[op["fused_dw"]]
                    cin_k = op["x"][2]
                    dww = np.zeros((9, cin_k), np.float32)
                    dww[:, :dwop["x"][2]] = dwop["w"]
                    keep.append(dww)
                    d.x = _tref((dwop["x"][0], dwop["x"][1], cin_k))
                    d.dw_w_host, d.dw_relu_in = dww.ctypes.data, int(dwop["relu_in"])
                check(L.bd_plan_add_conv(self.h, C.byref(d)))
            elif kind == G.OP_DWCONV and op.get("fuse") and fuse_sep:
                pass  # runs inside the pointwise convolution that follows
            elif kind == G.OP_DWCONV:
                w = np.ascontiguousarray(op["w"], np.float32)
                keep.append(w)
                check(L.bd_plan_add_dwconv(self.h, _tref(op["x"]), _tref(op["y"]), op["stride"], op["pad_t"],
                                           op["pad_l"], op["relu_in"], _ptr(w)))
            elif kind == G.OP_MAXPOOL:
                check(L.bd_plan_add_maxpool(self.h, _tref(op["x"]), _tref(op["y"]), op["k"], op["stride"],
                                            op["pad_t"], op["pad_l"]))
            elif kind == G.OP_ADDN:
                n = len(op["xs"])
                xs = (TRef * n)(*[_tref(r) for r in op["xs"]])
                fs = (C.c_int32 * n)(*op["fs"])
                check(L.bd_plan_add_addn(self.h, n, xs, fs, _tref(op["y"]), op["act"]))
            elif kind == G.OP_GAP:
                check(L.bd_plan_add_gap(self.h, _tref(op["x"]), op["y"]))
            elif kind == G.OP_DENSE:
                n = len(op["xs"])
                w = np.ascontiguousarray(op["w"], np.float32)
                b = np.ascontiguousarray(op["b"], np.float32)
                keep += [w, b]
                check(L.bd_plan_add_dense(self.h, n, (C.c_int32 * n)(*op["xs"]), op["y"], w.shape[1], w.shape[0],
                                          op["act"], _ptr(w), _ptr(b)))
            elif kind == G.OP_GATE:
                w = None if op["w"] is None else np.ascontiguousarray(op["w"], np.float32)
                keep.append(w)
                check(L.bd_plan_add_gate(self.h, op["mode"], _tref(op["x"]), _tref(op["y"]), op["v"], _tref(op["s"]),
                                         None if w is None else _ptr(w), float(op["b"])))
            elif kind == G.OP_SKFUSE:
                xs = (TRef * 4)(*[_tref(r) for r in op["xs"]])
                sc = np.ascontiguousarray(op["scale"], np.float32)
                sh = np.ascontiguousarray(op["shift"], np.float32)
                keep += [sc, sh]
                check(L.bd_plan_add_skfuse(self.h, xs, op["g"], (C.c_int32 * 5)(*op["logits"]), _tref(op["y"]),
                                           _ptr(sc), _ptr(sh)))
            elif kind == G.OP_BCAST:
                check(L.bd_plan_add_bcast(self.h, op["v"], _tref(op["y"])))
            elif kind == G.OP_SOFTMAX2:
                pass  # appended by bd_plan_finalize from (logits_buf, logits_up)
            else:
                raise NativeError(f"unknown plan op {kind}")
        check(L.bd_plan_finalize(self.h, plan.input, plan.logits, plan.logits_up))
        del keep

    # ------------------------------------------------------------------ execution
    def run_host(self, x, want_probs=True, want_mask=False):
        """x: (batch,512,512,3) float32 host array -> probs (batch,512,512,2) [, mask (batch,512,512) u8]."""
        n = self.plan.batch
        x = np.ascontiguousarray(x, np.float32)
        assert x.shape[0] == n, (x.shape, n)
        lb = self.plan.bufs[self.plan.logits]
        H, W = lb.H * self.plan.logits_up, lb.W * self.plan.logits_up
        probs = np.empty((n, H, W, 2), np.float32) if want_probs else None
        mask = np.empty((n, H, W), np.uint8) if want_mask else None
        check(lib().bd_plan_run_host(self.h, _ptr(x), None if probs is None else _ptr(probs),
                                     None if mask is None else _ptr(mask)))
        if want_probs and want_mask:
            return probs, mask
        return probs if want_probs else mask

    def run_device(self, x_ptr, probs_ptr, mask_ptr, stream=0):
        """Device pointers (ints, 0 = absent); asynchronous on ``stream``."""
        check(lib().bd_plan_run(self.h, x_ptr or None, probs_ptr or None, mask_ptr or None, stream or None))

    def buffer_ptr(self, buf):
        return lib().bd_plan_buffer_ptr(self.h, buf)

    def read_buffer(self, buf):
        """Whole plan buffer as float32 numpy (fp16 maps are widened)."""
        b = self.plan.bufs[buf]
        n = self.plan.batch
        if b.kind == "vec":
            out = np.empty((n, b.C), np.float32)
            check(lib().bd_plan_read_buffer(self.h, buf, _ptr(out), out.nbytes))
            return out
        if b.dtype == "f32":
            out = np.empty((n, b.H, b.W, b.C), np.float32)
            check(lib().bd_plan_read_buffer(self.h, buf, _ptr(out), out.nbytes))
            return out
        raw = np.empty((n, b.H, b.W, b.C), np.uint16)
        check(lib().bd_plan_read_buffer(self.h, buf, _ptr(raw), raw.nbytes))
        return G.h16_to_f32(raw)

    def write_buffer(self, buf, arr):
        b = self.plan.bufs[buf]
        if b.kind == "vec" or b.dtype == "f32":
            a = np.ascontiguousarray(arr, np.float32)
        else:
            a = G.to_h16(arr)
        check(lib().bd_plan_write_buffer(self.h, buf, _ptr(a), a.nbytes))

    def time_ops(self, stream=0):
        """Per plan op: device ms (CUDA events around every native op), kernel class, algorithmic FLOPs.  A depthwise
        stage fused into its pointwise convolution reports 0 ms; its FLOPs are counted on the fused op."""
        n = lib().bd_plan_num_ops(self.h)
        assert n == len(self.native_to_plan), (n, len(self.native_to_plan))
        ms_n = np.zeros(n, np.float32)
        check(lib().bd_plan_time_ops(self.h, _ptr(ms_n), stream or None))
        nplan = len(self.plan.ops)
        ms = np.zeros(nplan, np.float32)
        kinds = np.full(nplan, 2, np.int32)
        flops = np.zeros(nplan, np.float64)
        native_set = set(self.native_to_plan)
        for i in range(n):
            k, f = C.c_int(), C.c_double()
            check(lib().bd_plan_op_info(self.h, i, C.byref(k), C.byref(f)))
            j = self.native_to_plan[i]
            ms[j], kinds[j], flops[j] = ms_n[i], k.value, f.value
            op = self.plan.ops[j]
            if "flops" in op:
                flops[j] = op["flops"]  # algorithmic: the native count includes channel padding
            if op.get("fused_dw") is not None and op["fused_dw"] not in native_set:
                dw = self.plan.ops[op["fused_dw"]]
                flops[j] += 2.0 * self.plan.batch * op["Ho"] * op["Wo"] * dw["x"][2] * 9
        return ms, kinds, flops

    def save(self, path):
        """Write the plan file (bd_plan_save): what a non-Python host loads with bd_plan_load."""
        check(lib().bd_plan_save(self.h, os.fsencode(path)))

    @property
    def uses_graph(self):
        return bool(lib().bd_plan_uses_graph(self.h))

    @property
    def num_launches(self):
        return lib().bd_plan_num_launches(self.h)

    @property
    def arena_bytes(self):
        return lib().bd_plan_arena_bytes(self.h)

    def buffer_lifetime(self, buf):
        """(first step, last step, shared) of a buffer as the arena allocator saw it."""
        f, l, sh = C.c_int(), C.c_int(), C.c_int()
        check(lib().bd_plan_buffer_lifetime(self.h, buf, C.byref(f), C.byref(l), C.byref(sh)))
        return f.value, l.value, bool(sh.value)

    @property
    def arena_bytes_flat(self):
        """Arena size without buffer reuse (one range per buffer)."""
        return lib().bd_plan_arena_bytes_flat(self.h)

    def close(self):
        if self.h:
            lib().bd_plan_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
