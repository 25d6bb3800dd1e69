"""ctypes binding of the C ABI declared in include/asuna_b200.h.

`Library(path, prefix)` binds one shared object; the product uses
`Library.product()` = asuna_b200/libasuna_b200.so with prefix ``asuna_``.  The class is
prefix-generic because the CPU oracle (test infrastructure, oracle/binding.py) exports the
same entry points under ``oracle_`` -- the product never loads it.  There is no CPU
fallback: a missing CUDA library raises.
"""
import ctypes as C
import os

import numpy as np

from . import structs as S

_HERE = os.path.dirname(os.path.abspath(__file__))
# ASUNA_B200_LIB: developer knob for A/B-ing kernel variants on the GPU box (same ABI, another build)
PRODUCT_LIB = os.environ.get("ASUNA_B200_LIB") or os.path.join(_HERE, "libasuna_b200.so")

# every symbol include/asuna_b200.h declares (tests check the .so exports all of them)
ABI_SYMBOLS = (
    "abi_sizes", "create", "destroy", "last_error", "set_film", "add_texture", "set_envmap", "add_mesh",
    "add_material", "set_lights", "add_instance", "build_accel", "set_camera", "set_sunsky", "set_state",
    "reset_frame", "render_frames", "set_partition", "sync", "read_channel", "read_channel_async", "wait_reads", "export_partial",
    "import_partial", "post_process", "host_alloc", "host_free", "channel_device_ptr", "stream_handle", "set_counting", "set_profiling", "get_stats", "reset_stats", "trace_primary", "trace_rays",
    "occlusion_rays", "accel_stats")


class AsunaError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Library:
    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise AsunaError(f"{path} not found -- build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        self.check_abi()

    _product = None

    @classmethod
    def product(cls):
        if cls._product is None:
            cls._product = cls(PRODUCT_LIB, "asuna_")
        return cls._product

    def fn(self, name, restype=C.c_int):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    def has(self, name):
        return hasattr(self.lib, self.prefix + name)

    def check_abi(self):
        sizes = (C.c_uint32 * 6)()
        self.fn("abi_sizes", None)(sizes)
        if tuple(sizes) != S.EXPECTED_SIZES:
            raise AsunaError(f"wire-struct size mismatch: library {tuple(sizes)} vs binding {S.EXPECTED_SIZES}")


class Context:
    """One rendering context == one GPU (or one CPU oracle instance)."""

    def __init__(self, library=None, gpu_id=0):
        self.L = library or Library.product()
        self.h = C.c_void_p()
        rc = self.L.fn("create")(C.byref(self.h), C.c_int(gpu_id))
        if rc != 0 or not self.h:
            self.h = C.c_void_p()
            raise AsunaError(f"{self.L.prefix}create failed with {rc} (no CUDA device?)")
        self.width = self.height = 0
        self._pinned = []

    def close(self):
        if self.h:
            for p in self._pinned:
                self.L.fn("host_free")(self.h, p)
            self._pinned = []
            self.L.fn("destroy", None)(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        rc = self.L.fn(name)(self.h, *args)
        if rc < 0:
            msg = self.L.fn("last_error", C.c_char_p)(self.h)
            raise AsunaError(f"{self.L.prefix}{name} -> {rc}: {msg.decode() if msg else ''}")
        return rc

    # ---- scene upload
    def set_film(self, w, h):
        self.width, self.height = int(w), int(h)
        return self._call("set_film", C.c_uint32(w), C.c_uint32(h))

    def add_texture(self, rgba):
        a = np.ascontiguousarray(rgba, np.float32)
        assert a.ndim == 3 and a.shape[2] == 4
        return self._call("add_texture", _ptr(a), C.c_uint32(a.shape[1]), C.c_uint32(a.shape[0]))

    def set_envmap(self, rgba, marginal, conditional):
        a, m, c = (np.ascontiguousarray(x, np.float32) for x in (rgba, marginal, conditional))
        assert a.shape == m.shape == c.shape and a.shape[2] == 4
        return self._call("set_envmap", _ptr(a), _ptr(m), _ptr(c), C.c_uint32(a.shape[1]), C.c_uint32(a.shape[0]))

    def add_mesh(self, vertices, indices):
        v = np.ascontiguousarray(vertices, S.Vertex)
        i = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        return self._call("add_mesh", _ptr(v), C.c_uint32(v.size), _ptr(i), C.c_uint32(i.size))

    def add_material(self, material):
        m = np.ascontiguousarray(material, S.Material)
        return self._call("add_material", _ptr(m))

    def set_lights(self, lights):
        l = np.ascontiguousarray(lights, S.Light).reshape(-1)
        return self._call("set_lights", _ptr(l), C.c_uint32(l.size))

    def add_instance(self, xform_colmajor16, mesh, material, light_id=-1):
        x = np.ascontiguousarray(xform_colmajor16, np.float32).reshape(16)
        return self._call("add_instance", _ptr(x), C.c_uint32(mesh), C.c_uint32(material), C.c_int32(light_id))

    def build_accel(self):
        ms = C.c_float()
        self._call("build_accel", C.byref(ms))
        return ms.value

    # ---- per shot / frame
    def set_camera(self, cam):
        c = np.ascontiguousarray(cam, S.Camera)
        return self._call("set_camera", _ptr(c))

    def set_sunsky(self, ss):
        s = np.ascontiguousarray(ss, S.SunSky)
        return self._call("set_sunsky", _ptr(s))

    def set_state(self, st):
        s = np.ascontiguousarray(st, S.State)
        return self._call("set_state", _ptr(s))

    def reset_frame(self):
        return self._call("reset_frame")

    def render_frames(self, n):
        return self._call("render_frames", C.c_uint32(n))

    def set_partition(self, rank, world):
        return self._call("set_partition", C.c_uint32(rank), C.c_uint32(world))

    def sync(self):
        return self._call("sync")

    def read_channel(self, ch, out=None):
        """Image `ch` as (h, w, 4) float32.  `out` (e.g. from pinned_image()) is filled in place when given."""
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        assert out.dtype == np.float32 and out.size == self.height * self.width * 4 and out.flags.c_contiguous
        self._call("read_channel", C.c_int(ch), _ptr(out))
        return out

    def read_channel_async(self, ch, out):
        """Queues the read of image `ch` into the page-locked array `out` (pinned_image()) and returns at once; `out` is
        valid after wait_reads()."""
        assert out.dtype == np.float32 and out.size == self.height * self.width * 4 and out.flags.c_contiguous
        self._call("read_channel_async", C.c_int(ch), _ptr(out))

    def wait_reads(self):
        self._call("wait_reads")

    def post_process(self, post):
        """Tone-mapped radiance image, (h, w, 4) float32 (≙ PipelinePost::run + offline colour read-back)."""
        p = np.array(post, S.Post)
        out = np.empty((self.height, self.width, 4), np.float32)
        self._call("post_process", _ptr(p), _ptr(out))
        return out

    def pinned_image(self):
        """A page-locked (h, w, 4) float32 array owned by the library (asuna_host_alloc); freed with the context."""
        n = self.height * self.width * 4
        p = C.c_void_p()
        self._call("host_alloc", C.c_size_t(n * 4), C.byref(p))
        self._pinned.append(p)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).reshape(self.height, self.width, 4)

    def export_partial(self):
        p = C.c_void_p()
        self._call("export_partial", C.byref(p))
        return p.value

    def import_partial(self):
        return self._call("import_partial")

    def channel_device_ptr(self, ch):
        p = C.c_void_p()
        self._call("channel_device_ptr", C.c_int(ch), C.byref(p))
        return p.value

    def stream_handle(self):
        p = C.c_void_p()
        self._call("stream_handle", C.byref(p))
        return p.value or 0

    def set_counting(self, on):
        return self._call("set_counting", C.c_int(1 if on else 0))

    def set_profiling(self, on):
        return self._call("set_profiling", C.c_int(1 if on else 0))

    def stats(self):
        s = np.zeros((), S.Stats)
        self._call("get_stats", _ptr(s))
        return {k: s[k].item() for k in S.Stats.names if k != "pad"}

    def reset_stats(self):
        return self._call("reset_stats")

    # ---- introspection
    def trace_primary(self):
        ip = np.empty((self.height, self.width, 2), np.uint32)
        t = np.empty((self.height, self.width), np.float32)
        self._call("trace_primary", _ptr(ip), _ptr(t))
        return ip, t

    def trace_rays(self, rays):
        r = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        tuv = np.empty((r.shape[0], 3), np.float32)
        ip = np.empty((r.shape[0], 2), np.uint32)
        self._call("trace_rays", _ptr(r), C.c_uint32(r.shape[0]), _ptr(tuv), _ptr(ip))
        return tuv, ip

    def occlusion_rays(self, rays):
        r = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        occ = np.empty(r.shape[0], np.uint8)
        self._call("occlusion_rays", _ptr(r), C.c_uint32(r.shape[0]), _ptr(occ))
        return occ

    def accel_stats(self):
        out = (C.c_uint64 * 4)()
        self._call("accel_stats", out)
        return {"nodes": out[0], "leaf_prims": out[1], "tlas_nodes": out[2], "sah_cost": out[3] / 1000.0}
