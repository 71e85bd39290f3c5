"""Python mirror of the reference's host API for this path, over the C-ABI library.

Names, argument meaning and error behaviour follow libfloor (file:line relative to a2flo/floor):
  device_context::create_image / create_queue ... include/floor/device/device_context.hpp:106-116, 261-331
  device_queue::finish / start|stop_profiling ..... include/floor/device/device_queue.hpp:101-108, src/device/cuda/cuda_queue.cpp:58-70
  device_image (getters, write, map, unmap, generate_mip_map_chain) ... include/floor/device/device_image.hpp:116-162, 244-326
  cuda_image life-cycle (ctor upload -> chain, write -> chain, unmap -> chain) ... src/device/cuda/cuda_image.cpp:533-536, 667-670, 803-806
The C++ drop-in with the same shape lives in include/floor_b200/ (that is what a libfloor user links).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import image_types as it
from .image_types import IMAGE_TYPE, MEMORY_FLAG


def _L():
    from . import lib
    return lib()


def _check(rc):
    from . import check
    return check(rc)


def _u32x(vals, n):
    v = list(vals) + [0] * (n - len(vals))
    return (ctypes.c_uint32 * n)(*v[:n])


class device:
    """fl::device / cuda_device subset (include/floor/device/device.hpp:76-230)."""

    def __init__(self, index: int):
        from . import DeviceInfo
        info = DeviceInfo()
        _check(_L().flmip_get_device_info(index, ctypes.byref(info)))
        self.index = index
        self.name = info.name.decode(errors="replace")
        self.units = info.units
        self.global_mem_size = info.global_mem_size
        self.max_total_local_size = info.max_total_local_size
        self.max_image_2d_dim = tuple(info.max_image_2d_dim)
        self.max_image_3d_dim = tuple(info.max_image_3d_dim)
        self.max_mip_levels = info.max_mip_levels
        self.sm = (info.sm_major, info.sm_minor)
        self.driver_version = info.driver_version
        self.clock = info.clock_mhz
        self.mem_clock = info.mem_clock_mhz
        self.mem_bus_width = info.mem_bus_width
        self.l2_cache_size = info.l2_cache_size
        self.image_support = True
        self.image_mipmap_support = True
        self.image_mipmap_write_support = True
        # the reference CUDA device reports no cube write support (cuda_device.cpp:54-57); linear images lift that
        self.image_cube_write_support = True
        self.image_cube_array_write_support = True


class pinned_buffer:
    """page-locked host staging buffer exposed as a numpy uint8 array"""

    def __init__(self, size: int, device_index: int = 0, write_combined: bool = False):
        from . import HOST_WRITE_COMBINED
        self._dev = device_index
        self._ptr = ctypes.c_void_p()
        _check(_L().flmip_host_alloc_ex(device_index, size, HOST_WRITE_COMBINED if write_combined else 0, ctypes.byref(self._ptr)))
        self.size = size
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * size).from_address(self._ptr.value))

    @property
    def ptr(self) -> int:
        return self._ptr.value

    def free(self):
        if self._ptr:
            self.array = None
            _L().flmip_host_free(self._dev, self._ptr)
            self._ptr = ctypes.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class device_queue:
    def __init__(self, dev: device):
        self.dev = dev
        self._stream = ctypes.c_void_p()
        _check(_L().flmip_stream_create(dev.index, ctypes.byref(self._stream)))
        self._prof = None

    def get_queue_ptr(self) -> int:
        return self._stream.value

    def get_device(self) -> device:
        return self.dev

    def finish(self):
        _check(_L().flmip_stream_sync(self.dev.index, self._stream))

    def flush(self):
        pass

    def set_mip_chain_overlap(self, enable: bool = True):
        """chains on independent images overlap on this queue (flmip_stream_set_chain_overlap; no counterpart in the reference)"""
        _check(_L().flmip_stream_set_chain_overlap(self.dev.index, self._stream, 1 if enable else 0))

    def fence(self):
        """announces work enqueued on get_queue_ptr() behind the library's back (flmip_stream_fence)"""
        _check(_L().flmip_stream_fence(self.dev.index, self._stream))

    # -- profiling with CUDA events on this queue's stream (cuda_queue.cpp:58-70) --
    def record_event(self):
        ev = ctypes.c_void_p()
        _check(_L().flmip_event_create(self.dev.index, ctypes.byref(ev)))
        _check(_L().flmip_event_record(self.dev.index, ev, self._stream))
        return ev

    def elapsed_ms(self, start, stop, destroy: bool = True) -> float:
        ms = ctypes.c_float()
        _check(_L().flmip_event_sync(self.dev.index, stop))
        _check(_L().flmip_event_elapsed_ms(self.dev.index, start, stop, ctypes.byref(ms)))
        if destroy:
            _L().flmip_event_destroy(self.dev.index, start)
            _L().flmip_event_destroy(self.dev.index, stop)
        return ms.value

    def start_profiling(self):
        self._prof = self.record_event()

    def stop_profiling(self) -> int:
        """returns microseconds, like the reference"""
        stop = self.record_event()
        ms = self.elapsed_ms(self._prof, stop)
        self._prof = None
        return int(ms * 1000.0)

    def destroy(self):
        if self._stream:
            _L().flmip_stream_destroy(self.dev.index, self._stream)
            self._stream = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class device_image:
    def __init__(self, cqueue: device_queue, image_dim, image_type: int, data=None,
                 flags: int = MEMORY_FLAG.HOST_READ_WRITE, mip_level_limit: int = 0, no_double: bool = False,
                 force_generic: bool = False, units: bool | None = None, force_tiled: bool = False, no_tma_tiles: bool = False, tma_tiles: str | None = None):
        from . import IMAGE_NO_DOUBLE, IMAGE_FORCE_GENERIC, IMAGE_UNITS_ALWAYS, IMAGE_UNITS_NEVER, IMAGE_FORCE_TILED, IMAGE_NO_TMA_TILES, LevelInfo
        self.dev = cqueue.dev
        self._create_kw = {"no_double": no_double, "force_generic": force_generic, "units": units, "force_tiled": force_tiled, "no_tma_tiles": no_tma_tiles, "tma_tiles": tma_tiles}
        dim = list(image_dim) + [0] * (4 - len(image_dim))
        # device_image::handle_image_type (device_image.hpp:70-91)
        if flags & MEMORY_FLAG.GENERATE_MIP_MAPS:
            image_type |= IMAGE_TYPE.WRITE
        self.image_dim = tuple(dim)
        self._handle = ctypes.c_void_p()
        cflags = (IMAGE_NO_DOUBLE if no_double else 0) | (IMAGE_FORCE_GENERIC if force_generic else 0)
        cflags |= 0 if units is None else (IMAGE_UNITS_ALWAYS if units else IMAGE_UNITS_NEVER)
        cflags |= IMAGE_FORCE_TILED if force_tiled else 0
        cflags |= IMAGE_NO_TMA_TILES if no_tma_tiles else 0
        # tuning: "always" (whatever the tile count), "+split" / "+nosplit" (two levels per launch: always / never)
        if tma_tiles:
            cflags |= (64 if "always" in tma_tiles else 0) | (128 if "+split" in tma_tiles else 0) | (256 if "+nosplit" in tma_tiles else 0)
        _check(_L().flmip_image_create(self.dev.index, image_type, _u32x(dim, 4), mip_level_limit, cflags,
                                       ctypes.byref(self._handle)))
        n = ctypes.c_uint32()
        _check(_L().flmip_image_mip_level_count(self._handle, ctypes.byref(n)))
        self.mip_level_count = n.value
        if self.mip_level_count <= 1:
            image_type &= ~IMAGE_TYPE.FLAG_MIPMAPPED
        self.image_type = image_type
        self.is_mip_mapped = bool(image_type & IMAGE_TYPE.FLAG_MIPMAPPED)
        self.generate_mip_maps = self.is_mip_mapped and bool(flags & MEMORY_FLAG.GENERATE_MIP_MAPS)
        _check(_L().flmip_image_layer_count(self._handle, ctypes.byref(n)))
        self.layer_count = n.value
        self.flags = flags
        self.levels = []
        for lv in range(self.mip_level_count):
            li = LevelInfo()
            _check(_L().flmip_image_get_level_info(self._handle, lv, ctypes.byref(li)))
            self.levels.append({"dim": tuple(li.dim), "offset": li.offset, "size": li.size, "slice_size": li.slice_size})
        self.image_data_size_mip_maps = sum(l["size"] for l in self.levels)
        # with GENERATE_MIP_MAPS the user-visible size covers level 0 only (device_image.hpp:486)
        self.image_data_size = self.levels[0]["size"] if self.generate_mip_maps else self.image_data_size_mip_maps
        self._mappings = {}
        if data is not None and not (flags & MEMORY_FLAG.NO_INITIAL_COPY):
            host = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
            if host.size < self.image_data_size:
                # device_image.hpp:535-539 throws
                raise RuntimeError(f"image host data size {host.size} is smaller than the expected image size {self.image_data_size}")
            last = 0 if self.generate_mip_maps else self.mip_level_count - 1
            _check(_L().flmip_image_upload(self._handle, host.ctypes.data, host.size, 0, last, cqueue._stream))
            cqueue.finish()
            if self.generate_mip_maps:
                self.generate_mip_map_chain(cqueue)

    # ---- getters (device_image.hpp:244-326) ----
    def get_image_type(self): return self.image_type
    def get_image_dim(self): return self.image_dim
    def get_layer_count(self): return self.layer_count
    def get_mip_level_count(self): return self.mip_level_count
    def get_image_data_size(self): return self.image_data_size
    def get_generate_mip_maps(self): return self.generate_mip_maps
    def get_bytes_per_pixel(self): return it.bytes_per_pixel(self.image_type)
    def get_image_data_size_at_mip_level(self, level): return self.levels[level]["size"] if level < self.mip_level_count else 0

    def device_ptr(self) -> int:
        p = ctypes.c_uint64()
        _check(_L().flmip_image_device_ptr(self._handle, ctypes.byref(p)))
        return p.value

    def plan(self):
        a, b, c = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        _check(_L().flmip_image_plan(self._handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        d = ctypes.c_uint32()
        _check(_L().flmip_image_plan_tma_tile_launches(self._handle, ctypes.byref(d)))
        return {"single_pass": bool(a.value), "fast_levels": b.value, "launches": c.value, "tma_tile_launches": d.value}

    # ---- THE hot path ----
    def generate_mip_map_chain(self, cqueue: device_queue):
        """blocking, like the reference (device_image.cpp:235-328: wait_until_completion = true)"""
        _check(_L().flmip_mip_chain_generate(self._handle, cqueue._stream))
        cqueue.finish()

    def enqueue_mip_map_chain(self, cqueue: device_queue, first_level: int = 0):
        """non-blocking variant (no equivalent in the reference)"""
        _check(_L().flmip_mip_chain_generate_from(self._handle, first_level, cqueue._stream))

    # ---- data movement ----
    def write(self, cqueue: device_queue, src, offset, extent, mip_level_range, layer_range) -> bool:
        """cuda_image::write (cuda_image.cpp:588-673): returns False on invalid arguments instead of raising"""
        from . import FlmipError
        if src is None:
            return False
        host = np.ascontiguousarray(src).view(np.uint8).reshape(-1)
        try:
            _check(_L().flmip_image_write(self._handle, host.ctypes.data, host.size, _u32x(offset, 3), _u32x(extent, 3),
                                          _u32x(mip_level_range, 2), _u32x(layer_range, 2), cqueue._stream))
        except FlmipError:
            return False
        cqueue.finish()
        if self.generate_mip_maps:
            self.generate_mip_map_chain(cqueue)
        return True

    def zero(self, cqueue: device_queue) -> bool:
        _check(_L().flmip_image_zero(self._handle, cqueue._stream))
        cqueue.finish()
        return True

    def map(self, cqueue: device_queue, write_only: bool = False) -> np.ndarray:
        """cuda_image::map (cuda_image.cpp:703-769): host copy of image_data_size bytes (level 0 only when generating)"""
        buf = np.empty(self.image_data_size, dtype=np.uint8)
        if not write_only:
            cqueue.finish()
            last = 0 if self.generate_mip_maps else self.mip_level_count - 1
            _check(_L().flmip_image_download(self._handle, buf.ctypes.data, buf.size, 0, last, cqueue._stream))
            cqueue.finish()
        self._mappings[buf.ctypes.data] = buf
        return buf

    def unmap(self, cqueue: device_queue, mapped: np.ndarray, discard: bool = False) -> bool:
        """cuda_image::unmap (cuda_image.cpp:771-813): copy back, then regenerate the chain if generate_mip_maps"""
        if mapped is None or mapped.ctypes.data not in self._mappings:
            return False
        if not discard:
            last = 0 if self.generate_mip_maps else self.mip_level_count - 1
            _check(_L().flmip_image_upload(self._handle, mapped.ctypes.data, mapped.size, 0, last, cqueue._stream))
            cqueue.finish()
            if self.generate_mip_maps:
                self.generate_mip_map_chain(cqueue)
        del self._mappings[mapped.ctypes.data]
        return True

    def blit(self, cqueue: device_queue, src: "device_image") -> bool:
        """device_image::blit (device_image.hpp:96-101) with the checks of blit_check (device_image.cpp:470-501); blocking.
        The reference's CUDA image inherits the `return false` stub; on linear images it is one device-to-device copy."""
        from . import FlmipError
        if src.image_data_size != self.image_data_size:
            return False  # "blit: size mismatch"
        try:
            _check(_L().flmip_image_blit(self._handle, src._handle, cqueue._stream))
        except FlmipError:
            return False
        cqueue.finish()
        return True

    def blit_async(self, cqueue: device_queue, src: "device_image") -> bool:
        from . import FlmipError
        try:
            _check(_L().flmip_image_blit(self._handle, src._handle, cqueue._stream))
        except FlmipError:
            return False
        return True

    def clone(self, cqueue: device_queue, copy_contents: bool = False, flags_override: int = 0, image_type_override: int = 0) -> "device_image":
        """device_image::clone (device_image.cpp:446-466): same dim, type (unless overridden), flags and level count; host data
        is never copied into the clone; `copy_contents` blits"""
        flags = (flags_override if flags_override else self.flags) | MEMORY_FLAG.NO_INITIAL_COPY
        t = image_type_override if image_type_override else (self.image_type | (IMAGE_TYPE.FLAG_MIPMAPPED if self.mip_level_count > 1 else 0))
        ret = device_image(cqueue, self.image_dim, t, None, flags, self.mip_level_count, **self._create_kw)
        if copy_contents:
            ret.blit(cqueue, self)
        return ret

    # ---- interop with floor's tiled CUDA images (CUmipmappedArray, cuda_image.cpp:158-539) ----
    def create_tiled_twin(self) -> ctypes.c_void_p:
        """a CUmipmappedArray with the descriptor the reference's cuda_image would create for this image"""
        arr = ctypes.c_void_p()
        _check(_L().flmip_image_create_tiled_twin(self._handle, ctypes.byref(arr)))
        return arr

    def destroy_tiled(self, arr) -> None:
        _check(_L().flmip_tiled_destroy(self.dev.index, arr))

    def copy_to_tiled(self, cqueue: device_queue, arr, level_first: int = 0, level_last: int | None = None, sync: bool = True):
        last = self.mip_level_count - 1 if level_last is None else level_last
        _check(_L().flmip_image_copy_to_tiled(self._handle, arr, level_first, last, cqueue._stream))
        if sync:
            cqueue.finish()

    def copy_from_tiled(self, cqueue: device_queue, arr, level_first: int = 0, level_last: int | None = None, sync: bool = True):
        last = self.mip_level_count - 1 if level_last is None else level_last
        _check(_L().flmip_image_copy_from_tiled(self._handle, arr, level_first, last, cqueue._stream))
        if sync:
            cqueue.finish()

    def tiled_download(self, cqueue: device_queue, arr, level_first: int = 0, level_last: int | None = None) -> np.ndarray:
        """reads a CUmipmappedArray of this image's geometry back in floor's host layout"""
        last = self.mip_level_count - 1 if level_last is None else level_last
        n = self.levels[last]["offset"] + self.levels[last]["size"] - self.levels[level_first]["offset"]
        out = np.empty(n, dtype=np.uint8)
        _check(_L().flmip_tiled_download(self._handle, arr, out.ctypes.data, n, level_first, last, cqueue._stream))
        cqueue.finish()
        return out

    # ---- helpers beyond the reference API (tests / bench) ----
    def upload_levels(self, cqueue: device_queue, src, level_first: int = 0, level_last: int = 0, sync: bool = True, nbytes: int | None = None):
        if isinstance(src, int):
            ptr, size = src, nbytes
        else:
            host = np.ascontiguousarray(src).view(np.uint8).reshape(-1)
            ptr, size = host.ctypes.data, host.size
        _check(_L().flmip_image_upload(self._handle, ptr, size, level_first, level_last, cqueue._stream))
        if sync:
            cqueue.finish()

    def download_levels(self, cqueue: device_queue, level_first: int = 0, level_last: int | None = None, out=None, sync: bool = True):
        if level_last is None:
            level_last = self.mip_level_count - 1
        n = self.levels[level_last]["offset"] + self.levels[level_last]["size"] - self.levels[level_first]["offset"]
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        ptr = out if isinstance(out, int) else out.ctypes.data
        _check(_L().flmip_image_download(self._handle, ptr, n, level_first, level_last, cqueue._stream))
        if sync:
            cqueue.finish()
        return out

    def download_layers(self, cqueue: device_queue, layer_first: int, layer_count: int = 1, level_first: int = 0, level_last: int | None = None) -> np.ndarray:
        """layers [layer_first, +layer_count) of the given levels, laid out like an image of `layer_count` layers"""
        if level_last is None:
            level_last = self.mip_level_count - 1
        n = sum(self.levels[l]["slice_size"] for l in range(level_first, level_last + 1)) * layer_count
        out = np.empty(n, dtype=np.uint8)
        _check(_L().flmip_image_download_layers(self._handle, out.ctypes.data, n, level_first, level_last, layer_first, layer_count, cqueue._stream))
        cqueue.finish()
        return out

    def fill_synthetic(self, cqueue: device_queue, config_id: int, layer_id0: int = 0):
        _check(_L().flmip_image_fill_synthetic(self._handle, config_id, layer_id0, cqueue._stream))

    def destroy(self):
        if self._handle:
            _L().flmip_image_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class mip_chain_batch:
    """Batch of independent textures: one CUDA graph with a chain of kernel nodes per image and no edges between
    images, so `generate` is one graph launch and the chains run concurrently.  No equivalent in the reference, which loops
    generate_mip_map_chain over the images (one blocking launch per image, layer and level).  Destroy before the images."""

    def __init__(self, images):
        self.images = list(images)
        self.dev = self.images[0].dev
        arr = (ctypes.c_void_p * len(self.images))(*[im._handle for im in self.images])
        self._handle = ctypes.c_void_p()
        _check(_L().flmip_batch_create(arr, len(self.images), ctypes.byref(self._handle)))
        n = ctypes.c_uint32()
        _check(_L().flmip_batch_kernel_count(self._handle, ctypes.byref(n)))
        self.kernel_count = n.value

    def enqueue(self, cqueue: device_queue):
        _check(_L().flmip_batch_generate(self._handle, cqueue._stream))

    def generate(self, cqueue: device_queue):
        """blocking, like generate_mip_map_chain"""
        self.enqueue(cqueue)
        cqueue.finish()

    def destroy(self):
        if self._handle:
            _L().flmip_batch_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class device_context:
    """cuda_context subset (include/floor/device/cuda/cuda_context.hpp:40-41): constructible without floor::init"""

    def __init__(self):
        _check(_L().flmip_init())
        self.devices = [device(i) for i in range(_L().flmip_device_count())]
        self._default_queues = {}

    def is_supported(self) -> bool:
        return len(self.devices) > 0

    def get_platform_type(self) -> str:
        return "CUDA"

    def get_devices(self):
        return list(self.devices)

    def get_device(self, index: int | str = 0) -> device:
        """device::TYPE::GPU0 + index, or "FASTEST_GPU" / "FASTEST" / "ANY"; falls back to any device like the reference"""
        if isinstance(index, str):
            return self.fastest_gpu_device if index.upper().startswith("FASTEST") else self.devices[0]
        return self.devices[index] if 0 <= index < len(self.devices) else self.devices[0]

    @property
    def fastest_gpu_device(self) -> device:
        """cuda_context.cpp:340-395: score = cores per SM (128 for sm_100) x units x clock; the first device wins ties"""
        best = self.devices[0]
        for d in self.devices[1:]:
            if 128 * d.units * d.clock > 128 * best.units * best.clock:
                best = d
        return best

    def create_queue(self, dev: device) -> device_queue:
        return device_queue(dev)

    def get_device_default_queue(self, dev: device) -> device_queue:
        if dev.index not in self._default_queues:
            self._default_queues[dev.index] = device_queue(dev)
        return self._default_queues[dev.index]

    def create_image(self, cqueue: device_queue, image_dim, image_type: int, data=None,
                     flags: int = MEMORY_FLAG.HOST_READ_WRITE, mip_level_limit: int = 0, **kw) -> device_image:
        return device_image(cqueue, image_dim, image_type, data, flags, mip_level_limit, **kw)

    def create_mip_chain_batch(self, images) -> mip_chain_batch:
        return mip_chain_batch(images)
