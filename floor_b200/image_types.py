"""fl::IMAGE_TYPE / fl::MEMORY_FLAG values, bit-for-bit (reference:
include/floor/device/backend/image_types.hpp:24-425, include/floor/device/device_memory_flags.hpp:28-152).

Only the bit layout is mirrored -- the Python side is test / bench glue over the C-ABI library.
"""
from __future__ import annotations


class IMAGE_TYPE:
    NONE = 0
    # bits 0-5 format
    FORMAT_MASK = 0x3F
    FORMAT_2, FORMAT_4, FORMAT_8, FORMAT_16, FORMAT_24, FORMAT_32, FORMAT_64 = 2, 4, 11, 18, 20, 22, 24
    # bits 6-9 compression
    COMPRESSION_MASK = 0x3C0
    # bits 10-11 access
    READ, WRITE = 1 << 10, 2 << 10
    READ_WRITE = READ | WRITE
    # bits 12-13 data type
    DATA_TYPE_MASK = 0x3000
    INT, UINT, FLOAT = 1 << 12, 2 << 12, 3 << 12
    # bits 14-15 channels (count - 1)
    CHANNELS_MASK = 0xC000
    CHANNELS_1, CHANNELS_2, CHANNELS_3, CHANNELS_4 = 0 << 14, 1 << 14, 2 << 14, 3 << 14
    R, RG, RGB, RGBA = CHANNELS_1, CHANNELS_2, CHANNELS_3, CHANNELS_4
    # bits 16-17 dimensionality
    DIM_MASK = 0x30000
    DIM_1D, DIM_2D, DIM_3D = 1 << 16, 2 << 16, 3 << 16
    # bits 18-19 layout
    LAYOUT_RGBA, LAYOUT_BGRA, LAYOUT_ABGR, LAYOUT_ARGB = 0 << 18, 1 << 18, 2 << 18, 3 << 18
    # bits 20-31 flags
    FLAG_ARRAY = 1 << 20
    FLAG_BUFFER = 1 << 21
    FLAG_MSAA = 1 << 22
    FLAG_CUBE = 1 << 23
    FLAG_DEPTH = 1 << 24
    FLAG_STENCIL = 1 << 25
    FLAG_RENDER_TARGET = 1 << 26
    FLAG_MIPMAPPED = 1 << 27
    FLAG_FIXED_CHANNELS = 1 << 28
    FLAG_GATHER = 1 << 29
    FLAG_NORMALIZED = 1 << 30
    FLAG_SRGB = 1 << 31
    # base types
    IMAGE_1D = DIM_1D
    IMAGE_1D_ARRAY = DIM_1D | FLAG_ARRAY
    IMAGE_2D = DIM_2D
    IMAGE_2D_ARRAY = DIM_2D | FLAG_ARRAY
    IMAGE_CUBE = DIM_2D | FLAG_CUBE
    IMAGE_CUBE_ARRAY = DIM_2D | FLAG_CUBE | FLAG_ARRAY
    IMAGE_DEPTH = FLAG_DEPTH | CHANNELS_1 | IMAGE_2D
    IMAGE_DEPTH_ARRAY = FLAG_DEPTH | CHANNELS_1 | IMAGE_2D_ARRAY
    IMAGE_3D = DIM_3D
    # normalized
    R8 = CHANNELS_1 | FORMAT_8 | UINT | FLAG_NORMALIZED
    RG8 = CHANNELS_2 | FORMAT_8 | UINT | FLAG_NORMALIZED
    RGB8 = CHANNELS_3 | FORMAT_8 | UINT | FLAG_NORMALIZED
    RGBA8 = CHANNELS_4 | FORMAT_8 | UINT | FLAG_NORMALIZED
    BGRA8 = RGBA8 | LAYOUT_BGRA
    R16 = CHANNELS_1 | FORMAT_16 | UINT | FLAG_NORMALIZED
    RG16 = CHANNELS_2 | FORMAT_16 | UINT | FLAG_NORMALIZED
    RGBA16 = CHANNELS_4 | FORMAT_16 | UINT | FLAG_NORMALIZED
    R8I_NORM = CHANNELS_1 | FORMAT_8 | INT | FLAG_NORMALIZED
    RG8I_NORM = CHANNELS_2 | FORMAT_8 | INT | FLAG_NORMALIZED
    RGBA8I_NORM = CHANNELS_4 | FORMAT_8 | INT | FLAG_NORMALIZED
    R16I_NORM = CHANNELS_1 | FORMAT_16 | INT | FLAG_NORMALIZED
    RG16I_NORM = CHANNELS_2 | FORMAT_16 | INT | FLAG_NORMALIZED
    RGBA16I_NORM = CHANNELS_4 | FORMAT_16 | INT | FLAG_NORMALIZED
    # non-normalized
    R8UI = CHANNELS_1 | FORMAT_8 | UINT
    RG8UI = CHANNELS_2 | FORMAT_8 | UINT
    RGBA8UI = CHANNELS_4 | FORMAT_8 | UINT
    R8I = CHANNELS_1 | FORMAT_8 | INT
    RG8I = CHANNELS_2 | FORMAT_8 | INT
    RGBA8I = CHANNELS_4 | FORMAT_8 | INT
    R16UI = CHANNELS_1 | FORMAT_16 | UINT
    RG16UI = CHANNELS_2 | FORMAT_16 | UINT
    RGBA16UI = CHANNELS_4 | FORMAT_16 | UINT
    R16I = CHANNELS_1 | FORMAT_16 | INT
    RG16I = CHANNELS_2 | FORMAT_16 | INT
    RGBA16I = CHANNELS_4 | FORMAT_16 | INT
    R32UI = CHANNELS_1 | FORMAT_32 | UINT
    RG32UI = CHANNELS_2 | FORMAT_32 | UINT
    RGBA32UI = CHANNELS_4 | FORMAT_32 | UINT
    R32I = CHANNELS_1 | FORMAT_32 | INT
    RG32I = CHANNELS_2 | FORMAT_32 | INT
    RGBA32I = CHANNELS_4 | FORMAT_32 | INT
    R16F = CHANNELS_1 | FORMAT_16 | FLOAT
    RG16F = CHANNELS_2 | FORMAT_16 | FLOAT
    RGBA16F = CHANNELS_4 | FORMAT_16 | FLOAT
    R32F = CHANNELS_1 | FORMAT_32 | FLOAT
    RG32F = CHANNELS_2 | FORMAT_32 | FLOAT
    RGBA32F = CHANNELS_4 | FORMAT_32 | FLOAT
    RGB16F = CHANNELS_3 | FORMAT_16 | FLOAT
    RGB32F = CHANNELS_3 | FORMAT_32 | FLOAT
    D32F = IMAGE_DEPTH | FORMAT_32 | FLOAT
    # 3-channel formats (Host-Compute minifies them; the reference's CUDA backend cannot store them: cuda_image.cpp:173-180)
    RGB16 = CHANNELS_3 | FORMAT_16 | UINT | FLAG_NORMALIZED
    RGB8I_NORM = CHANNELS_3 | FORMAT_8 | INT | FLAG_NORMALIZED
    RGB16I_NORM = CHANNELS_3 | FORMAT_16 | INT | FLAG_NORMALIZED
    RGB8UI, RGB8I = CHANNELS_3 | FORMAT_8 | UINT, CHANNELS_3 | FORMAT_8 | INT
    RGB16UI, RGB16I = CHANNELS_3 | FORMAT_16 | UINT, CHANNELS_3 | FORMAT_16 | INT
    RGB32UI, RGB32I = CHANNELS_3 | FORMAT_32 | UINT, CHANNELS_3 | FORMAT_32 | INT
    # FORMAT_2 / FORMAT_4 normalized formats whose texel is a whole number of bytes (host_image.hpp:341-353, 1167-1186)
    RGBA2 = CHANNELS_4 | FORMAT_2 | UINT | FLAG_NORMALIZED
    RGBA2I_NORM = CHANNELS_4 | FORMAT_2 | INT | FLAG_NORMALIZED
    RG4 = CHANNELS_2 | FORMAT_4 | UINT | FLAG_NORMALIZED
    RG4I_NORM = CHANNELS_2 | FORMAT_4 | INT | FLAG_NORMALIZED
    RGBA4 = CHANNELS_4 | FORMAT_4 | UINT | FLAG_NORMALIZED
    RGBA4I_NORM = CHANNELS_4 | FORMAT_4 | INT | FLAG_NORMALIZED


class MEMORY_FLAG:
    NONE = 0
    READ, WRITE = 1 << 0, 1 << 1
    READ_WRITE = READ | WRITE
    HOST_READ, HOST_WRITE = 1 << 2, 1 << 3
    HOST_READ_WRITE = HOST_READ | HOST_WRITE
    NO_INITIAL_COPY = 1 << 4
    GENERATE_MIP_MAPS = 1 << 9


def dim_count(t: int) -> int:
    return (t >> 16) & 3


def channel_count(t: int) -> int:
    return ((t >> 14) & 3) + 1


def bits_per_channel(t: int) -> int:
    return {IMAGE_TYPE.FORMAT_2: 2, IMAGE_TYPE.FORMAT_4: 4, IMAGE_TYPE.FORMAT_8: 8, IMAGE_TYPE.FORMAT_16: 16, IMAGE_TYPE.FORMAT_32: 32}.get(t & IMAGE_TYPE.FORMAT_MASK, 0)


def bytes_per_pixel(t: int) -> int:
    return (bits_per_channel(t) * channel_count(t) + 7) // 8


def layer_count(dim, t: int) -> int:
    """image_layer_count (image_types.hpp:716-726): dim = (w, h, d or layers, layers-for-3D); 6 faces per cube"""
    d = list(dim) + [0] * (4 - len(dim))
    n = 1
    if t & IMAGE_TYPE.FLAG_ARRAY:
        n = d[1] if dim_count(t) == 1 else (d[2] if dim_count(t) == 2 else d[3])
    return n * 6 if t & IMAGE_TYPE.FLAG_CUBE else n


def mip_level_count(dim, t: int, mip_level_limit: int = 0) -> int:
    """image_mip_level_count (image_types.hpp:694-712) capped by mip_level_limit (device_image.hpp:483-485)"""
    if not (t & IMAGE_TYPE.FLAG_MIPMAPPED):
        return 1
    m = max(list(dim)[: dim_count(t)])
    n = max(m.bit_length(), 1)
    return min(n, mip_level_limit) if mip_level_limit > 0 else n


def level_size(dim, t: int, level: int) -> int:
    """bytes of one level over all layers; level dims are dim >> level without max(1) (image_types.hpp:751-766)"""
    texels = 1
    for d in list(dim)[: dim_count(t)]:
        texels *= d >> level
    return texels * bytes_per_pixel(t) * layer_count(dim, t)


def image_data_size(dim, t: int, mip_level_limit: int = 0) -> int:
    """image_data_size_from_types over all levels (image_types.hpp:731-769)"""
    return sum(level_size(dim, t, l) for l in range(mip_level_count(dim, t, mip_level_limit)))
