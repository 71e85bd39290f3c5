// fl::IMAGE_TYPE and the image size / level / layer arithmetic of the mip-chain path.
//
// Drop-in for include/floor/device/backend/image_types.hpp of a2flo/floor: the enum's bit layout (:24-236) and
// the value of every alias an application can pass to create_image (:240-365) must be bit-identical, so the
// enumerator names and values are the interface.  The aliases are generated from a format table instead of
// being spelled out, and the helpers are written for this path only (uncompressed formats).
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>

namespace fl {

struct uint2 { uint32_t x = 0, y = 0; };
struct uint3 { uint32_t x = 0, y = 0, z = 0; };
struct uint4 { uint32_t x = 0, y = 0, z = 0, w = 0; };
struct float3 { float x = 0.f, y = 0.f, z = 0.f; };

// clang-format off
// (name, channels 1..4, FORMAT_*, data type, extra flags): uncompressed color formats of image_types.hpp:240-340
#define FLB_COLOR_FORMATS(F) \
	F(R8, 1, 8, UINT, FLAG_NORMALIZED) F(RG8, 2, 8, UINT, FLAG_NORMALIZED) F(RGB8, 3, 8, UINT, FLAG_NORMALIZED) F(RGBA8, 4, 8, UINT, FLAG_NORMALIZED) \
	F(BGR8, 3, 8, UINT, FLAG_NORMALIZED | LAYOUT_BGRA) F(ABGR8, 4, 8, UINT, FLAG_NORMALIZED | LAYOUT_ABGR) F(BGRA8, 4, 8, UINT, FLAG_NORMALIZED | LAYOUT_BGRA) \
	F(R16, 1, 16, UINT, FLAG_NORMALIZED) F(RG16, 2, 16, UINT, FLAG_NORMALIZED) F(RGB16, 3, 16, UINT, FLAG_NORMALIZED) F(RGBA16, 4, 16, UINT, FLAG_NORMALIZED) \
	F(R8UI_NORM, 1, 8, UINT, FLAG_NORMALIZED) F(RG8UI_NORM, 2, 8, UINT, FLAG_NORMALIZED) F(RGB8UI_NORM, 3, 8, UINT, FLAG_NORMALIZED) F(RGBA8UI_NORM, 4, 8, UINT, FLAG_NORMALIZED) \
	F(R16UI_NORM, 1, 16, UINT, FLAG_NORMALIZED) F(RG16UI_NORM, 2, 16, UINT, FLAG_NORMALIZED) F(RGB16UI_NORM, 3, 16, UINT, FLAG_NORMALIZED) F(RGBA16UI_NORM, 4, 16, UINT, FLAG_NORMALIZED) \
	F(R8I_NORM, 1, 8, INT, FLAG_NORMALIZED) F(RG8I_NORM, 2, 8, INT, FLAG_NORMALIZED) F(RGB8I_NORM, 3, 8, INT, FLAG_NORMALIZED) F(RGBA8I_NORM, 4, 8, INT, FLAG_NORMALIZED) \
	F(R16I_NORM, 1, 16, INT, FLAG_NORMALIZED) F(RG16I_NORM, 2, 16, INT, FLAG_NORMALIZED) F(RGB16I_NORM, 3, 16, INT, FLAG_NORMALIZED) F(RGBA16I_NORM, 4, 16, INT, FLAG_NORMALIZED) \
	F(R8UI, 1, 8, UINT, NONE) F(RG8UI, 2, 8, UINT, NONE) F(RGB8UI, 3, 8, UINT, NONE) F(RGBA8UI, 4, 8, UINT, NONE) \
	F(R8I, 1, 8, INT, NONE) F(RG8I, 2, 8, INT, NONE) F(RGB8I, 3, 8, INT, NONE) F(RGBA8I, 4, 8, INT, NONE) \
	F(R16UI, 1, 16, UINT, NONE) F(RG16UI, 2, 16, UINT, NONE) F(RGB16UI, 3, 16, UINT, NONE) F(RGBA16UI, 4, 16, UINT, NONE) \
	F(R16I, 1, 16, INT, NONE) F(RG16I, 2, 16, INT, NONE) F(RGB16I, 3, 16, INT, NONE) F(RGBA16I, 4, 16, INT, NONE) \
	F(R32UI, 1, 32, UINT, NONE) F(RG32UI, 2, 32, UINT, NONE) F(RGB32UI, 3, 32, UINT, NONE) F(RGBA32UI, 4, 32, UINT, NONE) \
	F(R32I, 1, 32, INT, NONE) F(RG32I, 2, 32, INT, NONE) F(RGB32I, 3, 32, INT, NONE) F(RGBA32I, 4, 32, INT, NONE) \
	F(R16F, 1, 16, FLOAT, NONE) F(RG16F, 2, 16, FLOAT, NONE) F(RGB16F, 3, 16, FLOAT, NONE) F(RGBA16F, 4, 16, FLOAT, NONE) \
	F(R32F, 1, 32, FLOAT, NONE) F(RG32F, 2, 32, FLOAT, NONE) F(RGB32F, 3, 32, FLOAT, NONE) F(RGBA32F, 4, 32, FLOAT, NONE)

enum class IMAGE_TYPE : uint64_t {
	NONE = 0ull,
	// bits 60-63 extended flags, 35-37 anisotropy, 32-34 sample count
	__EXT_FLAG_MASK = 0xFull << 60, FLAG_TRANSIENT = 1ull << 60, FLAG_16_BIT_SAMPLING = 1ull << 61,
	__ANISOTROPY_MASK = 7ull << 35, __ANISOTROPY_SHIFT = 35ull,
	__SAMPLE_COUNT_MASK = 7ull << 32, __SAMPLE_COUNT_SHIFT = 32ull,
	// bits 20-31 flags
	__FLAG_MASK = 0xFFFC0000ull, __FLAG_SHIFT = 20ull,
	FLAG_ARRAY = 1ull << 20, FLAG_BUFFER = 1ull << 21, FLAG_MSAA = 1ull << 22, FLAG_CUBE = 1ull << 23, FLAG_DEPTH = 1ull << 24,
	FLAG_STENCIL = 1ull << 25, FLAG_RENDER_TARGET = 1ull << 26, FLAG_MIPMAPPED = 1ull << 27, FLAG_FIXED_CHANNELS = 1ull << 28,
	FLAG_GATHER = 1ull << 29, FLAG_NORMALIZED = 1ull << 30, FLAG_SRGB = 1ull << 31,
	// bits 18-19 channel layout
	__LAYOUT_MASK = 3ull << 18, __LAYOUT_SHIFT = 18ull,
	LAYOUT_RGBA = 0ull << 18, LAYOUT_BGRA = 1ull << 18, LAYOUT_ABGR = 2ull << 18, LAYOUT_ARGB = 3ull << 18,
	// bits 16-17 dimensionality (of the underlying image data)
	__DIM_MASK = 3ull << 16, __DIM_SHIFT = 16ull, DIM_1D = 1ull << 16, DIM_2D = 2ull << 16, DIM_3D = 3ull << 16,
	// bits 14-15 channel count - 1
	__CHANNELS_MASK = 3ull << 14, __CHANNELS_SHIFT = 14ull,
	CHANNELS_1 = 0ull << 14, CHANNELS_2 = 1ull << 14, CHANNELS_3 = 2ull << 14, CHANNELS_4 = 3ull << 14,
	R = CHANNELS_1, RG = CHANNELS_2, RGB = CHANNELS_3, RGBA = CHANNELS_4,
	// bits 12-13 storage data type
	__DATA_TYPE_MASK = 3ull << 12, __DATA_TYPE_SHIFT = 12ull, INT = 1ull << 12, UINT = 2ull << 12, FLOAT = 3ull << 12,
	// bits 10-11 access
	__ACCESS_MASK = 3ull << 10, __ACCESS_SHIFT = 10ull, READ = 1ull << 10, WRITE = 2ull << 10, READ_WRITE = READ | WRITE,
	// bits 6-9 compression: only the field is needed here (compressed images are rejected by this path)
	__COMPRESSION_MASK = 0xFull << 6, __COMPRESSION_SHIFT = 6ull, UNCOMPRESSED = 0ull,
	// bits 0-5 per-channel format
	__FORMAT_MASK = 0x3Full,
	FORMAT_1 = 1, FORMAT_2 = 2, FORMAT_3_3_2 = 3, FORMAT_4 = 4, FORMAT_4_2_0 = 5, FORMAT_4_1_1 = 6, FORMAT_4_2_2 = 7, FORMAT_5_5_5 = 8,
	FORMAT_5_5_5_ALPHA_1 = 9, FORMAT_5_6_5 = 10, FORMAT_8 = 11, FORMAT_9_9_9_EXP_5 = 12, FORMAT_10 = 13, FORMAT_10_10_10_ALPHA_2 = 14,
	FORMAT_11_11_10 = 15, FORMAT_12_12_12 = 16, FORMAT_12_12_12_12 = 17, FORMAT_16 = 18, FORMAT_16_8 = 19, FORMAT_24 = 20, FORMAT_24_8 = 21,
	FORMAT_32 = 22, FORMAT_32_8 = 23, FORMAT_64 = 24, FORMAT_8_8_8_ALPHA_1 = 25, FORMAT_11 = 26, __FORMAT_MAX = FORMAT_64,
	// base types
	IMAGE_1D = DIM_1D, IMAGE_1D_ARRAY = DIM_1D | FLAG_ARRAY, IMAGE_1D_BUFFER = DIM_1D | FLAG_BUFFER,
	IMAGE_2D = DIM_2D, IMAGE_2D_ARRAY = DIM_2D | FLAG_ARRAY, IMAGE_2D_MSAA = DIM_2D | FLAG_MSAA, IMAGE_2D_MSAA_ARRAY = DIM_2D | FLAG_MSAA | FLAG_ARRAY,
	IMAGE_CUBE = DIM_2D | FLAG_CUBE, IMAGE_CUBE_ARRAY = DIM_2D | FLAG_CUBE | FLAG_ARRAY,
	IMAGE_DEPTH = FLAG_DEPTH | CHANNELS_1 | IMAGE_2D, IMAGE_DEPTH_STENCIL = FLAG_DEPTH | CHANNELS_2 | IMAGE_2D | FLAG_STENCIL,
	IMAGE_DEPTH_ARRAY = FLAG_DEPTH | CHANNELS_1 | IMAGE_2D_ARRAY, IMAGE_DEPTH_CUBE = FLAG_DEPTH | CHANNELS_1 | IMAGE_CUBE,
	IMAGE_DEPTH_CUBE_ARRAY = FLAG_DEPTH | CHANNELS_1 | IMAGE_CUBE | FLAG_ARRAY,
	IMAGE_3D = DIM_3D,
#define FLB_ALIAS(name, ch, bits, dtype, extra) name = CHANNELS_##ch | FORMAT_##bits | dtype | (extra),
	FLB_COLOR_FORMATS(FLB_ALIAS)
#undef FLB_ALIAS
	D16 = IMAGE_DEPTH | FORMAT_16 | UINT, D24 = IMAGE_DEPTH | FORMAT_24 | UINT, D32 = IMAGE_DEPTH | FORMAT_32 | UINT,
	D32F = IMAGE_DEPTH | FORMAT_32 | FLOAT, DS24_8 = IMAGE_DEPTH_STENCIL | FORMAT_24_8 | UINT, DS32F_8 = IMAGE_DEPTH_STENCIL | FORMAT_32_8 | FLOAT,
};
// clang-format on

#define FLB_ENUM_OPS(E)                                                                                                         \
	constexpr E operator|(E a, E b) { return E(std::underlying_type_t<E>(a) | std::underlying_type_t<E>(b)); }                  \
	constexpr E operator&(E a, E b) { return E(std::underlying_type_t<E>(a) & std::underlying_type_t<E>(b)); }                  \
	constexpr E operator^(E a, E b) { return E(std::underlying_type_t<E>(a) ^ std::underlying_type_t<E>(b)); }                  \
	constexpr E operator~(E a) { return E(~std::underlying_type_t<E>(a)); }                                                     \
	constexpr E& operator|=(E& a, E b) { return a = a | b; }                                                                    \
	constexpr E& operator&=(E& a, E b) { return a = a & b; }                                                                    \
	template <E flag> constexpr bool has_flag(E v) { return (v & flag) == flag && std::underlying_type_t<E>(flag) != 0; }
FLB_ENUM_OPS(IMAGE_TYPE)

constexpr uint64_t image_type_bits(IMAGE_TYPE t) { return static_cast<uint64_t>(t); }

// ---- helpers with the reference's names (image_types.hpp:449-809), uncompressed formats only ----------------
constexpr uint32_t image_dim_count(IMAGE_TYPE t) { return uint32_t((image_type_bits(t) >> 16) & 3u); }
constexpr uint32_t image_channel_count(IMAGE_TYPE t) { return uint32_t((image_type_bits(t) >> 14) & 3u) + 1u; }
constexpr bool image_compressed(IMAGE_TYPE t) { return (t & IMAGE_TYPE::__COMPRESSION_MASK) != IMAGE_TYPE::UNCOMPRESSED; }
constexpr uint32_t image_bits_of_channel_format(IMAGE_TYPE t) {
	switch (t & IMAGE_TYPE::__FORMAT_MASK) {
		case IMAGE_TYPE::FORMAT_2: return 2;
		case IMAGE_TYPE::FORMAT_4: return 4;
		case IMAGE_TYPE::FORMAT_8: return 8;
		case IMAGE_TYPE::FORMAT_16: return 16;
		case IMAGE_TYPE::FORMAT_24: return 24;
		case IMAGE_TYPE::FORMAT_32: return 32;
		case IMAGE_TYPE::FORMAT_64: return 64;
		default: return 0; // packed / special formats: not minifiable (host_image.hpp floor_unreachable())
	}
}
constexpr uint32_t image_bits_per_pixel(IMAGE_TYPE t) { return image_bits_of_channel_format(t) * image_channel_count(t); }
constexpr uint32_t image_bytes_per_pixel(IMAGE_TYPE t) { return (image_bits_per_pixel(t) + 7u) / 8u; }
constexpr bool image_format_valid(IMAGE_TYPE t) { return image_bits_of_channel_format(t) != 0 && (t & IMAGE_TYPE::__DATA_TYPE_MASK) != IMAGE_TYPE::NONE; }

// image_dim = (w, h, d or layers, layers-for-3D); layers * 6 for cubes (image_types.hpp:716-726)
constexpr uint32_t image_layer_count(uint4 dim, IMAGE_TYPE t) {
	const uint32_t dc = image_dim_count(t);
	uint32_t layers = 1;
	if (has_flag<IMAGE_TYPE::FLAG_ARRAY>(t)) layers = dc == 1 ? dim.y : (dc == 2 ? dim.z : dim.w);
	return has_flag<IMAGE_TYPE::FLAG_CUBE>(t) ? layers * 6u : layers;
}
// number of levels of a mip-mapped image: position of the highest set bit of the largest dim (:702-712)
constexpr uint32_t image_mip_level_count_from_max_dim(uint32_t max_dim) {
	uint32_t n = 0;
	for (; max_dim != 0; max_dim >>= 1) ++n;
	return n == 0 ? 1u : n;
}
constexpr uint32_t image_mip_level_count(uint4 dim, IMAGE_TYPE t) {
	if (!has_flag<IMAGE_TYPE::FLAG_MIPMAPPED>(t)) return 1;
	const uint32_t dc = image_dim_count(t);
	uint32_t m = dim.x;
	if (dc >= 2 && dim.y > m) m = dim.y;
	if (dc >= 3 && dim.z > m) m = dim.z;
	return image_mip_level_count_from_max_dim(m);
}
// bytes of one layer of one level; level dims are `dim >> level` WITHOUT max(1): a level with a zero dim is empty
// (the reference's quirk for non-square images, image_types.hpp:751-766)
constexpr size_t image_slice_data_size_from_types(uint4 dim, IMAGE_TYPE t, uint32_t level = 0) {
	const uint32_t dc = image_dim_count(t);
	size_t texels = dim.x >> level;
	if (dc >= 2) texels *= dim.y >> level;
	if (dc >= 3) texels *= dim.z >> level;
	return texels * image_bytes_per_pixel(t);
}
constexpr size_t image_mip_level_data_size_from_types(uint4 dim, IMAGE_TYPE t, uint32_t level) {
	return image_slice_data_size_from_types(dim, t, level) * image_layer_count(dim, t);
}
// total bytes: level 0 only if `ignore_mip_levels` (GENERATE_MIP_MAPS images expose level 0 only, device_image.hpp:486)
constexpr size_t image_data_size_from_types(uint4 dim, IMAGE_TYPE t, bool ignore_mip_levels = false, uint32_t mip_level_limit = 0) {
	uint32_t levels = ignore_mip_levels ? 1u : image_mip_level_count(dim, t);
	if (mip_level_limit != 0 && mip_level_limit < levels) levels = mip_level_limit;
	size_t sum = 0;
	for (uint32_t l = 0; l < levels; ++l) sum += image_mip_level_data_size_from_types(dim, t, l);
	return sum;
}
constexpr size_t image_mip_level_data_offset_from_types(uint4 dim, IMAGE_TYPE t, uint32_t level) {
	size_t off = 0;
	for (uint32_t l = 0; l < level; ++l) off += image_mip_level_data_size_from_types(dim, t, l);
	return off;
}

// spot values of SURVEY.md section 8b (cross-checked there against the constants in the reference PTX)
static_assert(image_type_bits(IMAGE_TYPE::RGBA8) == 0x4000E00Bull && image_type_bits(IMAGE_TYPE::RGBA16) == 0x4000E012ull);
static_assert(image_type_bits(IMAGE_TYPE::RGBA8UI) == 0xE00Bull && image_type_bits(IMAGE_TYPE::RGBA32UI) == 0xE016ull);
static_assert(image_type_bits(IMAGE_TYPE::RGBA32I) == 0xD016ull && image_type_bits(IMAGE_TYPE::RGBA8I_NORM) == 0x4000D00Bull);
static_assert(image_type_bits(IMAGE_TYPE::RGBA16F) == 0xF012ull && image_type_bits(IMAGE_TYPE::RGBA32F) == 0xF016ull);
static_assert(image_type_bits(IMAGE_TYPE::R32F) == 0x3016ull && image_type_bits(IMAGE_TYPE::IMAGE_2D_ARRAY) == 0x120000ull);
static_assert(image_type_bits(IMAGE_TYPE::IMAGE_CUBE_ARRAY) == 0x920000ull && image_type_bits(IMAGE_TYPE::IMAGE_3D) == 0x30000ull);
static_assert(image_type_bits(IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGBA16F | IMAGE_TYPE::FLAG_MIPMAPPED | IMAGE_TYPE::READ_WRITE) == 0x802FC12ull);
static_assert(image_mip_level_count({ 8192, 8192, 0, 0 }, IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGBA16F | IMAGE_TYPE::FLAG_MIPMAPPED) == 14);
static_assert(image_data_size_from_types({ 8192, 8192, 0, 0 }, IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGBA16F | IMAGE_TYPE::FLAG_MIPMAPPED) == 715827880ull);
static_assert(image_data_size_from_types({ 512, 512, 512, 0 }, IMAGE_TYPE::IMAGE_3D | IMAGE_TYPE::R32F | IMAGE_TYPE::FLAG_MIPMAPPED) == 613566756ull);

} // namespace fl
