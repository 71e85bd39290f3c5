// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY: drives the REFERENCE's own Host-Compute mip_map_minify kernels.
//
// This file is ours; everything that computes a texel is the reference's, compiled from /root/reference by
// oracle/build_ref.py (which documents the few mechanical substitutions g++ needs):
//   * the kernels `libfloor_mip_map_minify_<IMAGE>_<SAMPLE>` + `fl::image_mip_map_minify`
//     (include/floor/device/backend/mip_map_minify.hpp:78-126) -- included verbatim below,
//   * `fl::image<>::read_lod_linear / write_lod` (include/floor/device/backend/image.hpp:458-577, 823-833, 1089-1214),
//   * the Host-Compute software sampler `fl::host_device_image` (include/floor/device/backend/host_image.hpp),
//   * `const_math::interpolate`, vector math, `soft_f16`, the IMAGE_TYPE size / level helpers (image_types.hpp:675-809).
// What this file restates (no arithmetic on texels):
//   * the level table of `host_image::create_internal` (src/device/host/host_image.cpp:81-108),
//   * the (layer x level) launch loop of `device_image::generate_mip_map_chain` (src/device/device_image.cpp:304-327),
//     one "work-item" = one call of the kernel with `global_id` set (the fibers of host_function.cpp are not needed for
//     a kernel without barriers),
//   * kernel selection by `minify_image_base_type` (mip_map_minify.hpp:53-69, device_image.cpp:255-287); cube images
//     have no kernel in the reference (static_assert :95-97), they run here as the 2D array of 6N layers they are
//     stored as (host_image.hpp:263-271) -- the same decision the oracle documents.
// All 17 kernels of FLOOR_MINIFY_IMAGE_TYPES are instantiated (1D, 1D-array, 2D, 2D-array, 3D x FLOAT / INT / UINT + the two
// depth kernels).
//
// Only tests/ may load the resulting oracle/_ref/libfloor_ref_minify*.so.
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cmath>
#include <type_traits>
#include <thread>
#include <vector>
#include <atomic>

#include <floor/core/essentials.hpp>
#include <floor/device/backend/host_pre.hpp>
#include <floor/core/enum_helpers.hpp>
#include <floor/math/vector_lib.hpp>
#include <floor/device/backend/host_limits.hpp>
#include <floor/device/backend/sampler.hpp>
#include <floor/device/backend/device_info.hpp>
#include <floor/device/backend/image_types.hpp>
#include <floor/device/backend/host_image.hpp>
#include <floor/device/backend/image.hpp>

// the kernel header is written in floor's device language: give it the three things the toolchain provides
#define FLOOR_DEVICE_HOST_COMPUTE_MINIFY 1
namespace fl { static thread_local uint3 global_id; }
using fl::global_id;
template <typename T> using param = const T&;
#define kernel_1d() static
#define kernel_2d() static
#define kernel_3d() static
#include <floor/device/backend/mip_map_minify.hpp>

namespace {

using fl::IMAGE_TYPE;

// layout of host_image::image_program_info (include/floor/device/host/host_image.hpp:79-98) == host_device_image
// (backend/host_image.hpp:837-839)
struct level_info_t {
	uint32_t dim_x, dim_y, dim_z, offset;
	int32_t clamp_dim_int[4];
	float clamp_dim_float[4];
	float clamp_dim_float_excl[4];
};
struct program_info_t {
	uint8_t* buffer;
	IMAGE_TYPE runtime_image_type;
	alignas(16) level_info_t level_info[fl::host_limits::max_mip_levels];
};
static_assert(sizeof(level_info_t) == 64);
static_assert(sizeof(program_info_t) == sizeof(fl::host_device_image<IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::FLOAT | IMAGE_TYPE::CHANNELS_4 | IMAGE_TYPE::READ_WRITE, false, false, false, false, false>));

fl::uint4 to_dim(const uint32_t d[4]) { return fl::uint4 { d[0], d[1], d[2], d[3] }; }

// cube (array) -> the 2D array it is stored as
void decube(fl::uint4& dim, IMAGE_TYPE& type) {
	if (fl::has_flag<IMAGE_TYPE::FLAG_CUBE>(type)) {
		const auto layers = fl::image_layer_count(dim, type);
		type = (type & ~IMAGE_TYPE::FLAG_CUBE) | IMAGE_TYPE::FLAG_ARRAY;
		dim.z = layers;
	}
}

// src/device/host/host_image.cpp:81-108
void fill_level_table(program_info_t& info, const fl::uint4& image_dim, IMAGE_TYPE image_type, uint32_t layer_count) {
	const auto dim_count = fl::image_dim_count(image_type);
	fl::uint4 mip_image_dim { image_dim.x, dim_count >= 2 ? image_dim.y : 0, dim_count >= 3 ? image_dim.z : 0, 0 };
	uint32_t level_offset = 0; // uint32_t in the reference
	for (size_t level = 0; level < fl::host_limits::max_mip_levels; ++level, mip_image_dim >>= 1) {
		auto& li = info.level_info[level];
		li.dim_x = mip_image_dim.x;
		li.dim_y = mip_image_dim.y;
		li.dim_z = mip_image_dim.z;
		const auto slice_data_size = fl::image_slice_data_size_from_types(mip_image_dim, image_type);
		const auto level_data_size = slice_data_size * layer_count;
		li.offset = level_offset;
		level_offset += uint32_t(level_data_size);
		const uint32_t d[3] { mip_image_dim.x, mip_image_dim.y, mip_image_dim.z };
		for (int i = 0; i < 3; ++i) {
			li.clamp_dim_int[i] = d[i] > 0 ? int(d[i] - 1) : 0;
			li.clamp_dim_float[i] = d[i] > 0 ? float(d[i]) : 0.0f;
			li.clamp_dim_float_excl[i] = d[i] > 0 ? std::nextafterf(float(d[i]), 0.0f) : 0.0f;
		}
		li.clamp_dim_int[3] = 0;
		li.clamp_dim_float[3] = 0.0f;
		li.clamp_dim_float_excl[3] = 0.0f;
	}
}

template <IMAGE_TYPE kernel_image_type>
using kernel_fn = void (*)(fl::image<kernel_image_type>, const fl::uint3&, const fl::float3&, const uint32_t&, const uint32_t&);

// one launch = every work-item of the (rounded-up) global range calls the kernel; the kernel's own bounds check
// (mip_map_minify.hpp:100) rejects the padding, so iterating level_size exactly is equivalent
template <IMAGE_TYPE kernel_image_type>
void launch(kernel_fn<kernel_image_type> fn, program_info_t* info, const fl::uint3& level_size, const fl::float3& inv_prev,
			uint32_t level, uint32_t layer, uint32_t dim_count, uint32_t threads) {
	fl::image<kernel_image_type> img;
	static_assert(sizeof(img) == sizeof(void*));
	std::memcpy((void*)&img, &info, sizeof(void*));
	const uint32_t nz = dim_count >= 3 ? level_size.z : 1u, ny = dim_count >= 2 ? level_size.y : 1u, nx = level_size.x;
	const uint64_t rows = uint64_t(nz) * ny;
	std::atomic<uint64_t> next { 0 };
	auto worker = [&]() {
		for (;;) {
			const uint64_t r0 = next.fetch_add(16);
			if (r0 >= rows) break;
			const uint64_t r1 = r0 + 16 < rows ? r0 + 16 : rows;
			for (uint64_t r = r0; r < r1; ++r) {
				for (uint32_t x = 0; x < nx; ++x) {
					fl::global_id = fl::uint3 { x, uint32_t(r % ny), uint32_t(r / ny) };
					fn(img, level_size, inv_prev, level, layer);
				}
			}
		}
	};
	if (threads <= 1 || rows < 64) {
		worker();
	} else {
		std::vector<std::thread> pool;
		for (uint32_t t = 0; t < threads; ++t) pool.emplace_back(worker);
		for (auto& t : pool) t.join();
	}
}

#define REF_KERNEL_TYPE(image_type, sample_type) (IMAGE_TYPE::image_type | IMAGE_TYPE::sample_type | IMAGE_TYPE::CHANNELS_4)
#define REF_CASE(image_type, sample_type) \
	if (base == (IMAGE_TYPE::image_type | IMAGE_TYPE::sample_type)) { \
		launch<REF_KERNEL_TYPE(image_type, sample_type)>(&libfloor_mip_map_minify_##image_type##_##sample_type, info, level_size, \
														  inv_prev, level, layer, dim_count, threads); \
		return true; \
	}

// depth kernels carry no CHANNELS_4 (mip_map_minify.hpp:113-115); their key keeps the channel bits (:64-67)
#define REF_CASE_DEPTH(image_type) \
	if (base == (IMAGE_TYPE::image_type | IMAGE_TYPE::FLOAT)) { \
		launch<(IMAGE_TYPE::image_type | IMAGE_TYPE::FLOAT)>(&libfloor_mip_map_minify_##image_type##_FLOAT, info, level_size, inv_prev, level, layer, \
															 dim_count, threads); \
		return true; \
	}

bool dispatch(IMAGE_TYPE base, program_info_t* info, const fl::uint3& level_size, const fl::float3& inv_prev, uint32_t level,
			  uint32_t layer, uint32_t dim_count, uint32_t threads) {
	REF_CASE(IMAGE_1D, FLOAT)
	REF_CASE(IMAGE_1D, INT)
	REF_CASE(IMAGE_1D, UINT)
	REF_CASE(IMAGE_1D_ARRAY, FLOAT)
	REF_CASE(IMAGE_1D_ARRAY, INT)
	REF_CASE(IMAGE_1D_ARRAY, UINT)
	REF_CASE_DEPTH(IMAGE_DEPTH)
	REF_CASE_DEPTH(IMAGE_DEPTH_ARRAY)
	REF_CASE(IMAGE_2D, FLOAT)
	REF_CASE(IMAGE_2D, INT)
	REF_CASE(IMAGE_2D, UINT)
	REF_CASE(IMAGE_2D_ARRAY, FLOAT)
	REF_CASE(IMAGE_2D_ARRAY, INT)
	REF_CASE(IMAGE_2D_ARRAY, UINT)
	REF_CASE(IMAGE_3D, FLOAT)
	REF_CASE(IMAGE_3D, INT)
	REF_CASE(IMAGE_3D, UINT)
	return false;
}

} // namespace

extern "C" {

// 0 = double scale for 9-16-bit normalized encoders (in-library Host-Compute), 1 = FLOOR_DEVICE_NO_DOUBLE build
int flr_no_double(void) {
#if defined(FLOOR_REF_NO_DOUBLE) // see build_ref.py: selects host_image.hpp:401 (`using fp_scale_type = float`)
	return 1;
#else
	return 0;
#endif
}

uint32_t flr_bytes_per_pixel(uint64_t type) { return fl::image_bytes_per_pixel(IMAGE_TYPE(type)); }
uint32_t flr_mip_level_count(const uint32_t dim[4], uint64_t type) { return fl::image_mip_level_count(to_dim(dim), IMAGE_TYPE(type)); }
uint32_t flr_layer_count(const uint32_t dim[4], uint64_t type) { return fl::image_layer_count(to_dim(dim), IMAGE_TYPE(type)); }
uint64_t flr_slice_data_size(const uint32_t dim[4], uint64_t type) { return fl::image_slice_data_size_from_types(to_dim(dim), IMAGE_TYPE(type)); }
// device_image ctor, include/floor/device/device_image.hpp:484-485
uint32_t flr_effective_level_count(const uint32_t dim[4], uint64_t type, uint32_t mip_level_limit) {
	const auto t = IMAGE_TYPE(type);
	return fl::has_flag<IMAGE_TYPE::FLAG_MIPMAPPED>(t) ?
		std::min(uint32_t(fl::image_mip_level_count(to_dim(dim), t)), mip_level_limit > 0u ? mip_level_limit : ~0u) : 1u;
}
uint64_t flr_image_data_size(const uint32_t dim[4], uint64_t type, uint32_t mip_level_limit) {
	return fl::image_data_size_from_types(to_dim(dim), IMAGE_TYPE(type), false, flr_effective_level_count(dim, type, mip_level_limit));
}
uint64_t flr_level_offset(const uint32_t dim[4], uint64_t type, uint32_t level) {
	return fl::image_mip_level_data_offset_from_types(to_dim(dim), IMAGE_TYPE(type), level);
}
uint64_t flr_level_data_size(const uint32_t dim[4], uint64_t type, uint32_t level) {
	return fl::image_mip_level_data_size_from_types(to_dim(dim), IMAGE_TYPE(type), level);
}
uint64_t flr_minify_base_type(uint64_t type) { return uint64_t(fl::minify_image_base_type(IMAGE_TYPE(type))); }

// buf: the whole level-major image, level 0 filled.  Returns 0, -1 = no kernel for this type, -2 = image too large for the
// reference's 32-bit level offsets (host_image.hpp:44).
int flr_generate_mip_map_chain(void* buf, const uint32_t dim_[4], uint64_t type_, uint32_t mip_level_limit, uint32_t threads) {
	auto image_dim = to_dim(dim_);
	auto image_type = IMAGE_TYPE(type_);
	const uint32_t mip_level_count = flr_effective_level_count(dim_, type_, mip_level_limit);
	decube(image_dim, image_type);
	const uint32_t layer_count = fl::image_layer_count(image_dim, image_type);
	if (fl::image_data_size_from_types(image_dim, image_type, false) > 0xFFFFFFFFull) return -2;

	alignas(64) program_info_t info {};
	info.buffer = (uint8_t*)buf;
	info.runtime_image_type = image_type;
	fill_level_table(info, image_dim, image_type, layer_count);

	const auto base = fl::minify_image_base_type(image_type);
	// device_image.cpp:290-327
	const auto dim_count = fl::image_dim_count(image_type);
	for (uint32_t layer = 0; layer < layer_count; ++layer) {
		fl::uint3 level_size { image_dim.x, dim_count >= 2 ? image_dim.y : 0u, dim_count >= 3 ? image_dim.z : 0u };
		fl::float3 inv_prev_level_size;
		for (uint32_t level = 0; level < mip_level_count; ++level, inv_prev_level_size = 1.0f / fl::float3(level_size), level_size >>= 1) {
			if (level == 0) continue;
			if (!dispatch(base, &info, level_size, inv_prev_level_size, level, layer, dim_count, threads)) return -1;
		}
	}
	return 0;
}

} // extern "C"
