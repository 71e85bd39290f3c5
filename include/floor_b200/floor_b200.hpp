// C++ drop-in for the mip-chain path of a2flo/floor (libfloor), over the C-ABI of libfloor_b200_mip.so.
//
// Mirrors, for THIS path only, the classes an application touches (file:line relative to a2flo/floor):
//   fl::device / fl::cuda_device ........ include/floor/device/device.hpp:36-230, include/floor/device/cuda/cuda_device.hpp:30-86
//   fl::device_queue / fl::cuda_queue ... include/floor/device/device_queue.hpp:95-108, src/device/cuda/cuda_queue.cpp:26-72
//   fl::device_image / fl::cuda_image ... include/floor/device/device_image.hpp:44-91, 101-162, 244-340, 470-540;
//                                         src/device/cuda/cuda_image.cpp:158-539 (create), 588-673 (write), 703-813 (map/unmap)
//   fl::device_context / fl::cuda_context include/floor/device/device_context.hpp:106-116, 261-331; include/floor/device/cuda/cuda_context.hpp:40-41
// Same names, argument meaning and error behaviour: constructors throw std::runtime_error on invalid type / flag
// combinations, everything else logs and returns false / nullptr; generate_mip_map_chain() is void and blocking.
// Header-only; link with -lfloor_b200_mip.  There is no CPU fallback: without a usable CUDA device
// cuda_context::is_supported() is false and create_image() returns nullptr.
#pragma once
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <span>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../floor_b200_mip.h"
#include "device_memory_flags.hpp"
#include "image_types.hpp"

namespace fl {

#define FLB_LOG_ERROR(...)                     \
	do {                                       \
		std::fprintf(stderr, "[floor_b200] "); \
		std::fprintf(stderr, __VA_ARGS__);     \
		std::fprintf(stderr, "\n");            \
	} while (0)

enum class PLATFORM_TYPE : uint64_t { NONE = 0u, OPENCL = 1u, CUDA = 2u, METAL = 3u, HOST = 4u, VULKAN = 5u };
enum class DEVICE_CONTEXT_FLAGS : uint32_t { NONE = 0u };

class device_context;
class device_queue;
class device_image;

//! What device_image::generate_mip_map_chain dispatches to.  The reference keeps one `minify_program` per device_context
//! (device_image.hpp:377-395) that provide_minify_program() fills from a compiled device_program (device_image.cpp:155-194); here a
//! program is an object that generates levels > first_level of an image on a queue (enqueue only -- the caller blocks).  The
//! built-in one launches the sm_100a kernels of libfloor_b200_mip.so; a context can register its own with
//! device_image::provide_minify_program (e.g. to trace, or to route some image types elsewhere) -- no FUBAR, no toolchain.
struct minify_program {
	virtual ~minify_program() = default;
	virtual bool minify(device_image& img, const device_queue& cqueue, uint32_t first_level) = 0;
};

//! fl::device: plain struct of public fields; the ones this path reads or a caller inspects
struct device {
	enum class TYPE : uint32_t {
		GPU = 1u << 31, CPU = 1u << 30, FASTEST_FLAG = 1u << 29,
		NONE = 0u, ANY = 1u, FASTEST = ANY | FASTEST_FLAG, FASTEST_GPU = GPU | FASTEST_FLAG, FASTEST_CPU = CPU | FASTEST_FLAG,
		ALL_GPU = GPU | (FASTEST_FLAG - 1u), ALL_CPU = CPU | (FASTEST_FLAG - 1u), ALL_DEVICES = GPU | CPU | (FASTEST_FLAG - 1u),
		GPU0 = GPU, GPU1, GPU2, GPU3, GPU4, GPU5, GPU6, GPU7, GPU255 = GPU0 + 255u,
		CPU0 = CPU, CPU255 = CPU0 + 255u,
	};
	TYPE type { TYPE::NONE };
	std::string name;
	uint32_t units { 0 };
	uint32_t clock { 0 }, mem_clock { 0 }, mem_bus_width { 0 }; //!< MHz, MHz, bits (device.hpp:101-109)
	uint64_t global_mem_size { 0 };
	uint32_t max_total_local_size { 0 };
	uint2 max_image_2d_dim;
	uint3 max_image_3d_dim;
	uint32_t max_mip_levels { 0 };
	bool image_support { true }, image_mipmap_support { true }, image_mipmap_write_support { true };
	bool image_depth_support { true }, image_depth_write_support { true };
	//! linear images lift the reference CUDA device's "no cube write" limit (src/device/cuda/cuda_device.cpp:54-57)
	bool image_cube_write_support { true }, image_cube_array_write_support { true };
	device_context* context { nullptr };
	bool is_gpu() const { return (uint32_t(type) & uint32_t(TYPE::GPU)) != 0; }
};

struct cuda_device : device {
	uint2 sm { 10, 0 };
	bool sm_aa { true };
	int32_t device_id { 0 }; //!< index for the C-ABI
	uint32_t driver_version { 0 };
	bool make_context_current() const { return true; } //!< the C-ABI makes the context current per call
};

//! device_queue: one CUstream (non-blocking, like cuda_context::create_queue, cuda_context.cpp:418-437)
class device_queue {
public:
	explicit device_queue(const cuda_device& dev_) : dev(dev_) {
		if (flmip_stream_create(dev.device_id, &stream) != FLMIP_OK) {
			FLB_LOG_ERROR("failed to create a queue: %s", flmip_last_error_string());
			stream = nullptr;
		}
	}
	~device_queue() {
		if (prof_start) flmip_event_destroy(dev.device_id, prof_start);
		if (stream) flmip_stream_destroy(dev.device_id, stream);
	}
	device_queue(const device_queue&) = delete;
	device_queue& operator=(const device_queue&) = delete;

	//! blocks until all currently scheduled work in this queue has been executed
	void finish() const {
		if (flmip_stream_sync(dev.device_id, stream) != FLMIP_OK) FLB_LOG_ERROR("queue finish failed: %s", flmip_last_error_string());
	}
	void flush() const {}
	//! not in the reference: chains on independent images enqueued on this queue overlap (flmip_stream_set_chain_overlap); work enqueued on
	//! get_queue_ptr() behind the library's back must then be announced with fence() before the next chain
	void set_mip_chain_overlap(const bool enable) const {
		if (flmip_stream_set_chain_overlap(dev.device_id, stream, enable ? 1 : 0) != FLMIP_OK) FLB_LOG_ERROR("set_mip_chain_overlap failed: %s", flmip_last_error_string());
	}
	void fence() const { flmip_stream_fence(dev.device_id, stream); }
	const void* get_queue_ptr() const { return stream; }
	void* get_queue_ptr() { return stream; }
	const cuda_device& get_device() const { return dev; }

	//! profiling with CUDA events on this queue (cuda_queue.cpp:58-70); stop returns microseconds
	void start_profiling() const {
		if (!prof_start) flmip_event_create(dev.device_id, &prof_start);
		flmip_event_record(dev.device_id, prof_start, stream);
	}
	uint64_t stop_profiling() const {
		flmip_event stop = nullptr;
		float ms = 0.f;
		flmip_event_create(dev.device_id, &stop);
		flmip_event_record(dev.device_id, stop, stream);
		flmip_event_sync(dev.device_id, stop);
		flmip_event_elapsed_ms(dev.device_id, prof_start, stop, &ms);
		flmip_event_destroy(dev.device_id, stop);
		return uint64_t(double(ms) * 1000.0);
	}

private:
	const cuda_device& dev;
	flmip_stream stream { nullptr };
	mutable flmip_event prof_start { nullptr };
};
using cuda_queue = device_queue;

//! device_image: linear device image in floor's host layout + the mip-chain life cycle
class device_image {
public:
	//! device_image.hpp:44-57: GENERATE_MIP_MAPS implies WRITE access
	static constexpr MEMORY_FLAG infer_rw_flags(IMAGE_TYPE type, MEMORY_FLAG flags) {
		if (has_flag<IMAGE_TYPE::READ>(type)) flags |= MEMORY_FLAG::READ;
		if (has_flag<IMAGE_TYPE::WRITE>(type)) flags |= MEMORY_FLAG::WRITE;
		if ((flags & MEMORY_FLAG::READ_WRITE) == MEMORY_FLAG::NONE) flags |= MEMORY_FLAG::READ_WRITE;
		if (has_flag<MEMORY_FLAG::GENERATE_MIP_MAPS>(flags)) flags |= MEMORY_FLAG::WRITE;
		return flags;
	}
	//! device_image.hpp:74-91: access bits follow the memory flags; a "mip-mapped" image with a single level is not mip-mapped
	static constexpr IMAGE_TYPE handle_image_type(uint4 dim, IMAGE_TYPE type, MEMORY_FLAG flags) {
		if (has_flag<MEMORY_FLAG::READ>(flags)) type |= IMAGE_TYPE::READ;
		if (has_flag<MEMORY_FLAG::WRITE>(flags) || has_flag<MEMORY_FLAG::GENERATE_MIP_MAPS>(flags)) type |= IMAGE_TYPE::WRITE;
		if (has_flag<IMAGE_TYPE::FLAG_MIPMAPPED>(type) && image_mip_level_count(dim, type) <= 1) type &= ~IMAGE_TYPE::FLAG_MIPMAPPED;
		return type;
	}

	device_image(const device_queue& cqueue, uint4 image_dim_, IMAGE_TYPE image_type_, std::span<uint8_t> host_data_, MEMORY_FLAG flags_,
				 uint32_t mip_level_limit = 0u, const char* debug_label_ = nullptr)
		: dev(cqueue.get_device()), host_data(host_data_), flags(infer_rw_flags(image_type_, flags_)), image_dim(image_dim_),
		  image_type(handle_image_type(image_dim_, image_type_, flags)), is_mip_mapped(has_flag<IMAGE_TYPE::FLAG_MIPMAPPED>(image_type)),
		  generate_mip_maps(is_mip_mapped && has_flag<MEMORY_FLAG::GENERATE_MIP_MAPS>(flags_)),
		  mip_level_count(is_mip_mapped ? std::min(image_mip_level_count(image_dim_, image_type), mip_level_limit > 0u ? mip_level_limit : ~0u) : 1u),
		  image_data_size(image_data_size_from_types(image_dim_, image_type, generate_mip_maps, mip_level_count)),
		  layer_count(image_layer_count(image_dim_, image_type)),
		  image_data_size_mip_maps(image_data_size_from_types(image_dim_, image_type, false, mip_level_count)),
		  debug_label(debug_label_ ? debug_label_ : "") {
		// constructor invariants of device_image.hpp:502-539: these throw, everything else logs and returns
		if (has_flag<IMAGE_TYPE::FLAG_MIPMAPPED>(image_type) &&
			(has_flag<IMAGE_TYPE::FLAG_RENDER_TARGET>(image_type) || has_flag<IMAGE_TYPE::FLAG_TRANSIENT>(image_type)))
			throw std::runtime_error("image can't be both mip-mapped and a render and/or transient target!");
		if (has_flag<IMAGE_TYPE::FLAG_MIPMAPPED>(image_type) && has_flag<IMAGE_TYPE::FLAG_MSAA>(image_type))
			throw std::runtime_error("image can't be both mip-mapped and a multi-sampled image!");
		if (image_compressed(image_type) && has_flag<IMAGE_TYPE::WRITE>(image_type)) throw std::runtime_error("image can not be compressed and writable!");
		if (!image_format_valid(image_type)) throw std::runtime_error("invalid image format: " + std::to_string(image_type_bits(image_type)));
		if (image_compressed(image_type) && generate_mip_maps)
			throw std::runtime_error("generating mip-maps for compressed image data is not supported!");
		if (host_data.data() != nullptr && host_data.size_bytes() < image_data_size)
			throw std::runtime_error("image host data size " + std::to_string(host_data.size_bytes()) + " is smaller than the expected image size " +
									 std::to_string(image_data_size));
		create_internal(cqueue);
	}
	~device_image() {
		for (auto& m : mappings) std::free(m.first);
		if (handle) flmip_image_destroy(handle);
	}
	device_image(const device_image&) = delete;
	device_image& operator=(const device_image&) = delete;

	//! true if the device allocation exists (cuda_image::create_internal succeeded)
	bool is_valid() const { return handle != nullptr; }

	// ---- THE hot path: device_image.hpp:161-162, device_image.cpp:235-328 --------------------------------------
	//! generates the whole mip chain from level 0; blocks until the last level has been written
	virtual void generate_mip_map_chain(const device_queue& cqueue) {
		if (!handle) return;
		if (!generate_mip_map_chain_async(cqueue, 0u)) {
			FLB_LOG_ERROR("mip-map minification failed: %s", flmip_last_error_string()); // device_image.cpp:250-253, 280-283: log + return
			return;
		}
		cqueue.finish();
	}
	//! non-blocking variant (no equivalent in the reference): enqueue only; levels > first_level are regenerated
	bool generate_mip_map_chain_async(const device_queue& cqueue, uint32_t first_level = 0u) {
		if (!handle) return false;
		if (auto prog = lookup_minify_program(dev.context)) return prog->minify(*this, cqueue, first_level); // device_image.cpp:259-284
		return flmip_mip_chain_generate_from(handle, first_level, const_cast<void*>(cqueue.get_queue_ptr())) == FLMIP_OK;
	}

	//! device_image::provide_minify_program (device_image.hpp:171-172, device_image.cpp:155-194): registers `prog` as the minify
	//! program of `ctx`; every image of that context dispatches generate_mip_map_chain to it from now on.  nullptr restores the
	//! built-in program.  Thread-safe (the reference guards its map with minify_programs_mtx, device_image.cpp:40).
	static bool provide_minify_program(device_context& ctx, std::shared_ptr<minify_program> prog) {
		std::lock_guard<std::mutex> lock(minify_programs_mtx());
		if (prog) minify_programs()[&ctx] = std::move(prog);
		else minify_programs().erase(&ctx);
		return true;
	}
	//! device_image::destroy_minify_programs (device_image.cpp:549-552)
	static void destroy_minify_programs() {
		std::lock_guard<std::mutex> lock(minify_programs_mtx());
		minify_programs().clear();
	}
	//! the built-in program: the single-pass / tile kernels of libfloor_b200_mip.so on the image's native handle
	static std::shared_ptr<minify_program> builtin_minify_program() {
		struct builtin : minify_program {
			bool minify(device_image& img, const device_queue& cqueue, uint32_t first_level) override {
				return flmip_mip_chain_generate_from(img.get_native_handle(), first_level, const_cast<void*>(cqueue.get_queue_ptr())) == FLMIP_OK;
			}
		};
		static const auto prog = std::make_shared<builtin>();
		return prog;
	}

	// ---- host <-> device ---------------------------------------------------------------------------------------
	//! device_image.hpp:116-120 / cuda_image.cpp:588-673: inclusive level and layer ranges, tightly packed source
	virtual bool write(const device_queue& cqueue, const void* src, size_t src_size, uint3 offset, uint3 extent, uint2 mip_level_range,
					   uint2 layer_range) {
		if (!handle || !src) return false;
		// write_check (device_image.cpp:503-547)
		if (!has_flag<MEMORY_FLAG::HOST_WRITE>(flags)) {
			FLB_LOG_ERROR("write: image is not host-writable");
			return false;
		}
		if (src_size == 0) {
			FLB_LOG_ERROR("write: trying to write 0 bytes!");
			return false;
		}
		const uint32_t o[3] = { offset.x, offset.y, offset.z }, e[3] = { extent.x, extent.y, extent.z };
		const uint32_t lv[2] = { mip_level_range.x, mip_level_range.y }, ly[2] = { layer_range.x, layer_range.y };
		if (flmip_image_write(handle, src, src_size, o, e, lv, ly, const_cast<void*>(cqueue.get_queue_ptr())) != FLMIP_OK) {
			FLB_LOG_ERROR("image write failed: %s", flmip_last_error_string());
			return false;
		}
		cqueue.finish();
		if (generate_mip_maps) generate_mip_map_chain(cqueue); // cuda_image.cpp:667-670
		return true;
	}
	template <typename data_type>
	bool write(const device_queue& cqueue, std::span<data_type> src, uint3 offset, uint3 extent, uint2 mip_level_range, uint2 layer_range) {
		return write(cqueue, (const void*)src.data(), src.size_bytes(), offset, extent, mip_level_range, layer_range);
	}

	//! cuda_image.cpp:703-769: returns a host copy of image_data_size bytes (level 0 only for GENERATE_MIP_MAPS images), owned by the image
	virtual void* map(const device_queue& cqueue, MEMORY_MAP_FLAG map_flags = (MEMORY_MAP_FLAG::READ_WRITE | MEMORY_MAP_FLAG::BLOCK)) {
		if (!handle) return nullptr;
		void* ptr = nullptr;
		if (posix_memalign(&ptr, 128, image_data_size ? image_data_size : 128) != 0) return nullptr;
		const bool write_only = has_flag<MEMORY_MAP_FLAG::WRITE_INVALIDATE>(map_flags);
		if (!write_only) {
			cqueue.finish();
			if (flmip_image_download(handle, ptr, image_data_size, 0, mappable_last_level(), const_cast<void*>(cqueue.get_queue_ptr())) != FLMIP_OK) {
				FLB_LOG_ERROR("image map failed: %s", flmip_last_error_string());
				std::free(ptr);
				return nullptr;
			}
			cqueue.finish();
		}
		std::lock_guard<std::mutex> lock(mappings_mtx);
		mappings.emplace(ptr, map_flags);
		return ptr;
	}
	//! cuda_image.cpp:771-813: copies back if mapped for writing, then regenerates the chain for GENERATE_MIP_MAPS images
	virtual bool unmap(const device_queue& cqueue, void* mapped_ptr, bool discard = false) {
		if (!handle || !mapped_ptr) return false;
		MEMORY_MAP_FLAG mf;
		{
			std::lock_guard<std::mutex> lock(mappings_mtx);
			const auto it = mappings.find(mapped_ptr);
			if (it == mappings.end()) {
				FLB_LOG_ERROR("invalid mapped pointer");
				return false;
			}
			mf = it->second;
			mappings.erase(it);
		}
		bool ok = true;
		const bool wrote = has_flag<MEMORY_MAP_FLAG::WRITE>(mf) || has_flag<MEMORY_MAP_FLAG::WRITE_INVALIDATE>(mf);
		if (wrote && !discard) {
			ok = flmip_image_upload(handle, mapped_ptr, image_data_size, 0, mappable_last_level(), const_cast<void*>(cqueue.get_queue_ptr())) == FLMIP_OK;
			cqueue.finish();
			if (ok && generate_mip_maps) generate_mip_map_chain(cqueue); // cuda_image.cpp:803-806
		}
		std::free(mapped_ptr);
		return ok;
	}
	//! cuda_image.cpp:675-701
	virtual bool zero(const device_queue& cqueue) {
		if (!handle || flmip_image_zero(handle, const_cast<void*>(cqueue.get_queue_ptr())) != FLMIP_OK) return false;
		cqueue.finish();
		return true;
	}
	//! device_image::blit (device_image.hpp:96-101) with the checks of blit_check (device_image.cpp:470-501); blocking.
	//! The reference's CUDA image inherits the `return false` stub (only Vulkan / Metal implement it); on linear images
	//! it is one device-to-device copy of every level both images have.
	virtual bool blit(const device_queue& cqueue, device_image& src) {
		if (!blit_async(cqueue, src)) return false;
		cqueue.finish();
		return true;
	}
	virtual bool blit_async(const device_queue& cqueue, device_image& src) {
		if (!handle || !src.handle) return false;
		if (src.get_image_data_size() != image_data_size) {
			FLB_LOG_ERROR("blit: size mismatch: src %zu != dst %zu", src.get_image_data_size(), image_data_size);
			return false;
		}
		if (flmip_image_blit(handle, src.handle, const_cast<void*>(cqueue.get_queue_ptr())) != FLMIP_OK) {
			FLB_LOG_ERROR("%s", flmip_last_error_string());
			return false;
		}
		return true;
	}
	//! device_image::clone (device_image.cpp:446-466): same dim, type (unless overridden), flags and level count; host data is
	//! never copied into the clone; `copy_contents` blits (which actually copies here)
	std::shared_ptr<device_image> clone(const device_queue& cqueue, bool copy_contents = false, MEMORY_FLAG flags_override = MEMORY_FLAG::NONE,
										IMAGE_TYPE image_type_override = IMAGE_TYPE::NONE, const char* clone_debug_label = nullptr) const {
		auto clone_flags = (flags_override != MEMORY_FLAG::NONE ? flags_override : flags);
		if (host_data.data() != nullptr) clone_flags |= MEMORY_FLAG::NO_INITIAL_COPY;
		std::shared_ptr<device_image> ret;
		try {
			ret = std::make_shared<device_image>(cqueue, image_dim, image_type_override == IMAGE_TYPE::NONE ? image_type : image_type_override, host_data,
												 clone_flags, mip_level_count, clone_debug_label);
		} catch (const std::exception& e) {
			FLB_LOG_ERROR("clone: %s", e.what());
			return {};
		}
		if (!ret->is_valid()) return {};
		if (copy_contents) ret->blit(cqueue, const_cast<device_image&>(*this));
		return ret;
	}

	// ---- interop with floor's tiled CUDA images (CUmipmappedArray + texture / surface objects, cuda_image.cpp:158-539) ----
	//! creates a CUmipmappedArray with the descriptor the original cuda_image uses for this image (caller destroys it)
	void* create_tiled_twin() const {
		void* arr = nullptr;
		if (!handle || flmip_image_create_tiled_twin(handle, &arr) != FLMIP_OK) {
			FLB_LOG_ERROR("%s", flmip_last_error_string());
			return nullptr;
		}
		return arr;
	}
	void destroy_tiled(void* mipmapped_array) const { flmip_tiled_destroy(dev.device_id, mipmapped_array); }
	//! linear -> CUmipmappedArray, levels [first, last]: publishes generated levels to kernels that sample through texture objects
	bool copy_to_tiled(const device_queue& cqueue, void* mipmapped_array, uint32_t first_level, uint32_t last_level) const {
		if (!handle || flmip_image_copy_to_tiled(handle, mipmapped_array, first_level, last_level, const_cast<void*>(cqueue.get_queue_ptr())) != FLMIP_OK) return false;
		cqueue.finish();
		return true;
	}
	//! CUmipmappedArray -> linear, levels [first, last]: pulls level 0 of an image the original cuda_image owns
	bool copy_from_tiled(const device_queue& cqueue, void* mipmapped_array, uint32_t first_level, uint32_t last_level) {
		if (!handle || flmip_image_copy_from_tiled(handle, mipmapped_array, first_level, last_level, const_cast<void*>(cqueue.get_queue_ptr())) != FLMIP_OK) return false;
		cqueue.finish();
		return true;
	}
	//! generate_mip_map_chain for an image that lives in a CUmipmappedArray (the original cuda_image's storage): level 0 in,
	//! single-pass chain, generated levels out -- three stream-ordered steps, one host wait
	bool generate_mip_map_chain_for_tiled(const device_queue& cqueue, void* mipmapped_array) {
		void* stream = const_cast<void*>(cqueue.get_queue_ptr());
		if (!handle || flmip_image_copy_from_tiled(handle, mipmapped_array, 0, 0, stream) != FLMIP_OK) return false;
		if (flmip_mip_chain_generate(handle, stream) != FLMIP_OK) return false;
		if (mip_level_count > 1 && flmip_image_copy_to_tiled(handle, mipmapped_array, 1, mip_level_count - 1, stream) != FLMIP_OK) return false;
		cqueue.finish();
		return true;
	}
	//! reads a CUmipmappedArray of this image's geometry back in floor's host layout
	bool read_tiled_levels(const device_queue& cqueue, void* mipmapped_array, void* dst, size_t dst_size, uint32_t first_level, uint32_t last_level) const {
		if (!handle || flmip_tiled_download(handle, mipmapped_array, dst, dst_size, first_level, last_level, const_cast<void*>(cqueue.get_queue_ptr())) != FLMIP_OK) return false;
		cqueue.finish();
		return true;
	}

	//! reads back whole levels [first, last] in host layout (what a harness needs to look at generated levels; the
	//! reference can only do this through map() on an image created without GENERATE_MIP_MAPS)
	bool read_levels(const device_queue& cqueue, void* dst, size_t dst_size, uint32_t first_level, uint32_t last_level) const {
		if (!handle || flmip_image_download(handle, dst, dst_size, first_level, last_level, const_cast<void*>(cqueue.get_queue_ptr())) != FLMIP_OK) return false;
		cqueue.finish();
		return true;
	}

	// ---- getters (device_image.hpp:244-340) --------------------------------------------------------------------
	const uint4& get_image_dim() const { return image_dim; }
	IMAGE_TYPE get_image_type() const { return image_type; }
	size_t get_image_data_size() const { return image_data_size; }
	size_t get_image_data_size_at_mip_level(uint32_t level) const { return level < mip_level_count ? image_mip_level_data_size_from_types(image_dim, image_type, level) : 0; }
	bool get_generate_mip_maps() const { return generate_mip_maps; }
	uint32_t get_dim_count() const { return image_dim_count(image_type); }
	uint32_t get_channel_count() const { return image_channel_count(image_type); }
	uint32_t get_bits_per_pixel() const { return image_bits_per_pixel(image_type); }
	uint32_t get_bytes_per_pixel() const { return image_bytes_per_pixel(image_type); }
	uint32_t get_mip_level_count() const { return mip_level_count; }
	uint32_t get_layer_count() const { return layer_count; }
	size_t get_slice_data_size() const { return image_slice_data_size_from_types(image_dim, image_type); }
	MEMORY_FLAG get_flags() const { return flags; }
	const std::string& get_debug_label() const { return debug_label; }
	const device& get_device() const { return dev; }
	//! CUdeviceptr of level 0 (linear memory: usable by any CUDA kernel)
	uint64_t get_device_ptr() const {
		uint64_t p = 0;
		if (handle) flmip_image_device_ptr(handle, &p);
		return p;
	}
	flmip_image get_native_handle() const { return handle; }

protected:
	uint32_t mappable_last_level() const { return generate_mip_maps ? 0u : mip_level_count - 1u; }

	static std::mutex& minify_programs_mtx() {
		static std::mutex m;
		return m;
	}
	static std::unordered_map<const device_context*, std::shared_ptr<minify_program>>& minify_programs() {
		static std::unordered_map<const device_context*, std::shared_ptr<minify_program>> progs;
		return progs;
	}
	static std::shared_ptr<minify_program> lookup_minify_program(const device_context* ctx) {
		std::lock_guard<std::mutex> lock(minify_programs_mtx());
		const auto it = minify_programs().find(ctx);
		return it != minify_programs().end() ? it->second : nullptr;
	}

	//! cuda_image::create_internal (cuda_image.cpp:158-539): allocate, initial copy unless NO_INITIAL_COPY, chain if GENERATE_MIP_MAPS
	void create_internal(const device_queue& cqueue) {
		const uint32_t dim[4] = { image_dim.x, image_dim.y, image_dim.z, image_dim.w };
		const int rc = flmip_image_create(dev.device_id, image_type_bits(image_type), dim, mip_level_count, 0u, &handle);
		if (rc != FLMIP_OK) {
			FLB_LOG_ERROR("failed to create image: %s", flmip_last_error_string()); // e.g. 3-channel formats, cuda_image.cpp:173-180
			handle = nullptr;
			return;
		}
		if (host_data.data() != nullptr && !has_flag<MEMORY_FLAG::NO_INITIAL_COPY>(flags)) {
			if (flmip_image_upload(handle, host_data.data(), host_data.size_bytes(), 0, mappable_last_level(), const_cast<void*>(cqueue.get_queue_ptr())) != FLMIP_OK) {
				FLB_LOG_ERROR("initial image copy failed: %s", flmip_last_error_string());
				return;
			}
			cqueue.finish();
			if (generate_mip_maps) generate_mip_map_chain(cqueue); // cuda_image.cpp:533-536
		}
	}

	const cuda_device& dev;
	std::span<uint8_t> host_data; //!< borrowed: the caller keeps it alive (device_memory::get_host_data)
	const MEMORY_FLAG flags;
	const uint4 image_dim;
	const IMAGE_TYPE image_type;
	const bool is_mip_mapped;
	const bool generate_mip_maps;
	const uint32_t mip_level_count;
	const size_t image_data_size; //!< level 0 only for GENERATE_MIP_MAPS images (device_image.hpp:486)
	const uint32_t layer_count;
	const size_t image_data_size_mip_maps;
	const std::string debug_label;
	flmip_image handle { nullptr };
	std::mutex mappings_mtx;
	std::unordered_map<void*, MEMORY_MAP_FLAG> mappings;
};
using cuda_image = device_image;

//! Batch of independent textures: one CUDA graph with a chain of kernel nodes per image and no edges between images;
//! generate() is one graph launch, the chains of the images run concurrently.  No equivalent in the reference, which loops
//! generate_mip_map_chain over the images.  Must not outlive its images.
class mip_chain_batch {
public:
	explicit mip_chain_batch(std::span<device_image* const> images) {
		std::vector<flmip_image> handles;
		for (auto* img : images) handles.push_back(img ? img->get_native_handle() : nullptr);
		if (flmip_batch_create(handles.data(), uint32_t(handles.size()), &handle) != FLMIP_OK) {
			FLB_LOG_ERROR("failed to create mip chain batch: %s", flmip_last_error_string());
			handle = nullptr;
		}
	}
	~mip_chain_batch() {
		if (handle) flmip_batch_destroy(handle);
	}
	mip_chain_batch(const mip_chain_batch&) = delete;
	mip_chain_batch& operator=(const mip_chain_batch&) = delete;
	bool is_valid() const { return handle != nullptr; }
	//! enqueue only
	bool generate_async(const device_queue& cqueue) const {
		return handle && flmip_batch_generate(handle, const_cast<void*>(cqueue.get_queue_ptr())) == FLMIP_OK;
	}
	//! blocking, like device_image::generate_mip_map_chain
	bool generate(const device_queue& cqueue) const {
		if (!generate_async(cqueue)) return false;
		cqueue.finish();
		return true;
	}

protected:
	flmip_batch handle { nullptr };
};

//! device_context / cuda_context: constructible without floor::init (cuda_context.hpp:40-41)
class device_context {
public:
	explicit device_context(DEVICE_CONTEXT_FLAGS = DEVICE_CONTEXT_FLAGS::NONE, bool has_toolchain = false, std::vector<std::string> whitelist = {}) {
		(void)has_toolchain;
		if (flmip_init() != FLMIP_OK) {
			FLB_LOG_ERROR("CUDA is not usable: %s", flmip_last_error_string());
			return;
		}
		const int count = flmip_device_count();
		for (int i = 0; i < count; ++i) {
			flmip_device_info info;
			if (flmip_get_device_info(i, &info) != FLMIP_OK) continue;
			if (!whitelist.empty()) {
				bool found = false;
				for (const auto& w : whitelist) found |= std::string(info.name).find(w) != std::string::npos;
				if (!found) continue;
			}
			auto dev = std::make_unique<cuda_device>();
			dev->type = device::TYPE(uint32_t(device::TYPE::GPU0) + uint32_t(devices.size()));
			dev->name = info.name;
			dev->units = info.units;
			dev->clock = info.clock_mhz;
			dev->mem_clock = info.mem_clock_mhz;
			dev->mem_bus_width = info.mem_bus_width;
			dev->global_mem_size = info.global_mem_size;
			dev->max_total_local_size = info.max_total_local_size;
			dev->max_image_2d_dim = { info.max_image_2d_dim[0], info.max_image_2d_dim[1] };
			dev->max_image_3d_dim = { info.max_image_3d_dim[0], info.max_image_3d_dim[1], info.max_image_3d_dim[2] };
			dev->max_mip_levels = info.max_mip_levels;
			dev->sm = { info.sm_major, info.sm_minor };
			dev->device_id = i;
			dev->driver_version = info.driver_version;
			dev->context = this;
			devices.emplace_back(std::move(dev));
		}
		for (const auto& dev : devices) default_queues.emplace_back(std::make_shared<device_queue>(*dev));
		supported = !devices.empty();
		// fastest device = highest cores per SM x units x clock; the first one wins ties (cuda_context.cpp:340-395; 128 cores per SM on sm_100)
		for (const auto& dev : devices) {
			const uint64_t score = 128ull * dev->units * dev->clock;
			if (!fastest_gpu_device || score > fastest_gpu_score) {
				fastest_gpu_device = dev.get();
				fastest_gpu_score = score;
			}
		}
	}
	virtual ~device_context() { device_image::provide_minify_program(*this, nullptr); }

	bool is_supported() const { return supported; }
	PLATFORM_TYPE get_platform_type() const { return PLATFORM_TYPE::CUDA; }
	std::vector<const device*> get_devices() const {
		std::vector<const device*> ret;
		for (const auto& d : devices) ret.push_back(d.get());
		return ret;
	}
	//! device_context::get_device (src/device/device_context.cpp:27-72): FASTEST / FASTEST_GPU -> the fastest GPU (there are no CPU
	//! devices in a CUDA context), GPU0 + n -> that GPU, anything else / out of range -> "any" = the first device
	const device* get_device(device::TYPE type) const {
		if (devices.empty()) return nullptr;
		if (type == device::TYPE::FASTEST || type == device::TYPE::FASTEST_GPU) return fastest_gpu_device;
		const uint32_t t = uint32_t(type);
		if (t >= uint32_t(device::TYPE::GPU0) && t <= uint32_t(device::TYPE::GPU255)) {
			const uint32_t idx = t - uint32_t(device::TYPE::GPU0);
			if (idx < devices.size()) return devices[idx].get();
		}
		return devices[0].get();
	}
	std::shared_ptr<device_queue> create_queue(const device& dev) const { return std::make_shared<device_queue>(static_cast<const cuda_device&>(dev)); }
	const device_queue* get_device_default_queue(const device& dev) const {
		const auto idx = size_t(static_cast<const cuda_device&>(dev).device_id);
		for (size_t i = 0; i < devices.size(); ++i)
			if (size_t(devices[i]->device_id) == idx) return default_queues[i].get();
		return nullptr;
	}

	//! device_context.hpp:261-267
	virtual std::shared_ptr<device_image> create_image(const device_queue& cqueue, uint4 image_dim, IMAGE_TYPE image_type, std::span<uint8_t> data,
													   MEMORY_FLAG flags = MEMORY_FLAG::HOST_READ_WRITE, uint32_t mip_level_limit = 0u,
													   const char* debug_label = nullptr) const {
		auto img = std::make_shared<device_image>(cqueue, image_dim, image_type, data, flags, mip_level_limit, debug_label); // may throw (invariants)
		if (!img->is_valid()) return nullptr;
		return img;
	}
	//! uninitialized image (device_context.hpp:270-277)
	std::shared_ptr<device_image> create_image(const device_queue& cqueue, uint4 image_dim, IMAGE_TYPE image_type,
											   MEMORY_FLAG flags = MEMORY_FLAG::HOST_READ_WRITE, uint32_t mip_level_limit = 0u,
											   const char* debug_label = nullptr) const {
		return create_image(cqueue, image_dim, image_type, std::span<uint8_t> {}, flags, mip_level_limit, debug_label);
	}
	template <typename data_type>
	std::shared_ptr<device_image> create_image(const device_queue& cqueue, uint4 image_dim, IMAGE_TYPE image_type, std::span<data_type> data,
											   MEMORY_FLAG flags = MEMORY_FLAG::HOST_READ_WRITE, uint32_t mip_level_limit = 0u,
											   const char* debug_label = nullptr) const {
		return create_image(cqueue, image_dim, image_type,
							std::span<uint8_t> { reinterpret_cast<uint8_t*>(const_cast<std::remove_const_t<data_type>*>(data.data())), data.size_bytes() }, flags,
							mip_level_limit, debug_label);
	}
	template <typename data_type>
	std::shared_ptr<device_image> create_image(const device_queue& cqueue, uint4 image_dim, IMAGE_TYPE image_type, const std::vector<data_type>& data,
											   MEMORY_FLAG flags = MEMORY_FLAG::HOST_READ_WRITE, uint32_t mip_level_limit = 0u,
											   const char* debug_label = nullptr) const {
		return create_image(cqueue, image_dim, image_type,
							std::span<uint8_t> { reinterpret_cast<uint8_t*>(const_cast<data_type*>(data.data())), data.size() * sizeof(data_type) }, flags,
							mip_level_limit, debug_label);
	}

protected:
	std::vector<std::unique_ptr<cuda_device>> devices;
	std::vector<std::shared_ptr<device_queue>> default_queues;
	bool supported { false };
	const device* fastest_gpu_device { nullptr };
	uint64_t fastest_gpu_score { 0 };
};
using cuda_context = device_context;

} // namespace fl
