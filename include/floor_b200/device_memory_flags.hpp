// fl::MEMORY_FLAG / fl::MEMORY_MAP_FLAG, bit-identical to include/floor/device/device_memory_flags.hpp:28-163 of a2flo/floor
// (only the flags this path reads carry behaviour here; the others are accepted and ignored).
#pragma once
#include <cstdint>

#include "image_types.hpp"

namespace fl {

enum class MEMORY_FLAG : uint32_t {
	NONE = 0u,
	READ = 1u << 0, WRITE = 1u << 1, READ_WRITE = READ | WRITE,
	HOST_READ = 1u << 2, HOST_WRITE = 1u << 3, HOST_READ_WRITE = HOST_READ | HOST_WRITE,
	NO_INITIAL_COPY = 1u << 4,
	HOST_READ_BACK_OPTIMIZE = 1u << 5, HOST_READ_STAGING = 1u << 6, USE_HOST_MEMORY = 1u << 7, RENDER_TARGET = 1u << 8,
	GENERATE_MIP_MAPS = 1u << 9, //!< generate the mip chain at creation, after write() and after unmap()
	VULKAN_SHARING = 1u << 10, METAL_SHARING = 1u << 11, VULKAN_SHARING_SYNC_SHARED = 1u << 12, METAL_SHARING_SYNC_SHARED = 1u << 13,
	VULKAN_ALIASING = 1u << 14, VULKAN_HOST_COHERENT = 1u << 15, NO_RESOURCE_TRACKING = 1u << 16, VULKAN_DESCRIPTOR_BUFFER = 1u << 17,
	SHARING_SYNC = 1u << 18, SHARING_RENDER_READ = 1u << 19, SHARING_RENDER_WRITE = 1u << 20,
	SHARING_RENDER_READ_WRITE = SHARING_RENDER_READ | SHARING_RENDER_WRITE,
	SHARING_COMPUTE_READ = 1u << 21, SHARING_COMPUTE_WRITE = 1u << 22, SHARING_COMPUTE_READ_WRITE = SHARING_COMPUTE_READ | SHARING_COMPUTE_WRITE,
	HEAP_ALLOCATION = 1u << 23, NO_HEAP_ALLOCATION = 1u << 24, VULKAN_MAY_USE_HOST_MEMORY = 1u << 25,
};
FLB_ENUM_OPS(MEMORY_FLAG)

enum class MEMORY_MAP_FLAG : uint32_t {
	NONE = 0u, READ = 1u << 0, WRITE = 1u << 1, WRITE_INVALIDATE = 1u << 2, READ_WRITE = READ | WRITE, BLOCK = 1u << 3,
};
FLB_ENUM_OPS(MEMORY_MAP_FLAG)

static_assert(static_cast<uint32_t>(MEMORY_FLAG::GENERATE_MIP_MAPS) == 0x200u && static_cast<uint32_t>(MEMORY_FLAG::NO_INITIAL_COPY) == 0x10u);

} // namespace fl
