/*
 * floor_b200_mip.h -- C-ABI of the B200-native mip-chain library (libfloor_b200_mip.so).
 *
 * libfloor has no C ABI / plugin interface for this path: the seam is the C++ virtual
 * fl::device_image::generate_mip_map_chain (include/floor/device/device_image.hpp:161-162).  This layer
 * sits where floor keeps its dlsym'd driver table (src/device/cuda/cuda_api.cpp:46-110): below the C++
 * device_context / device_queue / device_image classes (mirrored in include/floor_b200/) and above the
 * CUDA driver API.  Every entry point names the reference interface it replaces.
 *
 * Conventions: extern "C", plain pointers and sizes, opaque handles, int status (0 = ok, < 0 = error),
 * no exceptions; flmip_last_error_string() returns the calling thread's last message.  There is no CPU
 * fallback: without a usable libcuda / GPU every compute entry point fails with FLMIP_ERR_NO_CUDA.
 */
#ifndef FLOOR_B200_MIP_H
#define FLOOR_B200_MIP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLMIP_OK 0
#define FLMIP_ERR_NO_CUDA (-1)      /* libcuda missing / cuInit failed / no device */
#define FLMIP_ERR_INVALID (-2)      /* bad argument */
#define FLMIP_ERR_UNSUPPORTED (-3)  /* image type has no minification kernel (reference: device_image.cpp:278-283) */
#define FLMIP_ERR_DRIVER (-4)       /* a CUDA driver call failed (reference: CU_CALL_RET, cuda_common.hpp:42-61) */
#define FLMIP_ERR_OUT_OF_MEMORY (-5)

/* flags of flmip_image_create */
#define FLMIP_IMAGE_NO_DOUBLE 1u     /* FLOOR_DEVICE_NO_DOUBLE encoders for 16-bit normalized formats (host_image.hpp:398-402) */
#define FLMIP_IMAGE_FORCE_GENERIC 2u /* only the literal one-launch-per-level kernel (validation / A-B timing) */
#define FLMIP_IMAGE_UNITS_ALWAYS 4u  /* single-pass kernel: schedule 2x2(x2)-tile units whenever the tile grid allows it */
#define FLMIP_IMAGE_UNITS_NEVER 8u   /* single-pass kernel: always schedule single tiles (default: by image size) */
#define FLMIP_IMAGE_NO_TMA_TILES 32u /* tile kernel: never the persistent TMA form (flmip_ptile2d_*), always the LDG form (validation / A-B timing) */
#define FLMIP_IMAGE_TMA_TILES_ALWAYS 64u   /* tuning: the persistent TMA tile kernel wherever a level qualifies, however few tiles it has */
#define FLMIP_IMAGE_TMA_TILES_SPLIT 128u    /* tuning: always stream two levels per launch of it (default: by tiles per resident CTA) */
#define FLMIP_IMAGE_TMA_TILES_NO_SPLIT 256u /* tuning: never */
#define FLMIP_IMAGE_FORCE_TILED 16u  /* never use the persistent single-pass kernel: multi-level tile kernel for every level (validation / A-B timing) */

typedef struct flmip_image_s* flmip_image;
typedef struct flmip_batch_s* flmip_batch;
typedef void* flmip_stream; /* CUstream */
typedef void* flmip_event;  /* CUevent */

typedef struct flmip_device_info {
	char name[128];
	uint64_t global_mem_size;
	uint32_t sm_major, sm_minor;
	uint32_t units;                 /* SM count                     (fl::device::units) */
	uint32_t max_total_local_size;  /* max threads per block        (fl::device::max_total_local_size) */
	uint32_t max_image_2d_dim[2], max_image_3d_dim[3];
	uint32_t max_mip_levels;
	uint32_t driver_version;
	uint32_t clock_mhz;             /* SM clock                     (fl::device::clock; FASTEST_GPU score = units x clock, cuda_context.cpp:340-395) */
	uint32_t mem_clock_mhz;         /* fl::device::mem_clock */
	uint32_t mem_bus_width;         /* fl::device::mem_bus_width */
	uint32_t l2_cache_size;         /* bytes */
} flmip_device_info;

typedef struct flmip_level_info {
	uint32_t dim[3];      /* texels, unused dims 0; a level with a zero dim is empty (image_types.hpp:751-766) */
	uint64_t offset;      /* byte offset of the level in the level-major image */
	uint64_t size;        /* bytes of the level over all layers */
	uint64_t slice_size;  /* bytes of one layer */
} flmip_level_info;

/* -- bring-up: replaces cuda_api_init + the cuda_context ctor's device enumeration
 *    (src/device/cuda/cuda_api.cpp:53-, src/device/cuda/cuda_context.cpp:40-415) ------------------------ */
int flmip_init(void);
int flmip_device_count(void);
int flmip_get_device_info(int device, flmip_device_info* out);
const char* flmip_last_error_string(void);
/* number of kernels this library launched since load (bench bookkeeping) */
uint64_t flmip_launch_count(void);

/* -- queues: cuda_context::create_queue / cuda_queue::finish (cuda_context.cpp:418-437, cuda_queue.cpp:26-72) */
int flmip_stream_create(int device, flmip_stream* out);
int flmip_stream_destroy(int device, flmip_stream stream);
int flmip_stream_sync(int device, flmip_stream stream);
/* Chains on INDEPENDENT images overlap (opt-in per stream; no counterpart in the reference, whose generate_mip_map_chain blocks the host per
 * (layer, level) launch, device_image.cpp:304-327).  With overlap enabled, the first kernel of a chain whose image has no kernel in
 * the stream's open run (the chain kernels enqueued since the last one that waited for its predecessor) starts while the tail of the chain in
 * front of it is still running (last units, group / layer stages: microseconds without memory traffic), and waits for that chain before it
 * ENDS instead, so completion still follows stream order: whatever is enqueued behind a chain sees every chain before it complete.
 * A chain on an image that is still in the open run and anything else this library enqueues on the stream (copies, fills, events,
 * batches) wait as before; the later kernels of a multi-kernel chain (NPOT images) wait for the kernel they depend on, but release
 * their own dependents first, so that the next image's chain need not wait for them either.  The library cannot see work the caller enqueues on the stream by
 * other means: announce it with flmip_stream_fence(stream) AFTER enqueueing it and before the next chain (or leave overlap off, the
 * default).  Measured on a B200 (chains of different images back to back): 8192^2 RGBA16F 111.7 -> 102.5 us, 1024^2 RGBA8 10.5 -> 3.9 us. */
int flmip_stream_set_chain_overlap(int device, flmip_stream stream, int enable);
int flmip_stream_fence(int device, flmip_stream stream);
/* test hook (no GPU needed): drives the host bookkeeping behind the overlap -- op 0: opt `stream` in / out (arg), 1: a chain of `arg`
 * kernels on `image` (returns 1 if its first kernel would start without waiting), 2: the same for a chain the literal kernel starts,
 * 3: anything else enqueued on the stream, 4: forget the stream */
int flmip_overlap_bookkeeping(uint64_t stream, uint64_t image, uint32_t op, uint32_t arg);
/* profiling: cuda_queue::start_profiling / stop_profiling (cuda_queue.cpp:58-70) */
int flmip_event_create(int device, flmip_event* out);
int flmip_event_record(int device, flmip_event ev, flmip_stream stream);
int flmip_event_sync(int device, flmip_event ev);
int flmip_event_elapsed_ms(int device, flmip_event start, flmip_event stop, float* ms);
int flmip_event_destroy(int device, flmip_event ev);
/* page-locked host staging buffers; FLMIP_HOST_WRITE_COMBINED: upload-only buffers (fast for the GPU to read over PCIe, slow for
 * the CPU to read) */
#define FLMIP_HOST_WRITE_COMBINED 1u
int flmip_host_alloc(int device, size_t size, void** out);
int flmip_host_alloc_ex(int device, size_t size, uint32_t flags, void** out);
int flmip_host_free(int device, void* ptr);

/* -- images: cuda_image::create_internal (src/device/cuda/cuda_image.cpp:158-539), but as ONE linear
 *    allocation in floor's host layout (level-major, layers contiguous per level, tight rows:
 *    device_image.hpp:594-615, host_image.cpp:81-88) with 64-bit offsets.
 *    image_dim = (w, h, d or layers, layers-for-3D); cube: 6 faces per layer (image_types.hpp:716-726). */
int flmip_image_create(int device, uint64_t image_type, const uint32_t image_dim[4], uint32_t mip_level_limit, uint32_t flags,
					   flmip_image* out);
/* Opt-in for other contexts (the counterpart of device_image::provide_minify_program, src/device/device_image.cpp:155-194, which
 * registers a minify program per device_context): a context that already owns its CUcontext and its image memory hands both in
 * instead of adopting this library's allocations.
 *   flmip_device_attach_context: run `device` on the caller's CUcontext (floor's cuda_context creates its own with cuCtxCreate,
 *     cuda_context.cpp:117-124) instead of the primary one; must precede the first use of the device.
 *   flmip_image_create_external: an image over caller-owned LINEAR device memory in floor's host layout (level-major, tight rows,
 *     `size` >= image_data_size over all levels, 16-byte aligned); destroy releases the handle and its counters, never the memory.
 * Opaque tiled images (CUmipmappedArray) go through flmip_image_copy_{from,to}_tiled below. */
int flmip_device_attach_context(int device, void* cu_context);
int flmip_image_create_external(int device, uint64_t image_type, const uint32_t image_dim[4], uint32_t mip_level_limit, uint32_t flags,
								uint64_t device_ptr, uint64_t size, flmip_image* out);
int flmip_image_destroy(flmip_image img);
int flmip_image_mip_level_count(flmip_image img, uint32_t* out);
int flmip_image_layer_count(flmip_image img, uint32_t* out);
int flmip_image_data_size(flmip_image img, uint64_t* out);          /* all levels */
int flmip_image_get_level_info(flmip_image img, uint32_t level, flmip_level_info* out);
int flmip_image_device_ptr(flmip_image img, uint64_t* out);         /* CUdeviceptr of level 0 */
/* 1 if generate uses the single-pass kernel; *fast_levels = number of levels (incl. level 0) it covers */
int flmip_image_plan(flmip_image img, uint32_t* uses_single_pass, uint32_t* fast_levels, uint32_t* launches);
/* one entry of the sampler table the persistent tile kernel reads (host arithmetic only, no GPU needed): the reference's linear
 * fetch for destination texel g along an axis whose source level has n texels (mip_map_minify.hpp:106, host_image.hpp:141-174,
 * 869-894): fp32 bits of the weight t of the active texel B; bit 31: A = 2g + 1, B = 2g (else A = 2g, B = 2g + 1); bit 30: the
 * texel-2 fetch (g = 0: A = 2, B = 0) */
int flmip_sampler_table_entry(uint32_t g, uint32_t n, uint32_t* out);
/* number of those launches that are the persistent TMA tile kernel (flmip_ptile2d_*) */
int flmip_image_plan_tma_tile_launches(flmip_image img, uint32_t* out);

/* -- host <-> device: cuda_image::write / map / unmap copies (cuda_image.cpp:588-813).
 *    Whole levels [level_first, level_last] (inclusive, like mip_level_range) in host layout; async on `stream`. */
int flmip_image_upload(flmip_image img, const void* src, size_t src_size, uint32_t level_first, uint32_t level_last, flmip_stream stream);
int flmip_image_download(flmip_image img, void* dst, size_t dst_size, uint32_t level_first, uint32_t level_last, flmip_stream stream);
/* layers [layer_first, layer_first + layer_count) of levels [level_first, level_last], level-major like an image of layer_count
 * layers (validation of sampled layers of images too large to read back whole; the reference's map() has no ranges) */
int flmip_image_download_layers(flmip_image img, void* dst, size_t dst_size, uint32_t level_first, uint32_t level_last, uint32_t layer_first,
								uint32_t layer_count, flmip_stream stream);
/* sub-region write with floor's semantics (offset / extent in level-0 texels, inclusive level and layer ranges,
 * source tightly packed per level): cuda_image::write, cuda_image.cpp:588-673.  Every level of the range is validated before the
 * first copy: a region that leaves a level (odd dims: offset >> level + max(extent >> level, 1) > dim >> level) fails with
 * FLMIP_ERR_INVALID and writes nothing -- the reference's cuMemcpy3D into the level's CUarray fails at that level. */
int flmip_image_write(flmip_image img, const void* src, size_t src_size, const uint32_t offset[3], const uint32_t extent[3],
					  const uint32_t mip_level_range[2], const uint32_t layer_range[2], flmip_stream stream);
int flmip_image_zero(flmip_image img, flmip_stream stream);          /* cuda_image::zero, cuda_image.cpp:675-701 */
/* device_image::blit (device_image.hpp:96-101, checks of device_image.cpp:470-501): every level both images have, device to
 * device.  The reference's CUDA image inherits the `return false` stub, so its clone(copy_contents) copies nothing. */
int flmip_image_blit(flmip_image dst, flmip_image src, flmip_stream stream);

/* -- interop with floor's tiled CUDA images (CUmipmappedArray sampled through texture / surface objects,
 *    cuda_image.cpp:158-539): create an array with the reference's descriptor for this image, copy whole levels between
 *    the linear image and the array (one cuMemcpy3DAsync per level), read an array back to the host in floor's layout.
 *    `mipmapped_array` is a CUmipmappedArray; arrays created by the original cuda_image can be passed as they are. */
int flmip_image_create_tiled_twin(flmip_image img, void** out_mipmapped_array);
int flmip_tiled_destroy(int device, void* mipmapped_array);
int flmip_image_copy_to_tiled(flmip_image img, void* mipmapped_array, uint32_t level_first, uint32_t level_last, flmip_stream stream);
int flmip_image_copy_from_tiled(flmip_image img, void* mipmapped_array, uint32_t level_first, uint32_t level_last, flmip_stream stream);
int flmip_tiled_download(flmip_image geometry, void* mipmapped_array, void* dst, size_t dst_size, uint32_t level_first, uint32_t level_last,
						 flmip_stream stream);
/* the CUcontext the library uses on `device` (cuda_device::ctx, include/floor/device/cuda/cuda_device.hpp:30-86) */
int flmip_device_cu_context(int device, void** out);

/* -- THE hot path: device_image::generate_mip_map_chain (src/device/device_image.cpp:235-328) + the
 *    libfloor_mip_map_minify_* kernels (include/floor/device/backend/mip_map_minify.hpp:89-126).
 *    Enqueues on `stream` and returns; the C++ drop-in adds the blocking flmip_stream_sync the reference
 *    implies (wait_until_completion = true, device_image.cpp:322).
 *    ONE chain per image is in flight at a time (the launch shares the image's tile / group / layer counters): chains enqueued on
 *    one stream are ordered by the stream; a chain enqueued on a different stream than the image's previous one first makes that
 *    stream wait for everything enqueued on the old stream so far (event hand-over inside this call; also across
 *    flmip_stream_destroy and flmip_batch_generate).  Callable from any thread; calls on one image serialise on a per-image lock.
 *    first_level >= mip level count fails with FLMIP_ERR_INVALID; first_level == last level generates nothing. */
int flmip_mip_chain_generate(flmip_image img, flmip_stream stream);
/* regenerate only levels > first_level (dirty-level update); first_level = 0 is the full chain */
int flmip_mip_chain_generate_from(flmip_image img, uint32_t first_level, flmip_stream stream);

/* -- batches of independent textures (the reference loops generate_mip_map_chain over them, device_image.cpp:304-327
 *    per image): one CUDA graph with a chain of kernel nodes per image and no edges between images.  flmip_batch_generate is
 *    ONE graph launch on `stream`; the chains of the images run concurrently.  All images on one device, each at most once;
 *    destroy the batch before any of its images. */
int flmip_batch_create(const flmip_image* images, uint32_t count, flmip_batch* out);
int flmip_batch_generate(flmip_batch batch, flmip_stream stream);
int flmip_batch_kernel_count(flmip_batch batch, uint32_t* out); /* kernel nodes one flmip_batch_generate runs */
int flmip_batch_destroy(flmip_batch batch);

/* -- bench / validation helper: fill level 0 with the counter-based synthetic pattern of SURVEY.md 8d
 *    (the CPU checker defines the same pattern); global layer ids start at layer_id0 */
int flmip_image_fill_synthetic(flmip_image img, uint64_t config_id, uint64_t layer_id0, flmip_stream stream);

#ifdef __cplusplus
}
#endif
#endif
