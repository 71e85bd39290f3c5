// Parameter blocks shared by the host C-ABI layer (flmip.cpp) and the sm_100a kernels (mip_kernels.cu).
// Plain C structs: they are passed by value through cuLaunchKernelEx.
#pragma once
#include <stdint.h>

#define FLMIP_MAX_LEVELS 16u // == host_limits::max_mip_levels of the reference (host_image.hpp:462-479)
#define FLMIP_MAX_STAGES 8u  // upper bound of the TMA tile ring of the single-pass kernel
// persistent single-pass CTA = 8 consumer warps + 1 TMA producer warp + FLMIP_FINISHER_WARPS finisher warps
#define FLMIP_FINISHER_WARPS 4
#ifndef FLMIP_SCHED_PREFETCH
#define FLMIP_SCHED_PREFETCH 4u   // tile-index fetches the producer keeps in flight
#endif
#ifndef FLMIP_UNIT_PREFETCH
#define FLMIP_UNIT_PREFETCH 1u    // ... when the work items are units of 2 x 2 (x 2) tiles: a unit outlasts the round trip, and every prefetched unit is committed to its CTA (2 -> 1: C3 +0.7 %, C5 +0.1 %, C2 +-0, profiles/r2/06_unit_prefetch_ab.txt)
#endif
#define FLMIP_UNIT_PATCH_BYTES 512u    // remainders of the tiles of one unit: at most 8 x (1 x 2 x 2 texels of 16 bytes)
#define FLMIP_NO_TILE 0xFFFFFFFFu // end-of-work sentinel in the tile ring / cascade slots
#define FLMIP_BLOCK_THREADS ((8 + 1 + FLMIP_FINISHER_WARPS) * 32)

// storage element kinds the path supports (reference format LUT: src/device/cuda/cuda_image.cpp:197-248,
// normalized variants: include/floor/device/backend/host_image.hpp:487-561)
enum flmip_elem_kind : uint32_t {
	FLMIP_EK_F32 = 0,
	FLMIP_EK_F16,
	FLMIP_EK_UNORM8,
	FLMIP_EK_SNORM8,
	FLMIP_EK_UNORM16,
	FLMIP_EK_SNORM16,
	FLMIP_EK_U8,
	FLMIP_EK_I8,
	FLMIP_EK_U16,
	FLMIP_EK_I16,
	FLMIP_EK_U32,
	FLMIP_EK_I32,
	FLMIP_EK_COUNT, // kinds the single-pass / tile kernels are instantiated for
	// FORMAT_2 / FORMAT_4 normalized formats (host_image.hpp:341-353, 419-446): 2 / 4 bits per channel packed into whole bytes
	// (RGBA2: 1 byte, RG4: 1 byte, RGBA4: 2 bytes); served by the literal kernel (flmip_generic) only
	FLMIP_EK_UNORM4 = FLMIP_EK_COUNT,
	FLMIP_EK_SNORM4,
	FLMIP_EK_UNORM2,
	FLMIP_EK_SNORM2
};

// one destination level, any size (NPOT capable): the general path
struct flmip_generic_params {
	uint64_t base;                 // device address of the image
	uint64_t src_off, dst_off;     // byte offsets of level-1 / level
	uint64_t src_slice, dst_slice; // bytes of one layer in either level
	uint64_t total;                // dst texels per layer * layers
	uint32_t src_dim[3], dst_dim[3];
	float inv_prev[3];  // 1.0f / float(src_dim)   (device_image.cpp:311-312)
	float fdim[3];      // float(src_dim)          (host_image.cpp:96-101)
	float fdim_excl[3]; // nextafterf(fdim, 0)     (host_image.cpp:102-107)
	uint32_t dc, layers, elem_kind, channels, no_double;
};

// single-pass multi-level downsampler (power-of-two images)
struct flmip_fast_params {
	uint64_t base;
	uint64_t level_off[FLMIP_MAX_LEVELS];
	uint64_t counters;  // uint32[layers * groups] group counters, then uint32[layers] layer counters
	uint64_t sched;     // uint32[2]: next tile to hand out, producers done (both left at zero by the kernel)
	uint32_t dim[3];    // level-0 size in texels (z = 1 for 2D)
	uint32_t tiles[3];  // tiles per layer
	uint32_t groups[3]; // tile groups per layer
	uint32_t units[3];      // units per layer: tiles >> unit_shift (a unit = 2 x 2 (x 2) tiles, or one tile)
	uint32_t unit_cshift[3]; // log2(units): the single-pass kernel only runs on power-of-two images
	uint32_t unit_shift;    // 1: units of 2 x 2 (x 2) tiles, 0: units of one tile (some dim is a single tile wide)
	uint32_t layers, level_count, no_double;
	uint32_t total_units;   // units per layer * layers, handed out by the dynamic scheduler
	uint32_t stages;        // depth of the TMA tile ring in shared memory (<= FLMIP_MAX_STAGES)
	uint32_t late_wait;     // set per launch: the kernel before this one in the stream is a chain on another image (pdl_start / pdl_late_wait)
};

// multi-level tile kernel, any size (NPOT capable): one CTA reduces one source tile of level [0] through `nlev` levels.
// Source tile = 64 x 64 texels (2D, up to 6 levels) or 32 x 16 x 16 texels (3D, up to 4 levels).
#define FLMIP_TILE_MAX_LEVELS 6u
#define FLMIP_TILE2D_X 64u
#define FLMIP_TILE2D_Y 64u
#define FLMIP_TILE3D_X 32u
#define FLMIP_TILE3D_Y 16u
#define FLMIP_TILE3D_Z 16u
#define FLMIP_TILE3D_MAX_LEVELS 4u
struct flmip_tile_params {
	uint64_t base;
	uint64_t level_off[FLMIP_TILE_MAX_LEVELS + 1]; // [0] = source level, [k] = k-th produced level
	uint64_t slice[FLMIP_TILE_MAX_LEVELS + 1];     // bytes of one layer
	uint32_t dim[FLMIP_TILE_MAX_LEVELS + 1][3];    // texels (z = 1 for 2D)
	// sampler constants of produced level k, taken from its source level k - 1 (index k - 1):
	float inv_prev[FLMIP_TILE_MAX_LEVELS][3];  // 1.0f / float(dim)  (device_image.cpp:311-312)
	float fdim[FLMIP_TILE_MAX_LEVELS][3];      // float(dim)         (host_image.cpp:96-101)
	float fdim_excl[FLMIP_TILE_MAX_LEVELS][3]; // nextafterf(fdim,0) (host_image.cpp:102-107)
	uint32_t tiles[3];                         // source tiles per layer
	uint32_t nlev, layers, no_double;
	uint32_t block_sync; // 1: some produced level >= 2 contains a texel-2 fetch (block barriers instead of warp barriers)
	uint32_t late_wait;  // as in flmip_fast_params
};

// persistent TMA tile kernel, 2D images of any size whose source rows are 16-byte multiples (flmip_ptile2d_*): the single-pass
// structure of flmip_fast2d_* (TMA ring, warp-specialised CTA, finisher pool, last tile of a layer finishes the chain) with the
// general sampler weights of the reference, read from a per-image table instead of being 0.5.
// Sampler table: one uint32 per destination texel coordinate, level and axis = the fp32 weight t of the "active" texel B
// (0 < t <= 1, so bits 31 and 30 are free): bit 31 = roles swapped (A = 2g + 1, B = 2g instead of A = 2g, B = 2g + 1),
// bit 30 = the texel-2 fetch of the reference (g = 0: A = 2, B = 0); see axis_fetch() in mip_kernels.cu.
#ifndef FLMIP_PTILE_PATCH_BYTES
#define FLMIP_PTILE_PATCH_BYTES 16384u // the level the tiles end on, of one layer, must fit here for the chain to finish in one launch
#endif
#ifndef FLMIP_PTILE_TAIL_TAB
#define FLMIP_PTILE_TAIL_TAB 1024u      // sampler entries (x + y) of one level the last-tile stage keeps in shared memory
#endif
#define FLMIP_WTAB_SWAP 0x80000000u
#define FLMIP_WTAB_TEXEL2 0x40000000u
struct flmip_ptile_params {
	uint64_t base;
	uint64_t level_off[FLMIP_MAX_LEVELS]; // byte offset of every level of the image (absolute level numbers)
	uint64_t wtab;                        // device address of the sampler table
	uint64_t counters;                    // uint32[layers]: tiles of the layer that have finished their tile stage
	uint64_t sched;                       // uint32[2]: dynamic tile scheduler (as in flmip_fast_params)
	uint32_t wtab_off[FLMIP_MAX_LEVELS][2]; // first entry of (destination level, axis) in the table
	uint32_t dim[FLMIP_MAX_LEVELS][2];    // texels of every level
	uint32_t src_level;                   // the level the tensor map covers (0, or where a previous launch stopped)
	uint32_t tile_last;                   // last level the tiles produce on their own (<= src_level + 6)
	uint32_t last_level;                  // last level of this launch (> tile_last: the last tile of a layer carries on)
	uint32_t tiles[2], layers, total_tiles, stages, no_double;
	uint32_t unit_shift;                  // log2(tiles per unit): 2 when every CTA has many tiles (one publish per 4 tiles), else 0
	uint32_t vec1, vec2;                  // rows of level src + 1 / src + 2 start on 16- / 8-byte (16-byte texels: 16-byte) boundaries: full-width vector stores
	uint32_t late_wait;                   // as in flmip_fast_params
};

struct flmip_fill_params {
	uint64_t dst;             // device address of level 0 of the first layer to fill
	uint64_t elems_per_layer; // texels * channels
	uint64_t config_id, layer_id0;
	uint32_t layers, elem_kind;
};
