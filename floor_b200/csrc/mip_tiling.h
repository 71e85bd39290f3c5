// Compile-time tiling of the single-pass downsampler, shared by the kernels and the host planner.
#pragma once
#include <stdint.h>

#include "mip_params.h"

#if defined(__CUDACC__)
#define FLMIP_HD __host__ __device__
#else
#define FLMIP_HD
#endif

FLMIP_HD constexpr int flmip_elem_bytes(uint32_t ek) {
	return (ek == FLMIP_EK_F32 || ek == FLMIP_EK_U32 || ek == FLMIP_EK_I32) ? 4
		   : (ek == FLMIP_EK_F16 || ek == FLMIP_EK_UNORM16 || ek == FLMIP_EK_SNORM16 || ek == FLMIP_EK_U16 || ek == FLMIP_EK_I16) ? 2
																																   : 1;
}

FLMIP_HD constexpr int flmip_ilog2(uint32_t v) { return v <= 1 ? 0 : 1 + flmip_ilog2(v >> 1); }
FLMIP_HD constexpr uint32_t flmip_min3(uint32_t a, uint32_t b, uint32_t c) { return a < b ? (a < c ? a : c) : (b < c ? b : c); }

// Tile geometry per texel size (BPP = bytes per texel, power of two <= 16) and dimensionality.
//  2D: 512 B x 64 rows  = 32 KiB per CTA, 256 threads, each thread owns 32 B x 4 rows   (levels 1+2 in registers)
//  3D: 128 B x 16 x 16  = 32 KiB per CTA, 256 threads, each thread owns 32 B x 2 x 2    (level 1 in registers)
template <int BPP, int DIMS> struct flmip_tiling {
	static_assert(DIMS == 2 || DIMS == 3, "2D (incl. arrays / cubes) or 3D");
	static_assert(BPP == 1 || BPP == 2 || BPP == 4 || BPP == 8 || BPP == 16, "texel size");
	static constexpr uint32_t TILE_BYTES_X = (DIMS == 2 ? 512u : 128u);
	static constexpr uint32_t TX = TILE_BYTES_X / BPP;
	static constexpr uint32_t TY = (DIMS == 2 ? 64u : 16u);
	static constexpr uint32_t TZ = (DIMS == 2 ? 1u : 16u);
	static constexpr uint32_t TILE_BYTES = TILE_BYTES_X * TY * TZ;
	static constexpr uint32_t THREADS = 256u;
	static constexpr uint32_t THREADS_X = TILE_BYTES_X / 32u;       // 32 bytes (two 16-byte chunks) per thread and row
	static constexpr uint32_t THREADS_Y = (DIMS == 2 ? TY / 4u : TY / 2u);
	// levels produced in registers before the shared-memory cascade takes over
	static constexpr uint32_t IN_REG_LEVELS = (DIMS == 2) ? 2u : 1u;
	// levels one tile can finish on its own, and the texels of that level one tile holds
	static constexpr uint32_t TILE_LEVELS = (DIMS == 2 ? (uint32_t)flmip_ilog2(TX < TY ? TX : TY) : (uint32_t)flmip_ilog2(flmip_min3(TX, TY, TZ)));
	static constexpr uint32_t REM_TEXELS = (TX >> TILE_LEVELS) * (TY >> TILE_LEVELS) * (DIMS == 3 ? (TZ >> TILE_LEVELS) : 1u);
	// cascade scratch: level IN_REG_LEVELS of the tile; the group patch must fit into the same buffer
	static constexpr uint32_t CASCADE_BYTES = (TX >> IN_REG_LEVELS) * (TY >> IN_REG_LEVELS) * (DIMS == 3 ? (TZ >> IN_REG_LEVELS) : 1u) * BPP;
	static constexpr uint32_t group_for(uint32_t g) {
		return (DIMS == 2 ? g * g : g * g * g) * REM_TEXELS * BPP <= CASCADE_BYTES ? g : group_for(g / 2);
	}
	static constexpr uint32_t GROUP = group_for(DIMS == 2 ? 16u : 8u); // tiles per group and dimension
	// FLMIP_FINISHER_WARPS cascade slots (buf_a + buf_b each) + the CTA's patch buffer for the group / layer stages
	static constexpr uint32_t CASCADE_SMEM_BYTES = (FLMIP_FINISHER_WARPS + 1u) * (CASCADE_BYTES + CASCADE_BYTES / 4u);
	static constexpr uint32_t smem_bytes(uint32_t stages) { return stages * TILE_BYTES + CASCADE_SMEM_BYTES; }
	static_assert(GROUP >= 2, "group patch does not fit");
};

// run-time view of the same numbers for the host planner
struct flmip_tiling_rt {
	uint32_t tx, ty, tz, tile_levels, group, cascade_bytes, tile_bytes, cascade_smem_bytes, tile_bytes_x;
};
template <int BPP, int DIMS> constexpr flmip_tiling_rt flmip_tiling_make() {
	using T = flmip_tiling<BPP, DIMS>;
	return flmip_tiling_rt{ T::TX, T::TY, T::TZ, T::TILE_LEVELS, T::GROUP, T::CASCADE_BYTES, T::TILE_BYTES, T::CASCADE_SMEM_BYTES, T::TILE_BYTES_X };
}
inline flmip_tiling_rt flmip_tiling_lookup(uint32_t bpp, uint32_t dims) {
	if (dims == 2) {
		switch (bpp) {
			case 1: return flmip_tiling_make<1, 2>();
			case 2: return flmip_tiling_make<2, 2>();
			case 4: return flmip_tiling_make<4, 2>();
			case 8: return flmip_tiling_make<8, 2>();
			default: return flmip_tiling_make<16, 2>();
		}
	}
	switch (bpp) {
		case 1: return flmip_tiling_make<1, 3>();
		case 2: return flmip_tiling_make<2, 3>();
		case 4: return flmip_tiling_make<4, 3>();
		case 8: return flmip_tiling_make<8, 3>();
		default: return flmip_tiling_make<16, 3>();
	}
}
