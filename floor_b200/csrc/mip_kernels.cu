// sm_100a kernel library replacing libfloor's toolchain-compiled `libfloor_mip_map_minify_*` kernels
// (reference: include/floor/device/backend/mip_map_minify.hpp:89-126) for the CUDA backend.
//
//  * flmip_fast2d_* / flmip_fast3d_* : single-pass multi-level downsampler for power-of-two images.
//      TMA tile load -> 2x2 (2x2x2) reductions in registers / shared memory through all levels the tile
//      covers -> 16-byte vector stores -> last-CTA counters finish the tail levels in the same launch.
//  * flmip_generic                   : one destination level per launch, any size (NPOT), replaying the
//      Host-Compute sampler arithmetic literally (host_image.hpp:842-929).
//  * flmip_fill                      : counter-based synthetic level-0 data.
//
// Arithmetic contract (bit-exact with oracle/minify_oracle.c): every level is computed from the *stored*
// (quantised) previous level; x-lerp, then y, then z, each L(a,b) = (b - a) * 0.5f + a with a = even texel;
// encoders truncate.  No tensor cores: this is not a contraction.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "mip_params.h"
#include "mip_tiling.h"

namespace {

// ------------------------------------------------------------------------------------------------------
// element codecs
// ------------------------------------------------------------------------------------------------------
template <uint32_t EK> struct Codec {
	static constexpr bool IS_INT = (EK >= FLMIP_EK_U8);
	static constexpr bool IS_SIGNED_INT = (EK == FLMIP_EK_I8 || EK == FLMIP_EK_I16 || EK == FLMIP_EK_I32);
	static constexpr int BYTES = flmip_elem_bytes(EK);
	static constexpr uint32_t MASK = (BYTES == 4 ? 0xFFFFFFFFu : (BYTES == 2 ? 0xFFFFu : 0xFFu));

	// decode zero-extended storage bits into the compute domain (fp32 bits or widened 32-bit integer)
	// host_image.hpp:487-561 (float / normalized), :640-667 (int / uint)
	static __device__ __forceinline__ uint32_t dec(uint32_t raw) {
		if constexpr (EK == FLMIP_EK_F32 || EK == FLMIP_EK_U32 || EK == FLMIP_EK_I32 || EK == FLMIP_EK_U8 || EK == FLMIP_EK_U16) {
			return raw;
		} else if constexpr (EK == FLMIP_EK_F16) {
			return __float_as_uint(__half2float(__ushort_as_half((unsigned short)raw)));
		} else if constexpr (EK == FLMIP_EK_UNORM8) {
			return __float_as_uint(__fmul_rn(__uint2float_rn(raw), (float)(1.0 / 255.0)));
		} else if constexpr (EK == FLMIP_EK_SNORM8) {
			return __float_as_uint(__fmul_rn(__int2float_rn((int)(signed char)raw), (float)(1.0 / 127.0)));
		} else if constexpr (EK == FLMIP_EK_UNORM16) {
			return __float_as_uint(__fmul_rn(__uint2float_rn(raw), (float)(1.0 / 65535.0)));
		} else if constexpr (EK == FLMIP_EK_SNORM16) {
			return __float_as_uint(__fmul_rn(__int2float_rn((int)(short)raw), (float)(1.0 / 32767.0)));
		} else if constexpr (EK == FLMIP_EK_I8) {
			return (uint32_t)(int)(signed char)raw;
		} else { // I16
			return (uint32_t)(int)(short)raw;
		}
	}

	// encode back to storage bits: host_image.hpp:672-722 + insert_channels :391-460 (truncating), :801-825
	static __device__ __forceinline__ uint32_t enc(uint32_t v, uint32_t no_double) {
		if constexpr (EK == FLMIP_EK_F32 || EK == FLMIP_EK_U32 || EK == FLMIP_EK_I32) {
			return v;
		} else if constexpr (IS_INT) {
			return v & MASK;
		} else if constexpr (EK == FLMIP_EK_F16) {
			return (uint32_t)__half_as_ushort(__float2half_rn(__uint_as_float(v)));
		} else if constexpr (EK == FLMIP_EK_UNORM8) {
			return (uint32_t)__float2int_rz(__fmul_rn(__uint_as_float(v), 255.0f)) & 0xFFu;
		} else if constexpr (EK == FLMIP_EK_SNORM8) {
			return (uint32_t)__float2int_rz(__fmul_rn(__uint_as_float(v), 127.0f)) & 0xFFu;
		} else {
			// 9..16 bit normalized: fp_scale_type is double unless FLOOR_DEVICE_NO_DOUBLE (host_image.hpp:398-402)
			constexpr float scale_f = (EK == FLMIP_EK_UNORM16 ? 65535.0f : 32767.0f);
			constexpr double scale_d = (EK == FLMIP_EK_UNORM16 ? 65535.0 : 32767.0);
			const float f = __uint_as_float(v);
			const int q = no_double ? __float2int_rz(__fmul_rn(f, scale_f)) : __double2int_rz(__dmul_rn((double)f, scale_d));
			return (uint32_t)q & 0xFFFFu;
		}
	}

	// const_math.hpp:981-996 with t = 0.5 (power-of-two levels)
	static __device__ __forceinline__ uint32_t lerp_half(uint32_t a, uint32_t b) {
		if constexpr (!IS_INT) {
			const float fa = __uint_as_float(a), fb = __uint_as_float(b);
			if constexpr (EK == FLMIP_EK_F32) {
				// arbitrary fp32 inputs: keep the three roundings separate ((b-a)*0.5 may be subnormal)
				return __float_as_uint(__fadd_rn(__fmul_rn(__fsub_rn(fb, fa), 0.5f), fa));
			} else {
				// operands come from <= 16-bit storage: (b-a)*0.5 is exact, so one FMA gives the same bits
				return __float_as_uint(__fmaf_rn(__fsub_rn(fb, fa), 0.5f, fa));
			}
		} else {
			const uint32_t d = b - a; // in T: unsigned wraps
			if constexpr (IS_SIGNED_INT) {
				return (uint32_t)__float2int_rz(__fmul_rn(__int2float_rn((int)d), 0.5f)) + a;
			} else {
				return __float2uint_rz(__fmul_rn(__uint2float_rn(d), 0.5f)) + a;
			}
		}
	}

	// general weight (NPOT levels)
	static __device__ __forceinline__ uint32_t lerp_t(uint32_t a, uint32_t b, float t) {
		if constexpr (!IS_INT) {
			const float fa = __uint_as_float(a), fb = __uint_as_float(b);
			return __float_as_uint(__fadd_rn(__fmul_rn(__fsub_rn(fb, fa), t), fa));
		} else {
			const uint32_t d = b - a;
			const float s = __fmul_rn(IS_SIGNED_INT ? __int2float_rn((int)d) : __uint2float_rn(d), t);
			return (uint32_t)__float2ll_rz(s) + a; // via 64 bit like the oracle: no saturation surprises
		}
	}
};

template <int BYTES> __device__ __forceinline__ uint32_t get_elem(const uint32_t* w, int i) {
	if constexpr (BYTES == 4) return w[i];
	else if constexpr (BYTES == 2) return (w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
	else return (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
}
template <int BYTES> __device__ __forceinline__ void put_elem(uint32_t* w, int i, uint32_t raw) {
	if constexpr (BYTES == 4) w[i] = raw;
	else if constexpr (BYTES == 2) w[i >> 1] |= raw << (16 * (i & 1));
	else w[i >> 2] |= raw << (8 * (i & 3));
}

// texel load / store by size
template <int BPP> struct TexelIO {
	static constexpr int NW = (BPP + 3) / 4;
	template <bool CG> static __device__ __forceinline__ void load(const uint8_t* p, uint32_t (&w)[NW]) {
		if constexpr (BPP == 1) { w[0] = CG ? __ldcg(p) : *p; }
		else if constexpr (BPP == 2) { w[0] = CG ? __ldcg((const unsigned short*)p) : *(const unsigned short*)p; }
		else if constexpr (BPP == 4) { w[0] = CG ? __ldcg((const uint32_t*)p) : *(const uint32_t*)p; }
		else if constexpr (BPP == 8) {
			const uint2 v = CG ? __ldcg((const uint2*)p) : *(const uint2*)p;
			w[0] = v.x; w[1] = v.y;
		} else {
			const uint4 v = CG ? __ldcg((const uint4*)p) : *(const uint4*)p;
			w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
		}
	}
	static __device__ __forceinline__ void store(uint8_t* p, const uint32_t (&w)[NW]) {
		if constexpr (BPP == 1) *p = (uint8_t)w[0];
		else if constexpr (BPP == 2) *(unsigned short*)p = (unsigned short)w[0];
		else if constexpr (BPP == 4) *(uint32_t*)p = w[0];
		else if constexpr (BPP == 8) *(uint2*)p = make_uint2(w[0], w[1]);
		else *(uint4*)p = make_uint4(w[0], w[1], w[2], w[3]);
	}
};

// ------------------------------------------------------------------------------------------------------
// packed-row reductions: rows of NW 32-bit words holding whole texels -> NW/2 words of the next level
// ------------------------------------------------------------------------------------------------------
template <uint32_t EK, int CH, int NW>
__device__ __forceinline__ void reduce_rows_2d(const uint32_t (&r0)[NW], const uint32_t (&r1)[NW], uint32_t (&out)[NW / 2],
											   uint32_t no_double) {
	using C = Codec<EK>;
	constexpr int NT = NW * 4 / (C::BYTES * CH); // texels per source row
	static_assert(NT >= 2 && (NT % 2) == 0, "row must hold at least one x pair");
#pragma unroll
	for (int i = 0; i < NW / 2; ++i) out[i] = 0;
#pragma unroll
	for (int p = 0; p < NT / 2; ++p) {
#pragma unroll
		for (int c = 0; c < CH; ++c) {
			const uint32_t x0 = C::lerp_half(C::dec(get_elem<C::BYTES>(r0, (2 * p) * CH + c)), C::dec(get_elem<C::BYTES>(r0, (2 * p + 1) * CH + c)));
			const uint32_t x1 = C::lerp_half(C::dec(get_elem<C::BYTES>(r1, (2 * p) * CH + c)), C::dec(get_elem<C::BYTES>(r1, (2 * p + 1) * CH + c)));
			put_elem<C::BYTES>(out, p * CH + c, C::enc(C::lerp_half(x0, x1), no_double));
		}
	}
}

// r[z][y]: rows (y, y+1) of slices (z, z+1)
template <uint32_t EK, int CH, int NW>
__device__ __forceinline__ void reduce_rows_3d(const uint32_t (&r00)[NW], const uint32_t (&r01)[NW], const uint32_t (&r10)[NW],
											   const uint32_t (&r11)[NW], uint32_t (&out)[NW / 2], uint32_t no_double) {
	using C = Codec<EK>;
	constexpr int NT = NW * 4 / (C::BYTES * CH);
	static_assert(NT >= 2 && (NT % 2) == 0, "row must hold at least one x pair");
#pragma unroll
	for (int i = 0; i < NW / 2; ++i) out[i] = 0;
#pragma unroll
	for (int p = 0; p < NT / 2; ++p) {
#pragma unroll
		for (int c = 0; c < CH; ++c) {
			const int ea = (2 * p) * CH + c, eb = (2 * p + 1) * CH + c;
			const uint32_t x00 = C::lerp_half(C::dec(get_elem<C::BYTES>(r00, ea)), C::dec(get_elem<C::BYTES>(r00, eb)));
			const uint32_t x01 = C::lerp_half(C::dec(get_elem<C::BYTES>(r01, ea)), C::dec(get_elem<C::BYTES>(r01, eb)));
			const uint32_t x10 = C::lerp_half(C::dec(get_elem<C::BYTES>(r10, ea)), C::dec(get_elem<C::BYTES>(r10, eb)));
			const uint32_t x11 = C::lerp_half(C::dec(get_elem<C::BYTES>(r11, ea)), C::dec(get_elem<C::BYTES>(r11, eb)));
			const uint32_t y0 = C::lerp_half(x00, x01), y1 = C::lerp_half(x10, x11);
			put_elem<C::BYTES>(out, p * CH + c, C::enc(C::lerp_half(y0, y1), no_double));
		}
	}
}

// one destination texel from 4 / 8 individually addressed source texels (cascade levels)
template <uint32_t EK, int CH, int DIMS, bool CG>
__device__ __forceinline__ void reduce_texel(const uint8_t* src, uint32_t row_pitch, uint32_t slice_pitch,
											 uint32_t (&out)[TexelIO<Codec<EK>::BYTES * CH>::NW], uint32_t no_double) {
	using C = Codec<EK>;
	constexpr int BPP = C::BYTES * CH;
	using IO = TexelIO<BPP>;
	uint32_t t[DIMS == 3 ? 8 : 4][IO::NW];
#pragma unroll
	for (int k = 0; k < (DIMS == 3 ? 8 : 4); ++k) {
		IO::template load<CG>(src + (k & 1) * BPP + ((k >> 1) & 1) * (size_t)row_pitch + (k >> 2) * (size_t)slice_pitch, t[k]);
	}
#pragma unroll
	for (int i = 0; i < IO::NW; ++i) out[i] = 0;
#pragma unroll
	for (int c = 0; c < CH; ++c) {
		uint32_t v[DIMS == 3 ? 8 : 4];
#pragma unroll
		for (int k = 0; k < (DIMS == 3 ? 8 : 4); ++k) v[k] = C::dec(get_elem<C::BYTES>(t[k], c));
		uint32_t r = C::lerp_half(C::lerp_half(v[0], v[1]), C::lerp_half(v[2], v[3]));
		if constexpr (DIMS == 3) r = C::lerp_half(r, C::lerp_half(C::lerp_half(v[4], v[5]), C::lerp_half(v[6], v[7])));
		put_elem<C::BYTES>(out, c, C::enc(r, no_double));
	}
}

// ------------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	// try_wait suspends in hardware for a bounded time; the trip counter turns a lost TMA transaction (bad descriptor)
	// into a trap instead of a hung GPU
	for (uint32_t spins = 0;; ++spins) {
		uint32_t done;
		asm volatile(
			"{\n\t"
			".reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t"
			"}"
			: "=r"(done)
			: "r"(smem_u32(bar)), "r"(parity)
			: "memory");
		if (done) return;
		if (spins > (1u << 24)) __trap();
	}
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
					 smem_u32(dst)),
				 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
				 : "memory");
}

// ------------------------------------------------------------------------------------------------------
// warp-level cascade over a dense region held in shared memory
// ------------------------------------------------------------------------------------------------------
struct Region {
	uint32_t w, h, d;    // texels of the region at level `lvl`
	uint32_t ox, oy, oz; // origin of the region in level-`lvl` texel coordinates
	uint32_t lvl;
};

template <int BPP, int DIMS>
__device__ __forceinline__ uint8_t* level_layer_ptr(const flmip_fast_params& P, uint32_t level, uint32_t layer) {
	const uint64_t lw = P.dim[0] >> level, lh = P.dim[1] >> level, ld = (DIMS == 3 ? (P.dim[2] >> level) : 1u);
	return reinterpret_cast<uint8_t*>(P.base) + P.level_off[level] + (uint64_t)layer * (lw * lh * ld * BPP);
}

// true if level `lvl + 1` exists and is not empty (levels with a zero dim hold no texels: image_types.hpp:751-766)
template <int DIMS> __device__ __forceinline__ bool next_level_has_texels(const flmip_fast_params& P, uint32_t lvl) {
	const uint32_t n = lvl + 1;
	if (n >= P.level_count) return false;
	if ((P.dim[0] >> n) == 0 || (P.dim[1] >> n) == 0) return false;
	if (DIMS == 3 && (P.dim[2] >> n) == 0) return false;
	return true;
}

// Executed by one full warp.  Reduces the region as far as it goes, writing every level to global memory.
template <uint32_t EK, int CH, int DIMS>
__device__ __forceinline__ void cascade_warp(uint8_t*& src, uint8_t*& dst, Region& R, const flmip_fast_params& P, uint32_t layer,
											 uint32_t lane) {
	constexpr int BPP = Codec<EK>::BYTES * CH;
	using IO = TexelIO<BPP>;
	while (R.lvl + 1 < P.level_count && R.w >= 2 && R.h >= 2 && (DIMS < 3 || R.d >= 2)) {
		const uint32_t dw = R.w >> 1, dh = R.h >> 1, dd = (DIMS == 3 ? R.d >> 1 : 1u);
		const uint32_t L = R.lvl + 1;
		const uint32_t LW = P.dim[0] >> L, LH = P.dim[1] >> L;
		uint8_t* gdst = level_layer_ptr<BPP, DIMS>(P, L, layer);
		const uint32_t ox = R.ox >> 1, oy = R.oy >> 1, oz = R.oz >> 1;
		const uint32_t row_pitch = R.w * BPP, slice_pitch = R.w * R.h * BPP;
		for (uint32_t i = lane; i < dw * dh * dd; i += 32) {
			const uint32_t x = i % dw, y = (i / dw) % dh, z = i / (dw * dh);
			uint32_t out[IO::NW];
			reduce_texel<EK, CH, DIMS, false>(src + (size_t)(2 * z) * slice_pitch + (size_t)(2 * y) * row_pitch + (size_t)(2 * x) * BPP,
											  row_pitch, slice_pitch, out, P.no_double);
			IO::store(dst + (size_t)i * BPP, out);
			IO::store(gdst + ((uint64_t)(oz + z) * LH * LW + (uint64_t)(oy + y) * LW + (ox + x)) * BPP, out);
		}
		__syncwarp();
		uint8_t* t = src; src = dst; dst = t;
		R.w = dw; R.h = dh; R.d = dd; R.ox = ox; R.oy = oy; R.oz = oz; R.lvl = L;
	}
}

// copies a region of global level `R.lvl` (written by other CTAs) into shared memory, bypassing L1
template <int BPP, int DIMS>
__device__ __forceinline__ void gather_region(uint8_t* smem_dst, const Region& R, const flmip_fast_params& P, uint32_t layer, uint32_t lane) {
	using IO = TexelIO<BPP>;
	const uint32_t LW = P.dim[0] >> R.lvl, LH = P.dim[1] >> R.lvl;
	const uint8_t* g = level_layer_ptr<BPP, DIMS>(P, R.lvl, layer);
	for (uint32_t i = lane; i < R.w * R.h * R.d; i += 32) {
		const uint32_t x = i % R.w, y = (i / R.w) % R.h, z = i / (R.w * R.h);
		uint32_t t[IO::NW];
		IO::template load<true>(g + ((uint64_t)(R.oz + z) * LH * LW + (uint64_t)(R.oy + y) * LW + (R.ox + x)) * BPP, t);
		IO::store(smem_dst + (size_t)i * BPP, t);
	}
	__syncwarp();
}

// classic "last block" protocol (threadfence + atomic ticket); returns true for the warp that arrives last.
// The counter is reset by that warp, so a relaunch needs no memset.
__device__ __forceinline__ bool arrive_last(uint32_t* counter, uint32_t expected, uint32_t lane) {
	__threadfence();
	__syncwarp();
	uint32_t last = 0;
	if (lane == 0) {
		const uint32_t old = atomicAdd(counter, 1u);
		last = (old == expected - 1u);
		if (last) *counter = 0u;
	}
	last = __shfl_sync(0xFFFFFFFFu, last, 0);
	if (last) __threadfence();
	return last != 0;
}

// ------------------------------------------------------------------------------------------------------
// the single-pass kernel
// ------------------------------------------------------------------------------------------------------
template <uint32_t EK, int CH, int DIMS>
__device__ __forceinline__ void fast_body(const CUtensorMap& tmap, const flmip_fast_params& P) {
	using C = Codec<EK>;
	constexpr int BPP = C::BYTES * CH;
	using TL = flmip_tiling<BPP, DIMS>;
	constexpr int TPC = 16 / BPP;                 // texels per 16-byte chunk (0 if BPP == 16 -> handled as 1 chunk = 1 texel)
	constexpr bool WIDE = (BPP == 16);            // x pair spans the two chunks of a thread
	constexpr int ROW_BYTES = TL::TILE_BYTES_X;   // bytes of one tile row in shared memory
	(void)TPC;

	extern __shared__ __align__(128) uint8_t smem_raw[]; // TMA destination: 128-byte aligned
	uint8_t* tile = smem_raw;
	uint8_t* buf_a = tile + TL::TILE_BYTES;
	uint8_t* buf_b = buf_a + TL::CASCADE_BYTES;
	__shared__ uint64_t mbar;

	const uint32_t tid = threadIdx.x, lane = tid & 31u;

	// tile coordinates
	uint32_t b = blockIdx.x;
	const uint32_t tile_x = b % P.tiles[0]; b /= P.tiles[0];
	const uint32_t tile_y = b % P.tiles[1]; b /= P.tiles[1];
	uint32_t tile_z = 0, layer = 0;
	if constexpr (DIMS == 3) { tile_z = b % P.tiles[2]; layer = b / P.tiles[2]; }
	else { layer = b; }

	if (tid == 0) {
		mbar_init(&mbar, 1);
		fence_mbar_init();
	}
	__syncthreads();
	if (tid == 0) {
		mbar_arrive_expect_tx(&mbar, TL::TILE_BYTES);
		// innermost coordinate in uint32 units; 2D images use the third tensor dim for the layer
		tma_load_3d(tile, &tmap, &mbar, (int)(tile_x * (ROW_BYTES / 4)), (int)(tile_y * TL::TY), (int)(DIMS == 3 ? tile_z * TL::TZ : layer));
	}
	mbar_wait(&mbar, 0);

	uint8_t* const g1 = level_layer_ptr<BPP, DIMS>(P, 1, layer);
	const uint64_t l1_pitch = (uint64_t)(P.dim[0] >> 1) * BPP; // bytes per level-1 row

	if constexpr (DIMS == 2) {
		// thread = 2 chunks (32 B) x 4 rows; quarter-warps read conflict-free by swapping the chunk order on lane bit 2
		const uint32_t tx = tid % TL::THREADS_X, ty = tid / TL::THREADS_X;
		const uint32_t sel = (lane >> 2) & 1u;
		uint32_t raw[4][2][4];
#pragma unroll
		for (int r = 0; r < 4; ++r) {
#pragma unroll
			for (int k = 0; k < 2; ++k) {
				const uint4 v = *reinterpret_cast<const uint4*>(tile + (4 * ty + r) * ROW_BYTES + (2 * tx + (k ^ sel)) * 16);
				raw[r][k][0] = v.x; raw[r][k][1] = v.y; raw[r][k][2] = v.z; raw[r][k][3] = v.w;
			}
		}
		uint32_t l1[2][4]; // two level-1 rows of 16 bytes, logical (left, right) order
		if constexpr (!WIDE) {
#pragma unroll
			for (int j = 0; j < 2; ++j) {
				uint32_t o0[2], o1[2];
				reduce_rows_2d<EK, CH, 4>(raw[2 * j][0], raw[2 * j + 1][0], o0, P.no_double);
				reduce_rows_2d<EK, CH, 4>(raw[2 * j][1], raw[2 * j + 1][1], o1, P.no_double);
				l1[j][0] = sel ? o1[0] : o0[0]; l1[j][1] = sel ? o1[1] : o0[1];
				l1[j][2] = sel ? o0[0] : o1[0]; l1[j][3] = sel ? o0[1] : o1[1];
			}
		} else {
#pragma unroll
			for (int j = 0; j < 2; ++j) {
				uint32_t ra[8], rb[8];
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					ra[i] = sel ? raw[2 * j][1][i] : raw[2 * j][0][i];
					ra[4 + i] = sel ? raw[2 * j][0][i] : raw[2 * j][1][i];
					rb[i] = sel ? raw[2 * j + 1][1][i] : raw[2 * j + 1][0][i];
					rb[4 + i] = sel ? raw[2 * j + 1][0][i] : raw[2 * j + 1][1][i];
				}
				reduce_rows_2d<EK, CH, 8>(ra, rb, l1[j], P.no_double);
			}
		}
		// level 1: one 16-byte store per row
#pragma unroll
		for (int j = 0; j < 2; ++j) {
			const uint32_t row = tile_y * (TL::TY / 2) + 2 * ty + j;
			*reinterpret_cast<uint4*>(g1 + (uint64_t)row * l1_pitch + (uint64_t)tile_x * (ROW_BYTES / 2) + tx * 16) =
				make_uint4(l1[j][0], l1[j][1], l1[j][2], l1[j][3]);
		}
		if constexpr (!WIDE) {
			// level 2 in registers: 8 bytes per thread
			if (P.level_count > 2) {
				uint32_t l2[2];
				reduce_rows_2d<EK, CH, 4>(l1[0], l1[1], l2, P.no_double);
				uint8_t* const g2 = level_layer_ptr<BPP, DIMS>(P, 2, layer);
				const uint64_t l2_pitch = (uint64_t)(P.dim[0] >> 2) * BPP;
				const uint32_t row = tile_y * (TL::TY / 4) + ty;
				*reinterpret_cast<uint2*>(g2 + (uint64_t)row * l2_pitch + (uint64_t)tile_x * (ROW_BYTES / 4) + tx * 8) = make_uint2(l2[0], l2[1]);
				*reinterpret_cast<uint2*>(buf_a + ty * (ROW_BYTES / 4) + tx * 8) = make_uint2(l2[0], l2[1]);
			}
		} else {
#pragma unroll
			for (int j = 0; j < 2; ++j) {
				*reinterpret_cast<uint4*>(buf_a + (2 * ty + j) * (ROW_BYTES / 2) + tx * 16) = make_uint4(l1[j][0], l1[j][1], l1[j][2], l1[j][3]);
			}
		}
	} else {
		// 3D: thread = 2 chunks x 2 rows x 2 slices -> 16 bytes of level 1
		const uint32_t tx = tid % TL::THREADS_X, ty = (tid / TL::THREADS_X) % TL::THREADS_Y, tz = tid / (TL::THREADS_X * TL::THREADS_Y);
		const uint32_t sel = (lane >> 2) & 1u;
		static_assert(TL::THREADS_X == 4, "bank-conflict swizzle assumes 4 chunk pairs per tile row");
		uint32_t raw[2][2][2][4]; // [slice][row][k]
#pragma unroll
		for (int s = 0; s < 2; ++s) {
#pragma unroll
			for (int r = 0; r < 2; ++r) {
#pragma unroll
				for (int k = 0; k < 2; ++k) {
					const uint4 v = *reinterpret_cast<const uint4*>(tile + ((2 * tz + s) * TL::TY + (2 * ty + r)) * ROW_BYTES + (2 * tx + (k ^ sel)) * 16);
					raw[s][r][k][0] = v.x; raw[s][r][k][1] = v.y; raw[s][r][k][2] = v.z; raw[s][r][k][3] = v.w;
				}
			}
		}
		uint32_t l1[4];
		if constexpr (!WIDE) {
			uint32_t o0[2], o1[2];
			reduce_rows_3d<EK, CH, 4>(raw[0][0][0], raw[0][1][0], raw[1][0][0], raw[1][1][0], o0, P.no_double);
			reduce_rows_3d<EK, CH, 4>(raw[0][0][1], raw[0][1][1], raw[1][0][1], raw[1][1][1], o1, P.no_double);
			l1[0] = sel ? o1[0] : o0[0]; l1[1] = sel ? o1[1] : o0[1];
			l1[2] = sel ? o0[0] : o1[0]; l1[3] = sel ? o0[1] : o1[1];
		} else {
			uint32_t rr[2][2][8];
#pragma unroll
			for (int s = 0; s < 2; ++s)
#pragma unroll
				for (int r = 0; r < 2; ++r)
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						rr[s][r][i] = sel ? raw[s][r][1][i] : raw[s][r][0][i];
						rr[s][r][4 + i] = sel ? raw[s][r][0][i] : raw[s][r][1][i];
					}
			reduce_rows_3d<EK, CH, 8>(rr[0][0], rr[0][1], rr[1][0], rr[1][1], l1, P.no_double);
		}
		const uint32_t row = tile_y * (TL::TY / 2) + ty, slice = tile_z * (TL::TZ / 2) + tz;
		const uint64_t l1_rows = P.dim[1] >> 1;
		*reinterpret_cast<uint4*>(g1 + ((uint64_t)slice * l1_rows + row) * l1_pitch + (uint64_t)tile_x * (ROW_BYTES / 2) + tx * 16) =
			make_uint4(l1[0], l1[1], l1[2], l1[3]);
		*reinterpret_cast<uint4*>(buf_a + (tz * (TL::TY / 2) + ty) * (ROW_BYTES / 2) + tx * 16) = make_uint4(l1[0], l1[1], l1[2], l1[3]);
	}

	__syncthreads();
	if (tid >= 32) return;

	// ---- warp 0: remaining levels of this tile ---------------------------------------------------------
	constexpr uint32_t S0 = TL::IN_REG_LEVELS; // level held in buf_a
	if (S0 >= P.level_count) return;
	Region R;
	R.lvl = S0;
	R.w = TL::TX >> S0; R.h = TL::TY >> S0; R.d = (DIMS == 3 ? TL::TZ >> S0 : 1u);
	R.ox = tile_x * R.w; R.oy = tile_y * R.h; R.oz = (DIMS == 3 ? tile_z * R.d : 0u);
	uint8_t *src = buf_a, *dst = buf_b;
	cascade_warp<EK, CH, DIMS>(src, dst, R, P, layer, lane);
	if (!next_level_has_texels<DIMS>(P, R.lvl)) return;

	// ---- group stage: the last tile of a G x G (x G) tile group reduces the group's patch -----------------
	constexpr uint32_t G = TL::GROUP;
	const uint32_t gx = tile_x / G, gy = tile_y / G, gz = (DIMS == 3 ? tile_z / G : 0u);
	const uint32_t ntx = min(G, P.tiles[0] - gx * G), nty = min(G, P.tiles[1] - gy * G), ntz = (DIMS == 3 ? min(G, P.tiles[2] - gz * G) : 1u);
	const uint32_t groups_per_layer = P.groups[0] * P.groups[1] * P.groups[2];
	uint32_t* counters = reinterpret_cast<uint32_t*>(P.counters);
	if (!arrive_last(counters + (uint64_t)layer * groups_per_layer + (gz * P.groups[1] + gy) * P.groups[0] + gx, ntx * nty * ntz, lane)) return;
	// patch = this group's part of level R.lvl (R.w/h/d are the per-tile remainders here)
	R.ox = gx * G * R.w; R.oy = gy * G * R.h; R.oz = gz * G * R.d;
	R.w *= ntx; R.h *= nty; R.d *= ntz;
	src = buf_a; dst = buf_b;
	gather_region<BPP, DIMS>(src, R, P, layer, lane);
	cascade_warp<EK, CH, DIMS>(src, dst, R, P, layer, lane);
	if (!next_level_has_texels<DIMS>(P, R.lvl)) return;

	// ---- layer stage: the last group of a layer finishes the chain -----------------------------------------
	if (!arrive_last(counters + (uint64_t)P.layers * groups_per_layer + layer, groups_per_layer, lane)) return;
	R.ox = R.oy = R.oz = 0;
	R.w = P.dim[0] >> R.lvl; R.h = P.dim[1] >> R.lvl; R.d = (DIMS == 3 ? P.dim[2] >> R.lvl : 1u);
	src = buf_a; dst = buf_b;
	gather_region<BPP, DIMS>(src, R, P, layer, lane);
	cascade_warp<EK, CH, DIMS>(src, dst, R, P, layer, lane);
}

// ------------------------------------------------------------------------------------------------------
// general path: literal replay of host_device_image::read_linear + fixed_image::read/write
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float wrap01(float v) { // const_math.hpp:859-869 with max = 1
	if (v < 0.0f) return fminf(__fadd_rn(1.0f, fmodf(v, 1.0f)), __uint_as_float(__float_as_uint(1.0f) - 1u));
	return fmodf(v, 1.0f);
}

template <uint32_t EK>
__device__ __forceinline__ void generic_texel(const flmip_generic_params& P, uint64_t idx) {
	using C = Codec<EK>;
	const uint32_t dc = P.dc, ch = P.channels, bpp = C::BYTES * ch;
	// idx -> (x, y, z, layer)
	uint32_t g[3] = { 0, 0, 0 };
	uint64_t rem = idx;
	g[0] = (uint32_t)(rem % P.dst_dim[0]); rem /= P.dst_dim[0];
	if (dc >= 2) { g[1] = (uint32_t)(rem % P.dst_dim[1]); rem /= P.dst_dim[1]; }
	if (dc >= 3) { g[2] = (uint32_t)(rem % P.dst_dim[2]); rem /= P.dst_dim[2]; }
	const uint32_t layer = (uint32_t)rem;

	float coord[3], w[3];
	int so[3];
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		coord[d] = 0.0f; w[d] = 0.0f; so[d] = 0;
		if ((uint32_t)d < dc) {
			coord[d] = __fmul_rn(__uint2float_rn(g[d] * 2u + 1u), P.inv_prev[d]);     // mip_map_minify.hpp:106
			const float scaled = __fmul_rn(wrap01(coord[d]), P.fdim[d]);             // host_image.hpp:875
			const float frac = __fsub_rn(scaled, floorf(scaled));                    // const_math.hpp:308-313
			so[d] = frac < 0.5f ? -1 : 1;
			w[d] = frac < 0.5f ? __fadd_rn(frac, 0.5f) : __fsub_rn(1.5f, frac);
		}
	}
	const uint8_t* src = reinterpret_cast<const uint8_t*>(P.base) + P.src_off + (uint64_t)layer * P.src_slice;
	const uint32_t n = 1u << dc;
	uint32_t v[8][4];
	for (uint32_t k = 0; k < n; ++k) {
		uint32_t c[3] = { 0, 0, 0 };
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			if ((uint32_t)d < dc) {
				const int off = ((k >> d) & 1u) ? 0 : so[d];
				float m = __fadd_rn(__fmul_rn(coord[d], P.fdim[d]), (float)off);      // host_image.hpp:168-172
				m = m > P.fdim_excl[d] ? P.fdim_excl[d] : (m < 0.0f ? 0.0f : m);
				c[d] = (uint32_t)__float2ll_rz(m);
			}
		}
		const uint32_t texel = (dc == 1 ? c[0] : (dc == 2 ? P.src_dim[0] * c[1] + c[0] : P.src_dim[0] * P.src_dim[1] * c[2] + P.src_dim[0] * c[1] + c[0]));
		const uint8_t* p = src + (uint64_t)texel * bpp;
		for (uint32_t i = 0; i < ch; ++i) {
			uint32_t raw;
			if constexpr (C::BYTES == 4) raw = *reinterpret_cast<const uint32_t*>(p + 4 * i);
			else if constexpr (C::BYTES == 2) raw = *reinterpret_cast<const unsigned short*>(p + 2 * i);
			else raw = p[i];
			v[k][i] = C::dec(raw);
		}
	}
	for (uint32_t d = 0; d < dc; ++d) {
		const uint32_t step = 1u << d;
		for (uint32_t k = 0; k < n; k += 2u * step)
			for (uint32_t i = 0; i < ch; ++i) v[k][i] = C::lerp_t(v[k][i], v[k + step][i], w[d]);
	}
	const uint32_t texel = (dc == 1 ? g[0] : (dc == 2 ? P.dst_dim[0] * g[1] + g[0] : P.dst_dim[0] * P.dst_dim[1] * g[2] + P.dst_dim[0] * g[1] + g[0]));
	uint8_t* q = reinterpret_cast<uint8_t*>(P.base) + P.dst_off + (uint64_t)layer * P.dst_slice + (uint64_t)texel * bpp;
	for (uint32_t i = 0; i < ch; ++i) {
		const uint32_t raw = C::enc(v[0][i], P.no_double);
		if constexpr (C::BYTES == 4) *reinterpret_cast<uint32_t*>(q + 4 * i) = raw;
		else if constexpr (C::BYTES == 2) *reinterpret_cast<unsigned short*>(q + 2 * i) = (unsigned short)raw;
		else q[i] = (uint8_t)raw;
	}
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}

} // namespace

// ------------------------------------------------------------------------------------------------------
// entry points (looked up by name through cuModuleGetFunction)
// ------------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256) flmip_generic(const __grid_constant__ flmip_generic_params P) {
	const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P.total) return;
	switch (P.elem_kind) {
		case FLMIP_EK_F32: generic_texel<FLMIP_EK_F32>(P, idx); break;
		case FLMIP_EK_F16: generic_texel<FLMIP_EK_F16>(P, idx); break;
		case FLMIP_EK_UNORM8: generic_texel<FLMIP_EK_UNORM8>(P, idx); break;
		case FLMIP_EK_SNORM8: generic_texel<FLMIP_EK_SNORM8>(P, idx); break;
		case FLMIP_EK_UNORM16: generic_texel<FLMIP_EK_UNORM16>(P, idx); break;
		case FLMIP_EK_SNORM16: generic_texel<FLMIP_EK_SNORM16>(P, idx); break;
		case FLMIP_EK_U8: generic_texel<FLMIP_EK_U8>(P, idx); break;
		case FLMIP_EK_I8: generic_texel<FLMIP_EK_I8>(P, idx); break;
		case FLMIP_EK_U16: generic_texel<FLMIP_EK_U16>(P, idx); break;
		case FLMIP_EK_I16: generic_texel<FLMIP_EK_I16>(P, idx); break;
		case FLMIP_EK_U32: generic_texel<FLMIP_EK_U32>(P, idx); break;
		case FLMIP_EK_I32: generic_texel<FLMIP_EK_I32>(P, idx); break;
		default: break;
	}
}

// counter-based pattern of SURVEY.md 8d (the CPU checker defines the same function)
extern "C" __global__ void __launch_bounds__(256) flmip_fill(const __grid_constant__ flmip_fill_params P) {
	const uint64_t total = P.elems_per_layer * P.layers;
	const uint32_t ek = P.elem_kind;
	const int bytes = flmip_elem_bytes(ek);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t layer = i / P.elems_per_layer, e = i % P.elems_per_layer;
		const uint64_t r = splitmix64(splitmix64(splitmix64(0x9E3779B97F4A7C15ull + P.config_id) + (P.layer_id0 + layer)) + e);
		uint32_t v;
		if (ek == FLMIP_EK_F16) {
			v = (uint32_t)(r >> 63) << 15 | (1u + (uint32_t)((r >> 32) % 19u)) << 10 | (uint32_t)(r & 0x3FFu);
		} else if (ek == FLMIP_EK_F32) {
			v = __float_as_uint(__fmul_rn(__ull2float_rn(r >> 40), 0x1p-24f));
		} else if (ek == FLMIP_EK_I32) {
			v = (uint32_t)((int)(uint32_t)r >> 1);
		} else {
			v = (uint32_t)r;
		}
		uint8_t* p = reinterpret_cast<uint8_t*>(P.dst);
		if (bytes == 4) reinterpret_cast<uint32_t*>(p)[i] = v;
		else if (bytes == 2) reinterpret_cast<unsigned short*>(p)[i] = (unsigned short)v;
		else p[i] = (uint8_t)v;
	}
}

#define FLMIP_FAST_KERNEL(D, K, CHN)                                                                                            \
	extern "C" __global__ void __launch_bounds__(256) flmip_fast##D##d_k##K##_c##CHN(const __grid_constant__ CUtensorMap tmap,    \
																					   const __grid_constant__ flmip_fast_params P) { \
		fast_body<K, CHN, D>(tmap, P);                                                                                          \
	}
#define FLMIP_FAST_KERNELS_FOR_KIND(K) \
	FLMIP_FAST_KERNEL(2, K, 1) FLMIP_FAST_KERNEL(2, K, 2) FLMIP_FAST_KERNEL(2, K, 4) FLMIP_FAST_KERNEL(3, K, 1) FLMIP_FAST_KERNEL(3, K, 2) FLMIP_FAST_KERNEL(3, K, 4)

FLMIP_FAST_KERNELS_FOR_KIND(0)
FLMIP_FAST_KERNELS_FOR_KIND(1)
FLMIP_FAST_KERNELS_FOR_KIND(2)
FLMIP_FAST_KERNELS_FOR_KIND(3)
FLMIP_FAST_KERNELS_FOR_KIND(4)
FLMIP_FAST_KERNELS_FOR_KIND(5)
FLMIP_FAST_KERNELS_FOR_KIND(6)
FLMIP_FAST_KERNELS_FOR_KIND(7)
FLMIP_FAST_KERNELS_FOR_KIND(8)
FLMIP_FAST_KERNELS_FOR_KIND(9)
FLMIP_FAST_KERNELS_FOR_KIND(10)
FLMIP_FAST_KERNELS_FOR_KIND(11)
