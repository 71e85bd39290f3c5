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



// ------------------------------------------------------------------------------------------------------
// packed fp32 pairs: sm_100a issues two IEEE fp32 operations per instruction (FFMA2 / FADD2 / FMUL2, same
// per-lane rounding as the scalar forms), which halves the issue slots of the float formats' arithmetic
// ------------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(uint32_t lo, uint32_t hi) {
	f32x2_t r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
	return r;
}
__device__ __forceinline__ void unpk2(f32x2_t v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t splat2(float f) { return pk2(__float_as_uint(f), __float_as_uint(f)); }
__device__ __forceinline__ f32x2_t fma2_rn(f32x2_t a, f32x2_t b, f32x2_t c) {
	f32x2_t r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
	return r;
}
__device__ __forceinline__ f32x2_t add2_rn(f32x2_t a, f32x2_t b) {
	f32x2_t r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f32x2_t add2_rz(f32x2_t a, f32x2_t b) {
	f32x2_t r;
	asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f32x2_t sub2_rn(f32x2_t a, f32x2_t b) {
	f32x2_t r;
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f32x2_t mul2_rn(f32x2_t a, f32x2_t b) {
	f32x2_t r;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f32x2_t mul2_rz(f32x2_t a, f32x2_t b) {
	f32x2_t r;
	asm("mul.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}

// RN(a * b) that ptxas cannot merge with a following add: it contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (unlike
// the scalar mul.rn / add.rn, which are never fused; --fmad=false does not stop it, and it folds a literal -0 addend), which
// would round once where the reference rounds product and sum separately.  fma(a, b, -0) == RN(a * b) bit for bit (a -0
// addend keeps the sign of a zero product); the -0 comes from a special register ptxas cannot reason about.
__device__ __forceinline__ uint32_t opaque_neg_zero() {
	uint32_t r;
	asm("{\n\t.reg .u32 t;\n\tmov.u32 t, %%nsmid;\n\tshr.u32 t, t, 31;\n\tor.b32 %0, t, 0x80000000;\n\t}" : "=r"(r));
	return r;
}
__device__ __forceinline__ f32x2_t mul2_sep(f32x2_t a, f32x2_t b) {
	const uint32_t nz = opaque_neg_zero();
	return fma2_rn(a, b, pk2(nz, nz));
}

// ------------------------------------------------------------------------------------------------------
// element codecs
// ------------------------------------------------------------------------------------------------------
template <uint32_t EK> struct Codec {
	static constexpr bool IS_INT = (EK >= FLMIP_EK_U8);
	static constexpr bool IS_SIGNED_INT = (EK == FLMIP_EK_I8 || EK == FLMIP_EK_I16 || EK == FLMIP_EK_I32);
	static constexpr int BYTES = flmip_elem_bytes(EK);
	static constexpr uint32_t MASK = (BYTES == 4 ? 0xFFFFFFFFu : (BYTES == 2 ? 0xFFFFu : 0xFFu));

	// Conversions that stay off the quarter-rate XU pipe (I2F / F2I / F2F): an unsigned integer u < 2^23 placed in the
	// mantissa of 2^23 is the float 2^23 + u, so  float(u) * c == fma(2^23 + u, c, -(2^23 * c))  exactly (2^23 * c is a
	// power-of-two scaling of c, the fused product-sum is rounded once), and for 0 <= t < 2^23 the low mantissa bits
	// of  t + 2^23  rounded toward zero are trunc(t).
	static constexpr float MAGIC = 8388608.0f; // 2^23 == 0x4B000000
	static constexpr float UNORM_C = (EK == FLMIP_EK_UNORM8 ? (float)(1.0 / 255.0) : (float)(1.0 / 65535.0));
	static constexpr float UNORM_S = (EK == FLMIP_EK_UNORM8 ? 255.0f : 65535.0f);

	// decode zero-extended storage bits into the compute domain (fp32 bits or widened 32-bit integer)
	// host_image.hpp:487-561 (float / normalized), :640-667 (int / uint)
	static __device__ __forceinline__ uint32_t dec(uint32_t raw) {
		if constexpr (EK == FLMIP_EK_F32 || EK == FLMIP_EK_U32 || EK == FLMIP_EK_I32 || EK == FLMIP_EK_U8 || EK == FLMIP_EK_U16) {
			return raw;
		} else if constexpr (EK == FLMIP_EK_F16) {
			return __float_as_uint(__half2float(__ushort_as_half((unsigned short)raw)));
		} else if constexpr (EK == FLMIP_EK_UNORM8 || EK == FLMIP_EK_UNORM16) {
			// == __fmul_rn(__uint2float_rn(raw), UNORM_C)
			return __float_as_uint(__fmaf_rn(__uint_as_float(0x4B000000u | raw), UNORM_C, -(MAGIC * UNORM_C)));
		} else if constexpr (EK == FLMIP_EK_SNORM8) {
			return __float_as_uint(__fmul_rn(__int2float_rn((int)(signed char)raw), (float)(1.0 / 127.0)));
		} else if constexpr (EK == FLMIP_EK_SNORM16) {
			return __float_as_uint(__fmul_rn(__int2float_rn((int)(short)raw), (float)(1.0 / 32767.0)));
		} else if constexpr (EK == FLMIP_EK_I8) {
			return (uint32_t)(int)(signed char)raw;
		} else { // I16
			return (uint32_t)(int)(short)raw;
		}
	}

	// 0x4B000000 held in a register the compiler cannot see through: PRMT has one immediate slot, and it must go to
	// the (compile-time) selector -- with the constant as the immediate every PRMT needs a MOV for its selector.
	// (ptxas folds a constant mov, so the value is derived from a special register it cannot reason about.)
	static __device__ __forceinline__ uint32_t magic_bits() {
		uint32_t r;
		asm("{\n\t.reg .u32 t;\n\tmov.u32 t, %%nsmid;\n\tshr.u32 t, t, 31;\n\tor.b32 %0, t, 0x4B000000;\n\t}" : "=r"(r));
		return r;
	}

	// decode element i of a row of packed 32-bit words
	static __device__ __forceinline__ uint32_t dec_at(const uint32_t* w, int i) {
		if constexpr (EK == FLMIP_EK_UNORM8) {
			// one PRMT builds 0x4B0000uu
			return __float_as_uint(__fmaf_rn(__uint_as_float(__byte_perm(w[i >> 2], magic_bits(), 0x7440u | (uint32_t)(i & 3))), UNORM_C,
											 -(MAGIC * UNORM_C)));
		} else if constexpr (EK == FLMIP_EK_UNORM16) {
			return __float_as_uint(
				__fmaf_rn(__uint_as_float(__byte_perm(w[i >> 1], magic_bits(), (i & 1) ? 0x7432u : 0x7410u)), UNORM_C, -(MAGIC * UNORM_C)));
		} else if constexpr (EK == FLMIP_EK_F16) {
			const __half2 h = *reinterpret_cast<const __half2*>(&w[i >> 1]);
			return __float_as_uint((i & 1) ? __high2float(h) : __low2float(h));
		} else {
			return dec(get_elem<BYTES>(w, i));
		}
	}

	// encode back to storage bits: host_image.hpp:672-722 + insert_channels :391-460 (truncating), :801-825
	static __device__ __forceinline__ uint32_t enc(uint32_t v, uint32_t no_double) { return enc_dirty(v, no_double) & MASK; }

	// same, but bits above the storage width are unspecified (the packers select bytes with PRMT anyway)
	static __device__ __forceinline__ uint32_t enc_dirty(uint32_t v, uint32_t no_double) {
		if constexpr (EK == FLMIP_EK_F32 || EK == FLMIP_EK_U32 || EK == FLMIP_EK_I32) {
			return v;
		} else if constexpr (IS_INT) {
			return v;
		} else if constexpr (EK == FLMIP_EK_F16) {
			return (uint32_t)__half_as_ushort(__float2half_rn(__uint_as_float(v)));
		} else if constexpr (EK == FLMIP_EK_UNORM8) {
			// v in [0, 1] -> t in [0, 255]: trunc(t) sits in the low mantissa byte of t + 2^23 (RZ)
			return __float_as_uint(__fadd_rz(__fmul_rn(__uint_as_float(v), 255.0f), MAGIC));
		} else if constexpr (EK == FLMIP_EK_SNORM8) {
			return (uint32_t)__float2int_rz(__fmul_rn(__uint_as_float(v), 127.0f));
		} else {
			// 9..16 bit normalized: fp_scale_type is double unless FLOOR_DEVICE_NO_DOUBLE (host_image.hpp:398-402).
			// double(f) * scale is exact (24 + 16 significant bits), so the reference computes trunc(f * scale) of the
			// EXACT product; the fp32 product rounded toward zero has the same integer part (an integer n <= exact
			// is itself a float, hence <= RZ(exact)).  The all-float variant truncates the RN product instead.
			constexpr float scale = (EK == FLMIP_EK_UNORM16 ? 65535.0f : 32767.0f);
			const float f = __uint_as_float(v);
			const float t = no_double ? __fmul_rn(f, scale) : __fmul_rz(f, scale);
			if constexpr (EK == FLMIP_EK_UNORM16) return __float_as_uint(__fadd_rz(t, MAGIC));
			else return (uint32_t)__float2int_rz(t);
		}
	}

	// encode N consecutive elements into packed words (N * BYTES is a multiple of 4)
	template <int N> static __device__ __forceinline__ void enc_pack(const uint32_t (&v)[N], uint32_t* out, uint32_t no_double) {
		if constexpr (BYTES == 4) {
#pragma unroll
			for (int i = 0; i < N; ++i) out[i] = enc(v[i], no_double);
		} else if constexpr (BYTES == 2) {
#pragma unroll
			for (int i = 0; i < N; i += 2) {
				if constexpr (EK == FLMIP_EK_F16) {
					// one F2FP.PACK_AB converts both halves
					const __half2 h = __floats2half2_rn(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
					out[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
				} else {
					out[i >> 1] = __byte_perm(enc_dirty(v[i], no_double), enc_dirty(v[i + 1], no_double), 0x5410u);
				}
			}
		} else {
#pragma unroll
			for (int i = 0; i < N; i += 4) {
				const uint32_t lo = __byte_perm(enc_dirty(v[i], no_double), enc_dirty(v[i + 1], no_double), 0x0040u);
				const uint32_t hi = __byte_perm(enc_dirty(v[i + 2], no_double), enc_dirty(v[i + 3], no_double), 0x0040u);
				out[i >> 2] = __byte_perm(lo, hi, 0x5410u);
			}
		}
	}

	// ---- the same codecs on pairs of elements (float formats only) -----------------------------------------
	static __device__ __forceinline__ f32x2_t dec2_at(const uint32_t* w, int i0, int i1) {
		if constexpr (EK == FLMIP_EK_UNORM8) {
			const uint32_t m = magic_bits();
			return fma2_rn(pk2(__byte_perm(w[i0 >> 2], m, 0x7440u | (uint32_t)(i0 & 3)), __byte_perm(w[i1 >> 2], m, 0x7440u | (uint32_t)(i1 & 3))),
						   splat2(UNORM_C), splat2(-(MAGIC * UNORM_C)));
		} else if constexpr (EK == FLMIP_EK_UNORM16) {
			const uint32_t m = magic_bits();
			return fma2_rn(pk2(__byte_perm(w[i0 >> 1], m, (i0 & 1) ? 0x7432u : 0x7410u), __byte_perm(w[i1 >> 1], m, (i1 & 1) ? 0x7432u : 0x7410u)),
						   splat2(UNORM_C), splat2(-(MAGIC * UNORM_C)));
		} else {
			return pk2(dec_at(w, i0), dec_at(w, i1));
		}
	}
	static __device__ __forceinline__ f32x2_t lerp_half2(f32x2_t a, f32x2_t b) {
		static_assert(!IS_INT, "float formats only");
		// fp32 storage: product and sum are rounded separately ((b - a) * 0.5 can be an inexact subnormal), see mul2_sep()
		if constexpr (EK == FLMIP_EK_F32) return add2_rn(mul2_sep(sub2_rn(b, a), splat2(0.5f)), a);
		else return fma2_rn(sub2_rn(b, a), splat2(0.5f), a);
	}
	// unorm8 / unorm16: the encoder leaves the stored value q in the mantissa of the float 2^23 + q (bits 0x4B000000 | q) --
	// exactly what the decoder builds with a PRMT before its FMA.  A level computed from a level this thread has just encoded
	// can therefore decode these "magic" pairs directly and skip the byte extraction (dec2_magic(enc2_magic(v)) ==
	// dec2_at(packed enc(v))).
	static constexpr bool HAS_MAGIC = (EK == FLMIP_EK_UNORM8 || EK == FLMIP_EK_UNORM16);
	static __device__ __forceinline__ f32x2_t enc2_magic(f32x2_t v, uint32_t no_double) {
		static_assert(HAS_MAGIC, "unorm8 / unorm16 only");
		if constexpr (EK == FLMIP_EK_UNORM8) {
			return add2_rz(mul2_rn(v, splat2(255.0f)), splat2(MAGIC));
		} else {
			const f32x2_t t = no_double ? mul2_rn(v, splat2(65535.0f)) : mul2_rz(v, splat2(65535.0f));
			return add2_rz(t, splat2(MAGIC));
		}
	}
	static __device__ __forceinline__ f32x2_t dec2_magic(f32x2_t m) {
		static_assert(HAS_MAGIC, "unorm8 / unorm16 only");
		return fma2_rn(m, splat2(UNORM_C), splat2(-(MAGIC * UNORM_C)));
	}
	// storage bits of two elements; bits above the storage width are unspecified
	static __device__ __forceinline__ void enc2_dirty(f32x2_t v, uint32_t& lo, uint32_t& hi, uint32_t no_double) {
		if constexpr (HAS_MAGIC) {
			unpk2(enc2_magic(v, no_double), lo, hi);
		} else {
			uint32_t a, b;
			unpk2(v, a, b);
			lo = enc_dirty(a, no_double);
			hi = enc_dirty(b, no_double);
		}
	}
	// encode NP pairs into packed words
	template <int NP> static __device__ __forceinline__ void enc_pack2(const f32x2_t (&v)[NP], uint32_t* out, uint32_t no_double) {
		if constexpr (BYTES == 4) {
#pragma unroll
			for (int p = 0; p < NP; ++p) enc2_dirty(v[p], out[2 * p], out[2 * p + 1], no_double);
		} else if constexpr (BYTES == 2) {
#pragma unroll
			for (int p = 0; p < NP; ++p) {
				uint32_t a, b;
				if constexpr (EK == FLMIP_EK_F16) {
					unpk2(v[p], a, b);
					const __half2 h = __floats2half2_rn(__uint_as_float(a), __uint_as_float(b));
					out[p] = *reinterpret_cast<const uint32_t*>(&h);
				} else {
					enc2_dirty(v[p], a, b, no_double);
					out[p] = __byte_perm(a, b, 0x5410u);
				}
			}
		} else {
#pragma unroll
			for (int p = 0; p < NP; p += 2) {
				uint32_t a, b, c, d;
				enc2_dirty(v[p], a, b, no_double);
				enc2_dirty(v[p + 1], c, d, no_double);
				out[p >> 1] = __byte_perm(__byte_perm(a, b, 0x0040u), __byte_perm(c, d, 0x0040u), 0x5410u);
			}
		}
	}

	// pack NP already encoded "magic" pairs (unorm8 / unorm16) into words
	template <int NP> static __device__ __forceinline__ void pack2_magic(const f32x2_t (&m)[NP], uint32_t* out) {
		static_assert(HAS_MAGIC && BYTES <= 2, "unorm8 / unorm16 only");
		if constexpr (BYTES == 2) {
#pragma unroll
			for (int p = 0; p < NP; ++p) {
				uint32_t a, b;
				unpk2(m[p], a, b);
				out[p] = __byte_perm(a, b, 0x5410u);
			}
		} else {
#pragma unroll
			for (int p = 0; p < NP; p += 2) {
				uint32_t a, b, c, d;
				unpk2(m[p], a, b);
				unpk2(m[p + 1], c, d);
				out[p >> 1] = __byte_perm(__byte_perm(a, b, 0x0040u), __byte_perm(c, d, 0x0040u), 0x5410u);
			}
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


// texel load / store by size
template <int BPP> struct TexelIO {
	static constexpr int NW = (BPP + 3) / 4;
	template <bool CG> static __device__ __forceinline__ void load(const uint8_t* p, uint32_t (&w)[NW]) {
		if constexpr (BPP == 3) { // 3-channel texels (LDG tile kernel only): element-sized accesses, packed like the other sizes
			w[0] = (uint32_t)(CG ? __ldcg(p) : p[0]) | (uint32_t)(CG ? __ldcg(p + 1) : p[1]) << 8 | (uint32_t)(CG ? __ldcg(p + 2) : p[2]) << 16;
		} else if constexpr (BPP == 6) {
			const unsigned short* h = (const unsigned short*)p;
			w[0] = (uint32_t)(CG ? __ldcg(h) : h[0]) | (uint32_t)(CG ? __ldcg(h + 1) : h[1]) << 16;
			w[1] = (uint32_t)(CG ? __ldcg(h + 2) : h[2]);
		} else if constexpr (BPP == 12) {
			const uint32_t* q = (const uint32_t*)p;
			w[0] = CG ? __ldcg(q) : q[0]; w[1] = CG ? __ldcg(q + 1) : q[1]; w[2] = CG ? __ldcg(q + 2) : q[2];
		} else
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
		if constexpr (BPP == 3) { p[0] = (uint8_t)w[0]; p[1] = (uint8_t)(w[0] >> 8); p[2] = (uint8_t)(w[0] >> 16); }
		else if constexpr (BPP == 6) { unsigned short* h = (unsigned short*)p; h[0] = (unsigned short)w[0]; h[1] = (unsigned short)(w[0] >> 16); h[2] = (unsigned short)w[1]; }
		else if constexpr (BPP == 12) { uint32_t* q = (uint32_t*)p; q[0] = w[0]; q[1] = w[1]; q[2] = w[2]; }
		else
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
	constexpr int NO = NW * 2 / C::BYTES; // output elements (= half the elements of one source row)
	static_assert(NO >= CH && (NO % CH) == 0, "row must hold at least one x pair");
	if constexpr (!C::IS_INT) {
		static_assert((NO % 2) == 0, "pairs of output elements");
		f32x2_t v[NO / 2];
#pragma unroll
		for (int p = 0; p < NO / 2; ++p) {
			const int e0 = 2 * p, e1 = 2 * p + 1;
			const int ea0 = (2 * (e0 / CH)) * CH + (e0 % CH), eb0 = ea0 + CH, ea1 = (2 * (e1 / CH)) * CH + (e1 % CH), eb1 = ea1 + CH;
			const f32x2_t x0 = C::lerp_half2(C::dec2_at(r0, ea0, ea1), C::dec2_at(r0, eb0, eb1));
			const f32x2_t x1 = C::lerp_half2(C::dec2_at(r1, ea0, ea1), C::dec2_at(r1, eb0, eb1));
			v[p] = C::lerp_half2(x0, x1);
		}
		C::template enc_pack2<NO / 2>(v, out, no_double);
	} else {
		uint32_t v[NO];
#pragma unroll
		for (int e = 0; e < NO; ++e) {
			const int ea = (2 * (e / CH)) * CH + (e % CH), eb = ea + CH;
			const uint32_t x0 = C::lerp_half(C::dec_at(r0, ea), C::dec_at(r0, eb));
			const uint32_t x1 = C::lerp_half(C::dec_at(r1, ea), C::dec_at(r1, eb));
			v[e] = C::lerp_half(x0, x1);
		}
		C::template enc_pack<NO>(v, out, no_double);
	}
}

// unorm8 / unorm16: the same, also handing out the encoded result as "magic" pairs (see Codec::enc2_magic) ...
template <uint32_t EK, int CH, int NW>
__device__ __forceinline__ void reduce_rows_2d_keep(const uint32_t (&r0)[NW], const uint32_t (&r1)[NW], uint32_t (&out)[NW / 2],
													f32x2_t (&kept)[NW / Codec<EK>::BYTES], uint32_t no_double) {
	using C = Codec<EK>;
	constexpr int NO = NW * 2 / C::BYTES;
	static_assert(NO >= CH && (NO % CH) == 0 && (NO % 2) == 0, "row must hold at least one x pair");
#pragma unroll
	for (int p = 0; p < NO / 2; ++p) {
		const int e0 = 2 * p, e1 = 2 * p + 1;
		const int ea0 = (2 * (e0 / CH)) * CH + (e0 % CH), eb0 = ea0 + CH, ea1 = (2 * (e1 / CH)) * CH + (e1 % CH), eb1 = ea1 + CH;
		const f32x2_t x0 = C::lerp_half2(C::dec2_at(r0, ea0, ea1), C::dec2_at(r0, eb0, eb1));
		const f32x2_t x1 = C::lerp_half2(C::dec2_at(r1, ea0, ea1), C::dec2_at(r1, eb0, eb1));
		kept[p] = C::enc2_magic(C::lerp_half2(x0, x1), no_double);
	}
	C::template pack2_magic<NO / 2>(kept, out);
}
// ... and the next level from two rows of such pairs (NP pairs = NP * 2 / CH texels per row, CH >= 2: a pair never spans texels)
template <uint32_t EK, int CH, int NP>
__device__ __forceinline__ void reduce_magic_2d(const f32x2_t (&m0)[NP], const f32x2_t (&m1)[NP], uint32_t (&out)[NP * Codec<EK>::BYTES / 4],
												uint32_t no_double) {
	using C = Codec<EK>;
	constexpr int PPT = CH / 2; // pairs per texel
	static_assert(CH >= 2 && (NP % (2 * PPT)) == 0, "whole x pairs of texels");
	f32x2_t v[NP / 2];
#pragma unroll
	for (int p = 0; p < NP / 2; ++p) {
		const int a = (2 * (p / PPT)) * PPT + (p % PPT), b = a + PPT;
		const f32x2_t x0 = C::lerp_half2(C::dec2_magic(m0[a]), C::dec2_magic(m0[b]));
		const f32x2_t x1 = C::lerp_half2(C::dec2_magic(m1[a]), C::dec2_magic(m1[b]));
		v[p] = C::enc2_magic(C::lerp_half2(x0, x1), no_double);
	}
	C::template pack2_magic<NP / 2>(v, out);
}

// r[z][y]: rows (y, y+1) of slices (z, z+1)
template <uint32_t EK, int CH, int NW>
__device__ __forceinline__ void reduce_rows_3d(const uint32_t (&r00)[NW], const uint32_t (&r01)[NW], const uint32_t (&r10)[NW],
											   const uint32_t (&r11)[NW], uint32_t (&out)[NW / 2], uint32_t no_double) {
	using C = Codec<EK>;
	constexpr int NO = NW * 2 / C::BYTES;
	static_assert(NO >= CH && (NO % CH) == 0, "row must hold at least one x pair");
	if constexpr (!C::IS_INT) {
		static_assert((NO % 2) == 0, "pairs of output elements");
		f32x2_t v[NO / 2];
#pragma unroll
		for (int p = 0; p < NO / 2; ++p) {
			const int e0 = 2 * p, e1 = 2 * p + 1;
			const int ea0 = (2 * (e0 / CH)) * CH + (e0 % CH), eb0 = ea0 + CH, ea1 = (2 * (e1 / CH)) * CH + (e1 % CH), eb1 = ea1 + CH;
			const f32x2_t x00 = C::lerp_half2(C::dec2_at(r00, ea0, ea1), C::dec2_at(r00, eb0, eb1));
			const f32x2_t x01 = C::lerp_half2(C::dec2_at(r01, ea0, ea1), C::dec2_at(r01, eb0, eb1));
			const f32x2_t x10 = C::lerp_half2(C::dec2_at(r10, ea0, ea1), C::dec2_at(r10, eb0, eb1));
			const f32x2_t x11 = C::lerp_half2(C::dec2_at(r11, ea0, ea1), C::dec2_at(r11, eb0, eb1));
			v[p] = C::lerp_half2(C::lerp_half2(x00, x01), C::lerp_half2(x10, x11));
		}
		C::template enc_pack2<NO / 2>(v, out, no_double);
	} else {
		uint32_t v[NO];
#pragma unroll
		for (int e = 0; e < NO; ++e) {
			const int ea = (2 * (e / CH)) * CH + (e % CH), eb = ea + CH;
			const uint32_t x00 = C::lerp_half(C::dec_at(r00, ea), C::dec_at(r00, eb));
			const uint32_t x01 = C::lerp_half(C::dec_at(r01, ea), C::dec_at(r01, eb));
			const uint32_t x10 = C::lerp_half(C::dec_at(r10, ea), C::dec_at(r10, eb));
			const uint32_t x11 = C::lerp_half(C::dec_at(r11, ea), C::dec_at(r11, eb));
			v[e] = C::lerp_half(C::lerp_half(x00, x01), C::lerp_half(x10, x11));
		}
		C::template enc_pack<NO>(v, out, no_double);
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
		for (int k = 0; k < (DIMS == 3 ? 8 : 4); ++k) v[k] = C::dec_at(t[k], c);
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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { // raises the transaction count only; the arrival follows
	asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	// try_wait suspends the thread in hardware until the phase completes or a time limit expires.  Without a hint that limit
	// is short and the waiting warps (the producer, four mostly idle finisher warps per CTA) re-issue try_wait + branch all
	// the time: 19 % of all issued instructions of the RGBA8 kernel (ncu source view).  With the suspend-time hint (the value
	// CUTLASS's ClusterBarrier::wait uses) a waiter costs a handful of issue slots per tile; wake-up is still driven by the
	// phase completion.  Measured effect: C3 +0.7 %, others unchanged -- the spinners mostly used slots the consumers left
	// free.  The trip counter turns a lost TMA transaction (bad descriptor) into a trap instead of a hung GPU.
	for (uint32_t spins = 0;; ++spins) {
		uint32_t done;
		asm volatile(
			"{\n\t"
			".reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t"
			"}"
			: "=r"(done)
			: "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
			: "memory");
		if (done) return;
		if (spins > (1u << 20)) __trap();
	}
}
// Tuning alternative (-DFLMIP_FINISHER_SLEEP_WAIT) for waits that last about as long as a whole tile (the finisher pool waiting for its
// next cascade slot): the hardware-suspended try_wait above comes back every few dozen nanoseconds (ncu: ~100 trips of 6 instructions per
// wait, 8 % of all issued instructions of the persistent tile kernel); polling with an explicit sleep removes those instructions.
// Measured: no gain (N2 0.3219 vs 0.3218 ms, C3 slightly worse) -- the issue slots were not what the consumers lacked; not the default.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
	for (uint32_t spins = 0;; ++spins) {
		uint32_t done;
		asm volatile(
			"{\n\t"
			".reg .pred p;\n\t"
			"mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t"
			"}"
			: "=r"(done)
			: "r"(smem_u32(bar)), "r"(parity)
			: "memory");
		if (done) return;
		__nanosleep(128);
		if (spins > (1u << 24)) __trap();
	}
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
					 smem_u32(dst)),
				 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
				 : "memory");
}

// programmatic dependent launch (the host launches flmip_fast* / flmip_tile* with the stream-serialization attribute): everything
// before pdl_wait() may overlap the tail of the previous kernel in the stream; nothing in global memory may be touched before it.
// The dependents are released right after, so that at most one successor is ever pending.  No-ops for a plain launch.
__device__ __forceinline__ void pdl_wait_then_release() {
	asm volatile("griddepcontrol.wait;" ::: "memory");
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// Chains of independent images (`late_wait`, set per launch by the host on queues that have opted in: flmip_stream_set_chain_overlap).
//   0: wait for the kernel in front, then release the dependents (above).
//   1: the kernel in front is a chain on ANOTHER image (none of the kernels that can still be running touches this image): nothing this
//      kernel reads or writes depends on it, so it starts streaming while that kernel's tail (last units, group / layer stages: 7 - 10 us
//      without memory traffic) is still running, and its own dependents may follow as soon as its CTAs are resident.  Stream order for
//      whatever comes after is kept by waiting for the predecessor at the END instead (pdl_late_wait, one thread of one CTA: a grid is
//      complete when its last CTA is), so a grid never completes before the grid in front of it has completed and flushed.
//   2: a later kernel of a multi-kernel chain on such a queue: it needs its predecessor's output, so it waits at its start -- but it
//      releases its dependents first, so that the first kernel of the NEXT image's chain need not wait for this (small) kernel either.
__device__ __forceinline__ void pdl_start(uint32_t late_wait) {
	if (late_wait == 0u) {
		pdl_wait_then_release();
	} else {
		asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
		if (late_wait == 2u) asm volatile("griddepcontrol.wait;" ::: "memory");
	}
}
__device__ __forceinline__ void pdl_late_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#ifdef FLMIP_TIMELINE
// tuning builds only (make DEFS=-DFLMIP_TIMELINE): per-CTA time stamps behind the scheduler words, read back by flmip_debug_timeline
__device__ __forceinline__ uint64_t gtime() {
	uint64_t t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
#define FLMIP_STAMP(P, slot) (reinterpret_cast<unsigned long long*>(((P).sched + 15ull) & ~7ull)[blockIdx.x * 16u + (slot)] = gtime())
#define FLMIP_STAMP_MAX(P, slot) atomicMax(&reinterpret_cast<unsigned long long*>(((P).sched + 15ull) & ~7ull)[blockIdx.x * 16u + (slot)], (unsigned long long)gtime())
#else
#define FLMIP_STAMP(P, slot) ((void)0)
#define FLMIP_STAMP_MAX(P, slot) ((void)0)
#endif

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// All region sizes are powers of two, so texel indices decompose with shifts.
template <uint32_t EK, int CH, int DIMS>
__device__ __forceinline__ void cascade_warp(uint8_t*& src, uint8_t*& dst, Region& R, const flmip_fast_params& P, uint32_t layer,
											 uint32_t lane, uint32_t stamp_slot = 0u) {
	constexpr int BPP = Codec<EK>::BYTES * CH;
	using IO = TexelIO<BPP>;
	while (R.lvl + 1 < P.level_count && R.w >= 2 && R.h >= 2 && (DIMS < 3 || R.d >= 2)) {
		const uint32_t dw = R.w >> 1, dh = R.h >> 1, dd = (DIMS == 3 ? R.d >> 1 : 1u);
		const uint32_t sw = 31u - __clz(dw), sh = 31u - __clz(dh);
		const uint32_t L = R.lvl + 1;
		const uint32_t LW = P.dim[0] >> L, LH = P.dim[1] >> L;
		uint8_t* gdst = level_layer_ptr<BPP, DIMS>(P, L, layer);
		const uint32_t ox = R.ox >> 1, oy = R.oy >> 1, oz = R.oz >> 1;
		const uint32_t row_pitch = R.w * BPP, slice_pitch = R.w * R.h * BPP;
		for (uint32_t i = lane; i < dw * dh * dd; i += 32) {
			const uint32_t x = i & (dw - 1u), y = (i >> sw) & (dh - 1u), z = i >> (sw + sh);
			uint32_t out[IO::NW];
			reduce_texel<EK, CH, DIMS, false>(src + (size_t)(2 * z) * slice_pitch + (size_t)(2 * y) * row_pitch + (size_t)(2 * x) * BPP,
											  row_pitch, slice_pitch, out, P.no_double);
			IO::store(dst + (size_t)i * BPP, out);
			IO::store(gdst + ((uint64_t)(oz + z) * LH * LW + (uint64_t)(oy + y) * LW + (ox + x)) * BPP, out);
		}
		__syncwarp();
		uint8_t* t = src; src = dst; dst = t;
		R.w = dw; R.h = dh; R.d = dd; R.ox = ox; R.oy = oy; R.oz = oz; R.lvl = L;
#ifdef FLMIP_TIMELINE
		if (stamp_slot && stamp_slot < 16u && lane == 0) FLMIP_STAMP_MAX(P, stamp_slot++); // tuning builds: one stamp per level
#endif
	}
}

// copies a region of global level `R.lvl` (written by other CTAs) into shared memory, bypassing L1.
// Loads are issued in batches per lane: under load a round trip to L2 costs microseconds.
template <int BPP, int DIMS>
__device__ __forceinline__ void gather_region(uint8_t* smem_dst, const Region& R, const flmip_fast_params& P, uint32_t layer, uint32_t lane) {
	using IO = TexelIO<BPP>;
	constexpr uint32_t U = (IO::NW >= 4 ? 2u : (IO::NW == 2 ? 4u : 8u)); // <= 8 registers of payload in flight
	const uint32_t LW = P.dim[0] >> R.lvl, LH = P.dim[1] >> R.lvl;
	const uint32_t sw = 31u - __clz(R.w), sh = 31u - __clz(R.h);
	const uint8_t* g = level_layer_ptr<BPP, DIMS>(P, R.lvl, layer);
	const uint32_t n = R.w * R.h * R.d;
	{
		// rows of the region that are whole, aligned 16-byte vectors (the group patches of every benchmarked shape): asynchronous 16-byte
		// copies L2 -> shared memory, all of them in flight at once and none of them through registers -- one round trip where the texel
		// loop below needs up to four batches (C5: 4 KB per group patch)
		const uint32_t row_bytes = R.w * BPP;
		if ((row_bytes & 15u) == 0u && ((reinterpret_cast<uint64_t>(g) | ((uint64_t)LW * BPP) | ((uint64_t)R.ox * BPP)) & 15u) == 0u) {
			const uint32_t vsh = 31u - __clz(row_bytes >> 4), nv = (n * BPP) >> 4;
			for (uint32_t i = lane; i < nv; i += 32u) {
				const uint32_t vx = i & ((1u << vsh) - 1u), row = i >> vsh, y = row & (R.h - 1u), z = row >> sh;
				const uint8_t* src = g + ((uint64_t)(R.oz + z) * LH * LW + (uint64_t)(R.oy + y) * LW + R.ox) * BPP + (size_t)vx * 16u;
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst + (size_t)i * 16u)), "l"(src) : "memory");
			}
			asm volatile("cp.async.wait_all;" ::: "memory");
			__syncwarp();
			return;
		}
	}
	for (uint32_t base = 0; base < n; base += 32u * U) {
		uint32_t t[U][IO::NW];
#pragma unroll
		for (uint32_t u = 0; u < U; ++u) {
			const uint32_t i = base + u * 32u + lane;
			if (i < n) {
				const uint32_t x = i & (R.w - 1u), y = (i >> sw) & (R.h - 1u), z = i >> (sw + sh);
				IO::template load<true>(g + ((uint64_t)(R.oz + z) * LH * LW + (uint64_t)(R.oy + y) * LW + (R.ox + x)) * BPP, t[u]);
			}
		}
#pragma unroll
		for (uint32_t u = 0; u < U; ++u) {
			const uint32_t i = base + u * 32u + lane;
			if (i < n) IO::store(smem_dst + (size_t)i * BPP, t[u]);
		}
	}
	__syncwarp();
}

// classic "last block" protocol (threadfence + atomic ticket); returns true for the warp that arrives last.
// The counter is reset by that warp, so a relaunch needs no memset.
// Release / acquire ride on the atomic itself (atom.acq_rel.gpu) instead of a sequentially consistent __threadfence():
// the other lanes' stores are ordered before it by __syncwarp (cumulativity), their later loads after it likewise.
__device__ __forceinline__ uint32_t atom_add_acq_rel_gpu(uint32_t* p, uint32_t v) {
	uint32_t old;
	asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
	return old;
}
// the same with `count` arrivals at once (a unit of tiles that one CTA joined in shared memory)
__device__ __forceinline__ bool arrive_last_n(uint32_t* counter, uint32_t count, uint32_t expected, uint32_t lane) {
	__syncwarp();
	uint32_t last = 0;
	if (lane == 0) {
		const uint32_t old = atom_add_acq_rel_gpu(counter, count);
		last = (old + count == expected);
		if (last) *counter = 0u;
	}
	last = __shfl_sync(0xFFFFFFFFu, last, 0);
	__syncwarp();
	return last != 0;
}
__device__ __forceinline__ bool arrive_last(uint32_t* counter, uint32_t expected, uint32_t lane) {
	__syncwarp();
	uint32_t last = 0;
	if (lane == 0) {
		const uint32_t old = atom_add_acq_rel_gpu(counter, 1u);
		last = (old == expected - 1u);
		if (last) *counter = 0u;
	}
	last = __shfl_sync(0xFFFFFFFFu, last, 0);
	__syncwarp(); // extends lane 0's acquire to the other lanes' later loads (shfl.sync alone orders no memory accesses)
	return last != 0;
}

// ------------------------------------------------------------------------------------------------------
// the single-pass kernel: persistent CTAs, TMA producer warp + 8 consumer warps, `stages`-deep tile ring
// ------------------------------------------------------------------------------------------------------
// A "unit" is what one CTA works through in one go and what gets published to the other CTAs: 2 x 2 (x 2) adjacent
// tiles (unit_shift = 1) or a single tile (unit_shift = 0, images fewer than two tiles wide).  Tile index = unit * tiles
// per unit + sub-tile.
struct TileCoord {
	uint32_t x, y, z, layer; // tile coordinates
	uint32_t ux, uy, uz, k;  // unit coordinates, sub-tile within the unit
};
template <int DIMS> __device__ __forceinline__ TileCoord tile_coord(const flmip_fast_params& P, uint32_t t) {
	TileCoord c;
	const uint32_t us = P.unit_shift;
	c.k = t & ((1u << (us * DIMS)) - 1u);
	uint32_t u = t >> (us * DIMS);
	c.ux = u & (P.units[0] - 1u); u >>= P.unit_cshift[0];
	c.uy = u & (P.units[1] - 1u); u >>= P.unit_cshift[1];
	if constexpr (DIMS == 3) { c.uz = u & (P.units[2] - 1u); c.layer = u >> P.unit_cshift[2]; }
	else { c.uz = 0; c.layer = u; }
	c.x = (c.ux << us) | (c.k & us);
	c.y = (c.uy << us) | ((c.k >> 1) & us);
	c.z = (c.uz << us) | ((c.k >> 2) & us);
	return c;
}

// Levels IN_REG_LEVELS+1 .. of one tile from the tile's cascade slot; returns the tile's remainder region.
// Runs on a finisher warp, off the consumers' critical path.
template <uint32_t EK, int CH, int DIMS>
__device__ __forceinline__ Region finish_tile(uint8_t* buf_a, uint8_t* buf_b, const flmip_fast_params& P, const TileCoord& tc, uint32_t lane,
											  uint8_t*& remainder) {
	using C = Codec<EK>;
	constexpr int BPP = C::BYTES * CH;
	using TL = flmip_tiling<BPP, DIMS>;
	constexpr uint32_t S0 = TL::IN_REG_LEVELS; // level held in buf_a
	Region R;
	R.lvl = S0;
	R.w = TL::TX >> S0; R.h = TL::TY >> S0; R.d = (DIMS == 3 ? TL::TZ >> S0 : 1u);
	R.ox = tc.x * R.w; R.oy = tc.y * R.h; R.oz = (DIMS == 3 ? tc.z * R.d : 0u);
	uint8_t *src = buf_a, *dst = buf_b;
	cascade_warp<EK, CH, DIMS>(src, dst, R, P, tc.layer, lane);
	remainder = src; // the texels of level R.lvl this tile ends on
	return R;
}

// Group bookkeeping of a unit: its group's counters and how many units the group has.
template <int BPP, int DIMS> struct GroupOf {
	using TL = flmip_tiling<BPP, DIMS>;
	static constexpr uint32_t G = TL::GROUP;
	static constexpr uint32_t GS = (uint32_t)flmip_ilog2(G);
	uint32_t gx, gy, gz, ntx, nty, ntz;
	__device__ __forceinline__ GroupOf(const flmip_fast_params& P, const TileCoord& tc) { // G x G (x G) units
		gx = tc.ux >> GS; gy = tc.uy >> GS; gz = (DIMS == 3 ? tc.uz >> GS : 0u);
		ntx = min(G, P.units[0] - gx * G); nty = min(G, P.units[1] - gy * G); ntz = (DIMS == 3 ? min(G, P.units[2] - gz * G) : 1u);
	}
	__device__ __forceinline__ uint64_t index(const flmip_fast_params& P, uint32_t layer) const {
		const uint32_t groups_per_layer = P.groups[0] * P.groups[1] * P.groups[2];
		return (uint64_t)layer * groups_per_layer + (gz * P.groups[1] + gy) * P.groups[0] + gx;
	}
	__device__ __forceinline__ uint32_t* counter(const flmip_fast_params& P, uint32_t layer) const {
		return reinterpret_cast<uint32_t*>(P.counters) + index(P, layer);
	}
	__device__ __forceinline__ uint32_t units() const { return ntx * nty * ntz; }
};

// Last arriver of a group of units: reduces the group's patch and, if its group is the last one of the layer, the rest of
// the chain.  Runs in the CTA's patch buffer, serialised between the finisher warps of a CTA by a shared-memory lock.
template <uint32_t EK, int CH, int DIMS>
__device__ __forceinline__ void finish_group(uint8_t* patch_a, uint8_t* patch_b, uint32_t* patch_lock, const flmip_fast_params& P,
											 const TileCoord& tc, Region R, uint32_t lane) {
	using C = Codec<EK>;
	constexpr int BPP = C::BYTES * CH;
	using TL = flmip_tiling<BPP, DIMS>;
	constexpr uint32_t G = TL::GROUP;
	const uint32_t layer = tc.layer;
	const GroupOf<BPP, DIMS> grp(P, tc);

	if (lane == 0) FLMIP_STAMP_MAX(P, 4); // last arriver of a group: its publish has returned
	if (lane == 0) {
		while (atomicCAS(patch_lock, 0u, 1u) != 0u) __nanosleep(64);
	}
	__syncwarp();
	// patch = this group's part of the level the units ended on (R.w/h/d are the per-unit remainders here)
	R.ox = grp.gx * G * R.w; R.oy = grp.gy * G * R.h; R.oz = grp.gz * G * R.d;
	R.w *= grp.ntx; R.h *= grp.nty; R.d *= grp.ntz;
	uint8_t *src = patch_a, *dst = patch_b;
	gather_region<BPP, DIMS>(src, R, P, layer, lane);
	if (lane == 0) FLMIP_STAMP_MAX(P, 5); // group patch gathered
	cascade_warp<EK, CH, DIMS>(src, dst, R, P, layer, lane, 8u);
	if (lane == 0) FLMIP_STAMP_MAX(P, 6); // group patch reduced

	// ---- layer stage: the last group of a layer finishes the chain -----------------------------------------
	const uint32_t groups_per_layer = P.groups[0] * P.groups[1] * P.groups[2];
	if (next_level_has_texels<DIMS>(P, R.lvl) &&
		arrive_last(reinterpret_cast<uint32_t*>(P.counters) + (uint64_t)P.layers * groups_per_layer + layer, groups_per_layer, lane)) {
		R.ox = R.oy = R.oz = 0;
		R.w = P.dim[0] >> R.lvl; R.h = P.dim[1] >> R.lvl; R.d = (DIMS == 3 ? P.dim[2] >> R.lvl : 1u);
		src = patch_a; dst = patch_b;
		gather_region<BPP, DIMS>(src, R, P, layer, lane);
		cascade_warp<EK, CH, DIMS>(src, dst, R, P, layer, lane);
		if (lane == 0) FLMIP_STAMP_MAX(P, 7); // layer stage done: the chain of this layer is complete
	}
	__syncwarp();
	if (lane == 0) {
		__threadfence_block();
		atomicExch(patch_lock, 0u);
	}
}

// warp roles of one persistent CTA
constexpr int FLMIP_CONSUMER_WARPS = 8;                           // warps 0..7: tile -> levels 1 (+2) in registers
constexpr int FLMIP_CONSUMER_THREADS = FLMIP_CONSUMER_WARPS * 32;
constexpr int FLMIP_PRODUCER_WARP = FLMIP_CONSUMER_WARPS;         // warp 8: TMA loads
constexpr int FLMIP_FINISHER_WARP0 = FLMIP_PRODUCER_WARP + 1;     // warps 9..: remaining levels of a tile + last-arriver stages
static_assert(FLMIP_BLOCK_THREADS == (FLMIP_FINISHER_WARP0 + FLMIP_FINISHER_WARPS) * 32, "block size out of sync with mip_params.h");

template <uint32_t EK, int CH, int DIMS>
__device__ __forceinline__ void fast_body(const CUtensorMap& tmap, const flmip_fast_params& P) {
	using C = Codec<EK>;
	constexpr int BPP = C::BYTES * CH;
	using TL = flmip_tiling<BPP, DIMS>;
	constexpr bool WIDE = (BPP == 16);            // x pair spans the two chunks of a thread
	constexpr int ROW_BYTES = TL::TILE_BYTES_X;   // bytes of one tile row in shared memory
	static_assert(TL::THREADS == FLMIP_CONSUMER_THREADS, "one consumer thread per 32 B x 4 rows (2D) / 32 B x 2 x 2 (3D)");

	// [stages x tile][FLMIP_FINISHER_WARPS cascade slots (buf_a, buf_b)][patch buffer]; tiles are TMA destinations: 128-byte aligned
	extern __shared__ __align__(128) uint8_t smem_raw[];
	__shared__ uint64_t full_bar[FLMIP_MAX_STAGES], empty_bar[FLMIP_MAX_STAGES];
	__shared__ uint64_t slot_full[FLMIP_FINISHER_WARPS], slot_empty[FLMIP_FINISHER_WARPS];
	__shared__ uint32_t finisher_ticket, patch_lock, unit_count[2];
	__shared__ __align__(16) uint8_t unit_patch[2][FLMIP_UNIT_PATCH_BYTES + FLMIP_UNIT_PATCH_BYTES / 4u];
	__shared__ uint32_t stage_tile[FLMIP_MAX_STAGES], slot_tile[FLMIP_FINISHER_WARPS]; // tile index riding along with the data
	// ... and where the tile's part of level 1 ([0]) and level 2 ([1], 2D) starts in global memory: computed once per tile by the
	// producer instead of by every consumer thread (tile index -> coordinates -> 64-bit level / layer / row offsets)
	__shared__ __align__(16) uint64_t stage_org[FLMIP_MAX_STAGES][2];
	const uint32_t stages = P.stages;
	uint8_t* const cascade_base = smem_raw + (size_t)stages * TL::TILE_BYTES;

	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

	if (tid == 0) {
		finisher_ticket = 0;
		patch_lock = 0;
		unit_count[0] = unit_count[1] = 0;
		for (uint32_t s = 0; s < stages; ++s) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], FLMIP_CONSUMER_WARPS);
		}
		for (uint32_t f = 0; f < FLMIP_FINISHER_WARPS; ++f) {
			mbar_init(&slot_full[f], FLMIP_CONSUMER_WARPS);
			mbar_init(&slot_empty[f], 1);
		}
		fence_mbar_init();
	}
	__syncthreads();
	pdl_start(P.late_wait);
	if (tid == 0) FLMIP_STAMP(P, 0); // CTA may touch global memory from here

	constexpr uint32_t SLOT_BYTES = TL::CASCADE_BYTES + TL::CASCADE_BYTES / 4u;
	// the finisher pool is idle when the consumers' in-register levels end the chain
	const bool need_finish = P.level_count > TL::IN_REG_LEVELS + 1u;
	if (warp >= FLMIP_FINISHER_WARP0) {
		// ---- finishers: a pool of warps that take the CTA's tiles in ring order by ticket ---------------------
		if (!need_finish) return; // the consumers produce every level
		uint8_t* const patch_a = cascade_base + FLMIP_FINISHER_WARPS * SLOT_BYTES;
		uint8_t* const patch_b = patch_a + TL::CASCADE_BYTES;
		const uint32_t kbits = P.unit_shift * DIMS;      // log2(tiles per unit)
		const uint32_t unit_tiles = 1u << kbits;
		for (;;) {
			uint32_t n = 0;
			if (lane == 0) n = atomicAdd(&finisher_ticket, 1u);
			n = __shfl_sync(0xFFFFFFFFu, n, 0);
			const uint32_t slot = n % FLMIP_FINISHER_WARPS, use = n / FLMIP_FINISHER_WARPS;
			uint8_t* const buf_a = cascade_base + slot * SLOT_BYTES;
#ifdef FLMIP_FINISHER_SLEEP_WAIT
			mbar_wait_sleep(&slot_full[slot], use & 1u);
#else
			mbar_wait(&slot_full[slot], use & 1u);
#endif
			const uint32_t t = slot_tile[slot];
			if (t == FLMIP_NO_TILE) break; // the consumers are done: one sentinel per finisher warp
			const TileCoord tc = tile_coord<DIMS>(P, t);
			uint8_t* rem = nullptr;
			Region R = finish_tile<EK, CH, DIMS>(buf_a, buf_a + TL::CASCADE_BYTES, P, tc, lane, rem);
			bool carry_on = next_level_has_texels<DIMS>(P, R.lvl);
			if (carry_on && kbits != 0) {
				// ---- unit stage: the remainders of the unit's tiles meet in shared memory; whoever brings the last one
				//      reduces them one more level.  The CTA works through its units in order, two buffers suffice
				//      (a slot is only released after its tile is in, and FLMIP_FINISHER_WARPS <= tiles per unit + 1).
				const uint32_t q = (n >> kbits) & 1u;
				uint8_t* const up = unit_patch[q];
				const uint32_t kx = tc.k & 1u, ky = (tc.k >> 1) & 1u, kz = (DIMS == 3 ? tc.k >> 2 : 0u);
				const uint32_t sw = 31u - __clz(R.w), sh = 31u - __clz(R.h);
				for (uint32_t i = lane; i < R.w * R.h * R.d; i += 32) {
					const uint32_t x = i & (R.w - 1u), y = (i >> sw) & (R.h - 1u), z = i >> (sw + sh);
					uint32_t texel[TexelIO<BPP>::NW];
					TexelIO<BPP>::template load<false>(rem + (size_t)i * BPP, texel);
					TexelIO<BPP>::store(up + ((size_t)((kz * R.d + z) * (2u * R.h) + ky * R.h + y) * (2u * R.w) + kx * R.w + x) * BPP, texel);
				}
				__syncwarp();
				uint32_t last = 0;
				if (lane == 0) {
					__threadfence_block();
					last = (atomicAdd(&unit_count[q], 1u) == unit_tiles - 1u);
					if (last) {
						unit_count[q] = 0u;
						__threadfence_block();
					}
				}
				carry_on = __shfl_sync(0xFFFFFFFFu, last, 0) != 0;
				if (carry_on) {
					R.w *= 2u; R.h *= 2u; if (DIMS == 3) R.d *= 2u;
					R.ox = tc.ux * R.w; R.oy = tc.uy * R.h; R.oz = (DIMS == 3 ? tc.uz * R.d : 0u);
					uint8_t *src = up, *dst = up + FLMIP_UNIT_PATCH_BYTES;
					cascade_warp<EK, CH, DIMS>(src, dst, R, P, tc.layer, lane);
					carry_on = next_level_has_texels<DIMS>(P, R.lvl);
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&slot_empty[slot]); // the slot is free again before the slow part starts
			if (!carry_on) continue;
			// publish the unit; the last arriver of its group carries on
			const GroupOf<BPP, DIMS> grp(P, tc);
			if (arrive_last(grp.counter(P, tc.layer), grp.units(), lane)) finish_group<EK, CH, DIMS>(patch_a, patch_b, &patch_lock, P, tc, R, lane);
		}
		if (lane == 0) FLMIP_STAMP_MAX(P, 3); // last finisher warp of the CTA done
		return;
	}
	if (warp == FLMIP_PRODUCER_WARP) {
		// ---- producer: dynamic tile scheduler + TMA issue ----------------------------------------------------
		// Units are handed out by a global atomic counter, so an SM that gets less memory bandwidth simply takes
		// fewer of them.  An atomic round trip costs microseconds under load: lanes 0..PF-1 each keep one fetch in
		// flight, lane (it % PF) holds the tile of iteration `it`.
		// Depth: a fetched unit is committed to this CTA, so every fetch in flight is work the last wave cannot rebalance.  A unit
		// of 2 x 2 (x 2) tiles lasts longer than a round trip: 2 in flight (C2 +1.3 %, C5 +2.7 % over 4; 8 and 16 are slower
		// still, profiles/r1/10_timeline.txt); single tiles need FLMIP_SCHED_PREFETCH.
		const uint32_t PF = P.unit_shift ? FLMIP_UNIT_PREFETCH : FLMIP_SCHED_PREFETCH;
		uint32_t* const sched = reinterpret_cast<uint32_t*>(P.sched);
		// The first unit of a CTA is its own index (no round trip before the first load); the counter hands out the rest.
		uint32_t pf = FLMIP_NO_TILE;
		if (lane == 0) pf = blockIdx.x;
		else if (lane < PF) pf = gridDim.x + atomicAdd(&sched[0], 1u);
		uint32_t s = 0, use = 0, dead = 0;
		const uint32_t kbits = P.unit_shift * DIMS;
		for (uint32_t it = 0; dead < PF; ++it) {
			const uint32_t u = __shfl_sync(0xFFFFFFFFu, pf, it % PF);
			if (u >= P.total_units) { ++dead; continue; } // this lane ran off the end; the others may still hold units
			dead = 0;
			if (lane == it % PF) pf = gridDim.x + atomicAdd(&sched[0], 1u); // unit of iteration it + PF
			if (lane == 0) {
				for (uint32_t k = 0; k < (1u << kbits); ++k) {
					const uint32_t t = (u << kbits) | k;
					if (use > 0) mbar_wait(&empty_bar[s], (use - 1u) & 1u);
					const TileCoord tc = tile_coord<DIMS>(P, t);
					// the load first, then (behind its latency) what rides along with it; the arrival releases both to the consumers
					mbar_expect_tx(&full_bar[s], TL::TILE_BYTES);
					// innermost coordinate in uint32 units; 2D images use the third tensor dim for the layer
					tma_load_3d(smem_raw + (size_t)s * TL::TILE_BYTES, &tmap, &full_bar[s], (int)(tc.x * (ROW_BYTES / 4)), (int)(tc.y * TL::TY),
								(int)(DIMS == 3 ? tc.z * TL::TZ : tc.layer));
					stage_tile[s] = t;
					{
						const uint64_t l1_pitch = (uint64_t)(P.dim[0] >> 1) * BPP;
						const uint64_t row1 = (DIMS == 3 ? (uint64_t)tc.z * (TL::TZ / 2) * (P.dim[1] >> 1) : 0ull) + (uint64_t)tc.y * (TL::TY / 2);
						stage_org[s][0] = reinterpret_cast<uint64_t>(level_layer_ptr<BPP, DIMS>(P, 1, tc.layer)) + row1 * l1_pitch + (uint64_t)tc.x * (ROW_BYTES / 2);
						if constexpr (DIMS == 2) {
							const uint64_t l2_pitch = (uint64_t)(P.dim[0] >> 2) * BPP;
							stage_org[s][1] = reinterpret_cast<uint64_t>(level_layer_ptr<BPP, DIMS>(P, 2, tc.layer)) + (uint64_t)tc.y * (TL::TY / 4) * l2_pitch +
											  (uint64_t)tc.x * (ROW_BYTES / 4);
						}
					}
					mbar_arrive(&full_bar[s]);
					if (++s == stages) { s = 0; ++use; }
				}
			}
			// every lane tracks the ring position
			if (lane != 0) {
				const uint32_t adv = s + (1u << kbits);
				use += adv / stages;
				s = adv % stages;
			}
			__syncwarp();
		}
		if (lane == 0) {
			FLMIP_STAMP(P, 1); // the scheduler ran dry for this CTA: all of its loads are issued
			// end of work: a sentinel instead of a tile
			if (use > 0) mbar_wait(&empty_bar[s], (use - 1u) & 1u);
			stage_tile[s] = FLMIP_NO_TILE;
			mbar_arrive(&full_bar[s]);
			// the last producer to get here re-arms the scheduler for the next launch (every fetch has been made by then)
			__threadfence();
			if (atomicAdd(&sched[1], 1u) == gridDim.x - 1u) {
				sched[0] = 0u;
				sched[1] = 0u;
			}
		}
		return;
	}

	// ---- consumers ------------------------------------------------------------------------------------------
	// No barrier couples the consumer warps: each walks the tile ring on its own, so warps of one CTA overlap
	// different tiles.
	const uint32_t sel = (lane >> 2) & 1u; // quarter-warps read conflict-free by swapping the chunk order on lane bit 2
	uint32_t s = 0, use = 0, slot = 0, slot_use = 0;
	for (uint32_t it = 0;; ++it) {
		mbar_wait(&full_bar[s], use & 1u);
		const uint32_t t = stage_tile[s];
		if (t == FLMIP_NO_TILE) {
			if (tid == 0) FLMIP_STAMP(P, 2); // consumers saw the sentinel: every tile of the CTA is in registers / written
			break;
		}
		const uint8_t* const tile = smem_raw + (size_t)s * TL::TILE_BYTES;
		uint8_t* const buf_a = cascade_base + slot * SLOT_BYTES;
		const ulonglong2 org = *reinterpret_cast<const ulonglong2*>(stage_org[s]); // read before the stage is handed back
		uint8_t* const g1 = reinterpret_cast<uint8_t*>(org.x);     // the tile's part of level 1
		const uint64_t l1_pitch = (uint64_t)(P.dim[0] >> 1) * BPP; // bytes per level-1 row

		if constexpr (DIMS == 2) {
			// thread = 2 chunks (32 B) x 4 rows
			const uint32_t tx = tid % TL::THREADS_X, ty = tid / TL::THREADS_X;
			uint32_t raw[4][2][4];
#pragma unroll
			for (int r = 0; r < 4; ++r) {
#pragma unroll
				for (int k = 0; k < 2; ++k) {
					const uint4 v = *reinterpret_cast<const uint4*>(tile + (4 * ty + r) * ROW_BYTES + (2 * tx + (k ^ sel)) * 16);
					raw[r][k][0] = v.x; raw[r][k][1] = v.y; raw[r][k][2] = v.z; raw[r][k][3] = v.w;
				}
			}
			// the tile now lives in registers: hand the stage back to the producer
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty_bar[s]);

			uint32_t l1[2][4]; // two level-1 rows of 16 bytes, logical (left, right) order
			// unorm8 / unorm16 texels of 2 .. 4 bytes with >= 2 channels (RGBA8, RG8, RG16): level 2 of a 16-byte chunk depends on
			// that chunk alone and is computed from the encoder's own "magic" pairs of level 1 (no byte extraction)
			constexpr bool MAGIC_L2 = C::HAS_MAGIC && CH >= 2 && BPP <= 4;
			constexpr int MNP = MAGIC_L2 ? 4 / C::BYTES : 1;
			f32x2_t mg[2][2][MNP]; // [level-1 row][chunk][pair]
			if constexpr (!WIDE) {
#pragma unroll
				for (int j = 0; j < 2; ++j) {
					uint32_t o0[2], o1[2];
					if constexpr (MAGIC_L2) {
						reduce_rows_2d_keep<EK, CH, 4>(raw[2 * j][0], raw[2 * j + 1][0], o0, mg[j][0], P.no_double);
						reduce_rows_2d_keep<EK, CH, 4>(raw[2 * j][1], raw[2 * j + 1][1], o1, mg[j][1], P.no_double);
					} else {
						reduce_rows_2d<EK, CH, 4>(raw[2 * j][0], raw[2 * j + 1][0], o0, P.no_double);
						reduce_rows_2d<EK, CH, 4>(raw[2 * j][1], raw[2 * j + 1][1], o1, P.no_double);
					}
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
				*reinterpret_cast<uint4*>(g1 + (uint64_t)(2 * ty + j) * l1_pitch + tx * 16) =
					make_uint4(l1[j][0], l1[j][1], l1[j][2], l1[j][3]);
			}
			if constexpr (!WIDE) {
				// level 2 in registers: 8 bytes per thread
				if (P.level_count > 2) {
					uint32_t l2[2];
					if constexpr (MAGIC_L2) {
						uint32_t w0[1], w1[1];
						reduce_magic_2d<EK, CH, MNP>(mg[0][0], mg[1][0], w0, P.no_double);
						reduce_magic_2d<EK, CH, MNP>(mg[0][1], mg[1][1], w1, P.no_double);
						l2[0] = sel ? w1[0] : w0[0];
						l2[1] = sel ? w0[0] : w1[0];
					} else {
						reduce_rows_2d<EK, CH, 4>(l1[0], l1[1], l2, P.no_double);
					}
					uint8_t* const g2 = reinterpret_cast<uint8_t*>(org.y);
					const uint64_t l2_pitch = (uint64_t)(P.dim[0] >> 2) * BPP;
					*reinterpret_cast<uint2*>(g2 + (uint64_t)ty * l2_pitch + tx * 8) = make_uint2(l2[0], l2[1]);
					if (need_finish) {
						if (slot_use > 0) mbar_wait(&slot_empty[slot], (slot_use - 1u) & 1u);
						*reinterpret_cast<uint2*>(buf_a + ty * (ROW_BYTES / 4) + tx * 8) = make_uint2(l2[0], l2[1]);
					}
				}
			} else {
				// 16-byte texels: a thread holds one level-1 texel per row, its x neighbour lives in lane ^ 1.
				// Even lanes fetch it with shuffles and produce the level-2 texel.
				if (P.level_count > 2) {
					uint32_t ra[8], rb[8];
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						ra[i] = l1[0][i]; rb[i] = l1[1][i];
						ra[4 + i] = __shfl_xor_sync(0xFFFFFFFFu, l1[0][i], 1);
						rb[4 + i] = __shfl_xor_sync(0xFFFFFFFFu, l1[1][i], 1);
					}
					if (!(lane & 1u)) {
						uint32_t l2[4];
						reduce_rows_2d<EK, CH, 8>(ra, rb, l2, P.no_double);
						uint8_t* const g2 = reinterpret_cast<uint8_t*>(org.y);
						const uint64_t l2_pitch = (uint64_t)(P.dim[0] >> 2) * BPP;
						*reinterpret_cast<uint4*>(g2 + (uint64_t)ty * l2_pitch + (tx >> 1) * 16) =
							make_uint4(l2[0], l2[1], l2[2], l2[3]);
						if (need_finish) {
							if (slot_use > 0) mbar_wait(&slot_empty[slot], (slot_use - 1u) & 1u);
							*reinterpret_cast<uint4*>(buf_a + ty * (ROW_BYTES / 4) + (tx >> 1) * 16) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
						}
					}
				}
			}
		} else {
			// 3D: thread = 2 chunks x 2 rows x 2 slices -> 16 bytes of level 1
			const uint32_t tx = tid % TL::THREADS_X, ty = (tid / TL::THREADS_X) % TL::THREADS_Y, tz = tid / (TL::THREADS_X * TL::THREADS_Y);
			static_assert(TL::THREADS_X == 4, "bank-conflict swizzle assumes 4 chunk pairs per tile row");
			uint32_t raw[2][2][2][4]; // [slice][row][k]
#pragma unroll
			for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
				for (int r = 0; r < 2; ++r) {
#pragma unroll
					for (int k = 0; k < 2; ++k) {
						const uint4 v = *reinterpret_cast<const uint4*>(tile + ((2 * tz + sl) * TL::TY + (2 * ty + r)) * ROW_BYTES + (2 * tx + (k ^ sel)) * 16);
						raw[sl][r][k][0] = v.x; raw[sl][r][k][1] = v.y; raw[sl][r][k][2] = v.z; raw[sl][r][k][3] = v.w;
					}
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty_bar[s]);

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
				for (int sl = 0; sl < 2; ++sl)
#pragma unroll
					for (int r = 0; r < 2; ++r)
#pragma unroll
						for (int i = 0; i < 4; ++i) {
							rr[sl][r][i] = sel ? raw[sl][r][1][i] : raw[sl][r][0][i];
							rr[sl][r][4 + i] = sel ? raw[sl][r][0][i] : raw[sl][r][1][i];
						}
				reduce_rows_3d<EK, CH, 8>(rr[0][0], rr[0][1], rr[1][0], rr[1][1], l1, P.no_double);
			}
			const uint64_t l1_rows = P.dim[1] >> 1;
			*reinterpret_cast<uint4*>(g1 + ((uint64_t)tz * l1_rows + ty) * l1_pitch + tx * 16) =
				make_uint4(l1[0], l1[1], l1[2], l1[3]);
			if (need_finish) {
				if (slot_use > 0) mbar_wait(&slot_empty[slot], (slot_use - 1u) & 1u);
				*reinterpret_cast<uint4*>(buf_a + (tz * (TL::TY / 2) + ty) * (ROW_BYTES / 2) + tx * 16) = make_uint4(l1[0], l1[1], l1[2], l1[3]);
			}
		}

		// this warp's part of buf_a is written: hand the slot to the finisher pool (arrive = release)
		if (need_finish) {
			if (tid == 0) slot_tile[slot] = t; // only after this thread's wait on slot_empty above
			__syncwarp();
			if (lane == 0) mbar_arrive(&slot_full[slot]);
		}

		if (++s == stages) { s = 0; ++use; }
		if (++slot == FLMIP_FINISHER_WARPS) { slot = 0; ++slot_use; }
	}
	// release the finisher pool: one sentinel per finisher warp in the next FLMIP_FINISHER_WARPS slots
	if (need_finish) {
		for (uint32_t k = 0; k < FLMIP_FINISHER_WARPS; ++k) {
			if (slot_use > 0) mbar_wait(&slot_empty[slot], (slot_use - 1u) & 1u);
			if (tid == 0) slot_tile[slot] = FLMIP_NO_TILE;
			__syncwarp();
			if (lane == 0) mbar_arrive(&slot_full[slot]);
			if (++slot == FLMIP_FINISHER_WARPS) { slot = 0; ++slot_use; }
		}
	}
}

// ------------------------------------------------------------------------------------------------------
// general path: literal replay of host_device_image::read_linear + fixed_image::read/write
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float wrap01(float v) { // const_math.hpp:859-869 with max = 1
	if (v < 0.0f) return fminf(__fadd_rn(1.0f, fmodf(v, 1.0f)), __uint_as_float(__float_as_uint(1.0f) - 1u));
	return fmodf(v, 1.0f);
}

// FORMAT_2 / FORMAT_4 normalized texels (host_image.hpp:333-383, 391-460): channel i sits at bits 6 - 2i of byte 0 (2 bits) or in
// the high (even i) / low nibble of byte i / 2 (4 bits).  The signed variants reproduce the reference as written: its sign fix-up
// `x & high_bit != 0u` parses as `x & 1`, so a channel with bit 0 set becomes -(x ^ high_bit) (pinned against the reference's
// own build: tests/test_reference_pin.py).
template <int BITS, bool SIGNED> __device__ __forceinline__ float packed_dec(const uint8_t* p, uint32_t i) {
	const uint32_t v = BITS == 2 ? (p[0] >> (6u - 2u * i)) & 0x3u : (p[i >> 1] >> ((i & 1u) ? 0u : 4u)) & 0xFu;
	if constexpr (!SIGNED) {
		return __fmul_rn(__uint2float_rn(v), BITS == 2 ? (float)(1.0 / 3.0) : (float)(1.0 / 15.0));
	} else {
		int sv = (int)v;
		if (v & 1u) sv = -(int)(signed char)(v ^ (1u << (BITS - 1)));
		return __fmul_rn(__int2float_rn(sv), BITS == 2 ? 1.0f : (float)(1.0 / 7.0));
	}
}
template <int BITS, bool SIGNED> __device__ __forceinline__ uint32_t packed_enc(float c) {
	if constexpr (!SIGNED) {
		const uint32_t q = (uint32_t)(unsigned char)__float2int_rz(__fmul_rn(c, (float)((1 << BITS) - 1)));
		return q & ((1u << BITS) - 1u);
	} else {
		const int q = (int)(signed char)__float2int_rz(__fmul_rn(c, (float)((1 << (BITS - 1)) - 1)));
		return ((uint32_t)q & ((1u << (BITS - 1)) - 1u)) | (q < 0 ? 1u << (BITS - 1) : 0u);
	}
}

template <uint32_t EK>
__device__ __forceinline__ void generic_texel(const flmip_generic_params& P, uint64_t idx) {
	constexpr bool PACKED = EK >= FLMIP_EK_COUNT;
	constexpr int PBITS = (EK == FLMIP_EK_UNORM2 || EK == FLMIP_EK_SNORM2) ? 2 : 4;
	constexpr bool PSIGNED = (EK == FLMIP_EK_SNORM4 || EK == FLMIP_EK_SNORM2);
	using C = Codec<PACKED ? (uint32_t)FLMIP_EK_UNORM8 : EK>; // packed formats sample as float: only lerp_t is used from the codec
	const uint32_t dc = P.dc, ch = P.channels, bpp = PACKED ? (PBITS * ch) >> 3 : C::BYTES * ch;
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
			if constexpr (PACKED) {
				v[k][i] = __float_as_uint(packed_dec<PBITS, PSIGNED>(p, i));
			} else {
				uint32_t raw;
				if constexpr (C::BYTES == 4) raw = *reinterpret_cast<const uint32_t*>(p + 4 * i);
				else if constexpr (C::BYTES == 2) raw = *reinterpret_cast<const unsigned short*>(p + 2 * i);
				else raw = p[i];
				v[k][i] = C::dec(raw);
			}
		}
	}
	for (uint32_t d = 0; d < dc; ++d) {
		const uint32_t step = 1u << d;
		for (uint32_t k = 0; k < n; k += 2u * step)
			for (uint32_t i = 0; i < ch; ++i) v[k][i] = C::lerp_t(v[k][i], v[k + step][i], w[d]);
	}
	const uint32_t texel = (dc == 1 ? g[0] : (dc == 2 ? P.dst_dim[0] * g[1] + g[0] : P.dst_dim[0] * P.dst_dim[1] * g[2] + P.dst_dim[0] * g[1] + g[0]));
	uint8_t* q = reinterpret_cast<uint8_t*>(P.base) + P.dst_off + (uint64_t)layer * P.dst_slice + (uint64_t)texel * bpp;
	if constexpr (PACKED) {
		uint32_t bytes[2] = { 0u, 0u }; // the texel's bytes are rebuilt from zero (insert_channels memsets them)
		for (uint32_t i = 0; i < ch; ++i) {
			const uint32_t bits = packed_enc<PBITS, PSIGNED>(__uint_as_float(v[0][i]));
			if constexpr (PBITS == 2) bytes[0] |= bits << (6u - 2u * i);
			else bytes[i >> 1] |= bits << ((i & 1u) ? 0u : 4u);
		}
		for (uint32_t b = 0; b < bpp; ++b) q[b] = (uint8_t)bytes[b];
		return;
	}
	for (uint32_t i = 0; i < ch; ++i) {
		const uint32_t raw = C::enc(v[0][i], P.no_double);
		if constexpr (C::BYTES == 4) *reinterpret_cast<uint32_t*>(q + 4 * i) = raw;
		else if constexpr (C::BYTES == 2) *reinterpret_cast<unsigned short*>(q + 2 * i) = (unsigned short)raw;
		else q[i] = (uint8_t)raw;
	}
}


// ------------------------------------------------------------------------------------------------------
// multi-level tile kernel (flmip_tile2d_* / flmip_tile3d_*): any image size.
//
// For NPOT levels the sample point (2g+1) * fl(1/N) * N is 2g+1 +- a few ulp, so along each axis the reference's
// sampler (host_image.hpp:141-174, 869-929) reads the "active" texel B (the one the point falls into: 2g or 2g+1)
// and its neighbour A on the side of the point, with a weight of B of 0.5 +- eps:  out = (B - A) * t + A.
// axis_fetch() replays that literally.  The two texels are {2g, 2g+1} with ONE exception that the reference really
// has: for g = 0 and sizes N with fl(fl(1/N) * N) == pred(1.0f) (41, 47, 55, 61, 82, ...), the neighbour coordinate
// 0.99999994 + 1 rounds to 2.0, so A is texel 2 and texel 1 is skipped (tests/test_npot_weights.py pins this down
// for every size the device supports).  Either way level l+k of a tile depends only on texels [2g, 2g+2] of the
// previous level, so one CTA can take a 64 x 64 (32 x 16 x 16) source tile through up to 6 (4) levels in shared
// memory; the host cuts a launch short where texel 2 would lie outside a 2-texel-wide remainder.  Partial tiles at
// the image border are masked.  Every level is re-decoded from its stored (quantised) bits.
// ------------------------------------------------------------------------------------------------------
struct __align__(16) axis_f {
	float t;       // weight of B
	uint32_t a, b; // texel indices of A (outside) and B (active) in the source level
	uint32_t pad;  // 16 bytes: one LDS.128 per table entry
};
// Same values as the literal replay in generic_texel(), computed without the quarter-rate XU pipe and without fmodf:
// for 0 <= x < 2^23, x + 2^23 rounded toward zero holds floor(x) in its mantissa (the codecs use the same identity),
// float(u) for u < 2^23 is (2^23 | u) - 2^23, and wrap(coord, 1) is the identity for 0 <= coord < 1 (always the case:
// (2g+1) * fl(1/N) < 1 for g < N / 2); anything else takes the literal path.
__device__ __forceinline__ axis_f axis_fetch(uint32_t g, float inv_prev, float fdim, float fdim_excl) {
	constexpr float MAGIC = 8388608.0f;
	const uint32_t n = g * 2u + 1u;
	const float fn = n < 0x800000u ? __fsub_rn(__uint_as_float(0x4B000000u | n), MAGIC) : __uint2float_rn(n);
	const float coord = __fmul_rn(fn, inv_prev);                             // mip_map_minify.hpp:106
	const float m = __fmul_rn(coord, fdim);                                  // host_image.hpp:168-172, 875
	axis_f r;
	if (coord >= 0.0f && coord < 1.0f && m < 4194304.0f) {
		const float fl = __fsub_rn(__fadd_rz(m, MAGIC), MAGIC);              // floorf(m), m >= 0
		const float frac = __fsub_rn(m, fl);                                 // const_math.hpp:308-313
		const bool lo = frac < 0.5f;
		r.t = lo ? __fadd_rn(frac, 0.5f) : __fsub_rn(1.5f, frac);
		const float mb = m > fdim_excl ? fdim_excl : m;
		float ma = __fadd_rn(m, lo ? -1.0f : 1.0f);
		ma = ma > fdim_excl ? fdim_excl : (ma < 0.0f ? 0.0f : ma);
		r.b = __float_as_uint(__fadd_rz(mb, MAGIC)) & 0x7FFFFFu;
		r.a = __float_as_uint(__fadd_rz(ma, MAGIC)) & 0x7FFFFFu;
	} else {
		const float scaled = __fmul_rn(wrap01(coord), fdim);
		const float frac = __fsub_rn(scaled, floorf(scaled));
		r.t = frac < 0.5f ? __fadd_rn(frac, 0.5f) : __fsub_rn(1.5f, frac);
		const float ma = __fadd_rn(m, frac < 0.5f ? -1.0f : 1.0f);
		r.b = (uint32_t)__float2ll_rz(m > fdim_excl ? fdim_excl : (m < 0.0f ? 0.0f : m));
		r.a = (uint32_t)__float2ll_rz(ma > fdim_excl ? fdim_excl : (ma < 0.0f ? 0.0f : ma));
	}
	return r;
}

template <int D> struct tile_geo;
template <> struct tile_geo<2> { static constexpr int TX = FLMIP_TILE2D_X, TY = FLMIP_TILE2D_Y, TZ = 1, MAXLEV = FLMIP_TILE_MAX_LEVELS; };
template <> struct tile_geo<3> { static constexpr int TX = FLMIP_TILE3D_X, TY = FLMIP_TILE3D_Y, TZ = FLMIP_TILE3D_Z, MAXLEV = FLMIP_TILE3D_MAX_LEVELS; };

// reduce the 2^D raw texels of one fetch (index bit d = 0: texel A of axis d, 1: texel B) to one stored texel: x, then y, then z
template <uint32_t EK, int CH, int D, int NW>
__device__ __forceinline__ void tile_reduce_block(const uint32_t (&raw)[1 << D][NW], const axis_f (&af)[D], uint32_t no_double, uint32_t (&out)[NW]) {
	using C = Codec<EK>;
	if constexpr (!C::IS_INT && (CH % 2) == 0 && ((CH * C::BYTES) % 4) == 0) {
		// float formats, channel pairs: packed fp32 (FADD2 / FMUL2, per-lane IEEE rounding) -- (b - a) * t + a in three roundings
		f32x2_t p[1 << D][CH / 2];
#pragma unroll
		for (int k = 0; k < (1 << D); ++k)
#pragma unroll
			for (int j = 0; j < CH / 2; ++j) p[k][j] = C::dec2_at(raw[k], 2 * j, 2 * j + 1);
#pragma unroll
		for (int d = 0; d < D; ++d) {
			const int step = 1 << d;
			const f32x2_t t2 = splat2(af[d].t);
			// t is not 0.5 here, so the product is not exact: it has to be rounded on its own (mul2_sep)
#pragma unroll
			for (int k = 0; k < (1 << D); k += 2 * step)
#pragma unroll
				for (int j = 0; j < CH / 2; ++j) p[k][j] = add2_rn(mul2_sep(sub2_rn(p[k + step][j], p[k][j]), t2), p[k][j]);
		}
		C::template enc_pack2<CH / 2>(p[0], out, no_double);
	} else {
		uint32_t v[1 << D][CH];
#pragma unroll
		for (int k = 0; k < (1 << D); ++k)
#pragma unroll
			for (int i = 0; i < CH; ++i) v[k][i] = C::dec_at(raw[k], i);
#pragma unroll
		for (int d = 0; d < D; ++d) {
			const int step = 1 << d;
#pragma unroll
			for (int k = 0; k < (1 << D); k += 2 * step)
#pragma unroll
				for (int i = 0; i < CH; ++i) v[k][i] = C::lerp_t(v[k][i], v[k + step][i], af[d].t);
		}
		if constexpr ((CH * C::BYTES) % 4 == 0) {
			C::template enc_pack<CH>(v[0], out, no_double);
		} else {
#pragma unroll
			for (int w = 0; w < NW; ++w) out[w] = 0u;
#pragma unroll
			for (int i = 0; i < CH; ++i) put_elem<C::BYTES>(out, i, C::enc(v[0][i], no_double));
		}
	}
}

template <uint32_t EK, int CH, int D>
__device__ __forceinline__ void tile_body(const flmip_tile_params& P) {
	using C = Codec<EK>;
	using G = tile_geo<D>;
	constexpr int BPP = C::BYTES * CH;
	using IO = TexelIO<BPP>;
	constexpr int NW = IO::NW;
	constexpr int N1 = (G::TX / 2) * (G::TY / 2) * (D == 3 ? G::TZ / 2 : 1); // texels of level 1 of a tile
	constexpr int N2 = N1 >> D;
	constexpr int THREADS = 256;
	__shared__ uint32_t buf0[N1 * NW];
	__shared__ uint32_t buf1[N2 * NW];
	__shared__ uint32_t buf2[(N2 >> D) * NW]; // 2D: level 3 must not land in buf0 while other warps still read level 1 from it

	const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
	uint32_t tile = blockIdx.x;
	uint32_t ti[3];
	ti[0] = tile % P.tiles[0]; tile /= P.tiles[0];
	ti[1] = tile % P.tiles[1]; tile /= P.tiles[1];
	ti[2] = 0;
	if constexpr (D == 3) { ti[2] = tile % P.tiles[2]; tile /= P.tiles[2]; }
	const uint32_t layer = tile;
	uint8_t* const base = reinterpret_cast<uint8_t*>(P.base);

	// ---- sampler table of the tile: (A, B, t) of every destination row / column / slice of every level this launch produces,
	//      computed once per CTA (126 entries in 2D) instead of by every thread for its own texels.  Heap layout per axis: level k
	//      holds its T >> k entries at [T >> k, 2 * (T >> k)).  Level 1 keeps A and B as texel indices of the source level (they
	//      address global memory), levels >= 2 as indices into the tile's part of level k - 1 in shared memory (the host guarantees
	//      they lie inside the tile's remainder).  Touches only kernel parameters and shared memory: runs before the PDL wait.
	__shared__ axis_f axtab[D][64];
	{
		constexpr uint32_t TD[3] = { (uint32_t)G::TX, (uint32_t)G::TY, (uint32_t)G::TZ };
		constexpr uint32_t TOTAL = G::TX + G::TY + (D == 3 ? G::TZ : 0);
		for (uint32_t idx = t; idx < TOTAL; idx += THREADS) {
			uint32_t d = 0, j = idx;
			if (j >= TD[0]) { j -= TD[0]; d = 1; }
			if (D == 3 && d == 1 && j >= TD[1]) { j -= TD[1]; d = 2; }
			if (j == 0) continue; // heap slot 0 is unused
			const uint32_t td = (d == 0 ? TD[0] : (d == 1 ? TD[1] : TD[2]));
			const uint32_t k = (uint32_t)__clz(j) - (uint32_t)__clz(td); // level: T >> k <= j < 2 * (T >> k)
			if (k > P.nlev || P.dim[k][d] == 0u) continue;
			const uint32_t e = td >> k, pe = 2u * e;                    // extents of levels k and k - 1 of the tile
			const uint32_t g = min(ti[d] * e + (j - e), P.dim[k][d] - 1u); // destination index (clamped: partial tiles are masked later)
			axis_f f = axis_fetch(g, P.inv_prev[k - 1][d], P.fdim[k - 1][d], P.fdim_excl[k - 1][d]);
			if (k >= 2u) {
				f.a = min(f.a - ti[d] * pe, pe - 1u);
				f.b = min(f.b - ti[d] * pe, pe - 1u);
			}
			f.pad = 0u;
			axtab[d][j] = f;
		}
	}
	__syncthreads();
	pdl_start(P.late_wait);

	// ---- level 1: straight from global memory (each warp reads whole 2 * BPP * 32 byte row segments) -------------
	// 2D: warp w owns rows 4w .. 4w+3 of level 1 (so that levels 2 and 3 stay inside the warp); 3D: o = t + c * 256
	{
		constexpr int EX = G::TX / 2, EY = G::TY / 2; // extents of level 1 of the tile
		constexpr int PER_THREAD = N1 / THREADS;
		constexpr int U = ((1 << D) * NW >= 16) ? 2 : 4; // outputs whose loads are in flight together
		static_assert(PER_THREAD % U == 0, "unroll");
		static_assert(D == 3 || (EX == 32 && PER_THREAD == 4), "2D mapping: one row of 32 outputs per warp and step");
		const uint8_t* const src = base + P.level_off[0] + (uint64_t)layer * P.slice[0];
		uint8_t* const dst = base + P.level_off[1] + (uint64_t)layer * P.slice[1];
		const uint32_t W0 = P.dim[0][0], H0 = P.dim[0][1];
		const uint32_t W1 = P.dim[1][0], H1 = P.dim[1][1], D1 = P.dim[1][2];
		// the x index (and in 3D the y index) of a thread's outputs does not depend on the step c
		static_assert(THREADS % EX == 0 && (D < 3 || THREADS % (EX * EY) == 0), "loop-invariant axes");
		constexpr int INV = (D == 3 ? 2 : 1); // number of loop-invariant axes
		axis_f fix[INV];
		fix[0] = axtab[0][EX + t % EX];
		if constexpr (D == 3) fix[1] = axtab[1][EY + (t / EX) % EY];
		const uint64_t pitch = (uint64_t)W0 * BPP;
		const uint32_t xoff[2] = { fix[0].a * BPP, fix[0].b * BPP };
#pragma unroll 1
		for (int c = 0; c < PER_THREAD; c += U) {
			uint32_t raw[U][1 << D][NW];
			uint32_t g[U][3], o[U];
			float wt[U][D];
			bool ok[U];
#pragma unroll
			for (int u = 0; u < U; ++u) {
				o[u] = (D == 2 ? (warp * PER_THREAD + (uint32_t)(c + u)) * 32u + lane : t + (uint32_t)(c + u) * THREADS);
				const uint32_t ox = o[u] % EX, oy = (o[u] / EX) % EY, oz = o[u] / (EX * EY);
				g[u][0] = ti[0] * EX + ox; g[u][1] = ti[1] * EY + oy; g[u][2] = (D == 3 ? ti[2] * (G::TZ / 2) + oz : 0u);
				ok[u] = g[u][0] < W1 && g[u][1] < H1 && (D < 3 || g[u][2] < D1);
				uint32_t s[3][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 } };
				wt[u][0] = fix[0].t;
				if constexpr (D == 2) {
					const axis_f f = axtab[1][EY + oy]; // the same row for the whole warp: a broadcast read
					wt[u][1] = f.t; s[1][0] = f.a; s[1][1] = f.b;
				} else {
					wt[u][1] = fix[1].t; s[1][0] = fix[1].a; s[1][1] = fix[1].b;
					const axis_f f = axtab[2][G::TZ / 2 + oz];
					wt[u][D - 1] = f.t; s[2][0] = f.a; s[2][1] = f.b;
				}
				if (ok[u]) {
#pragma unroll
					for (int r = 0; r < (1 << (D - 1)); ++r) {
						// one row pointer per (y, z) choice, two texels (A, B along x) from it
						const uint64_t row = (D == 3 ? (uint64_t)s[2][(r >> 1) & 1] * H0 : 0ull) + s[1][r & 1];
						const uint8_t* const rp = src + row * pitch;
						IO::template load<false>(rp + xoff[0], raw[u][2 * r]);
						IO::template load<false>(rp + xoff[1], raw[u][2 * r + 1]);
					}
				}
			}
#pragma unroll
			for (int u = 0; u < U; ++u) {
				if (ok[u]) {
					axis_f af[D];
#pragma unroll
					for (int d = 0; d < D; ++d) af[d].t = wt[u][d];
					uint32_t out[NW];
					tile_reduce_block<EK, CH, D, NW>(raw[u], af, P.no_double, out);
					IO::store(dst + (((uint64_t)g[u][2] * H1 + g[u][1]) * W1 + g[u][0]) * BPP, out);
#pragma unroll
					for (int w = 0; w < NW; ++w) buf0[o[u] * NW + w] = out[w];
				}
			}
		}
	}

	// ---- levels 2 .. nlev: shared memory ping-pong, every level also written to global ----------------------------
	// one output texel of level k: tile-local coordinates (ox, oy, oz), slot `oi` of the level's buffer
	auto produce = [&](uint32_t k, const uint32_t* in, uint32_t* outb, uint32_t ox, uint32_t oy, uint32_t oz, uint32_t oi) {
		const uint32_t ex = G::TX >> k, ey = G::TY >> k, ez = (D == 3 ? G::TZ >> k : 1u); // extents of level k of the tile
		const uint32_t g[3] = { ti[0] * ex + ox, ti[1] * ey + oy, (D == 3 ? ti[2] * ez + oz : 0u) };
		if (g[0] < P.dim[k][0] && g[1] < P.dim[k][1] && (D < 3 || g[2] < P.dim[k][2])) {
			const uint32_t pe[3] = { 2u * ex, 2u * ey, 2u * ez }; // extents of level k - 1 of the tile
			const uint32_t ext[3] = { ex, ey, ez }, oo[3] = { ox, oy, oz };
			axis_f af[D];
			uint32_t s[3][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 } };
#pragma unroll
			for (int d = 0; d < D; ++d) {
				af[d] = axtab[d][ext[d] + oo[d]]; // A, B: tile-local indices in level k - 1
				s[d][0] = af[d].a;
				s[d][1] = af[d].b;
			}
			uint32_t raw[1 << D][NW];
#pragma unroll
			for (int b = 0; b < (1 << D); ++b) {
				const uint32_t idx = ((s[2][(b >> 2) & 1] * pe[1] + s[1][(b >> 1) & 1]) * pe[0] + s[0][b & 1]) * NW;
#pragma unroll
				for (int w = 0; w < NW; ++w) raw[b][w] = in[idx + w];
			}
			uint32_t out[NW];
			tile_reduce_block<EK, CH, D, NW>(raw, af, P.no_double, out);
			uint8_t* const dst = base + P.level_off[k] + (uint64_t)layer * P.slice[k];
			IO::store(dst + (((uint64_t)g[2] * P.dim[k][1] + g[1]) * P.dim[k][0] + g[0]) * BPP, out);
#pragma unroll
			for (int w = 0; w < NW; ++w) outb[oi * NW + w] = out[w];
		}
	};

	if constexpr (D == 2) {
		// Levels 2 and 3 read only what the same warp wrote (warp w: rows 4w..4w+3 of level 1, rows 2w, 2w+1 of level 2, row w
		// of level 3), so a warp barrier is enough -- unless the launch contains a texel-2 fetch (P.block_sync), which may
		// reach into the next warp's rows.  Levels 4 .. 6 (16 + 4 + 1 texels) are finished by warp 0 alone.
		const uint32_t nlev = P.nlev;
		const bool bs = P.block_sync != 0u;
		if (nlev < 2u) return;
		if (bs) __syncthreads(); else __syncwarp();
		produce(2u, buf0, buf1, t & 15u, t >> 4, 0u, t);           // 16 x 16, slot = row-major index = t
		if (nlev < 3u) return;
		if (bs) __syncthreads(); else __syncwarp();
		if (lane < 8u) produce(3u, buf1, buf2, lane, warp, 0u, warp * 8u + lane); // 8 x 8, row w by warp w
		if (nlev < 4u) return;
		__syncthreads();
		if (warp != 0u) return;
		if (lane < 16u) produce(4u, buf2, buf1, lane & 3u, lane >> 2, 0u, lane);
		if (nlev < 5u) return;
		__syncwarp();
		if (lane < 4u) produce(5u, buf1, buf2, lane & 1u, lane >> 1, 0u, lane);
		if (nlev < 6u) return;
		__syncwarp();
		if (lane == 0u) produce(6u, buf2, buf1, 0u, 0u, 0u, 0u);
	} else {
#pragma unroll 1
		for (uint32_t k = 2; k <= P.nlev; ++k) {
			__syncthreads();
			const uint32_t ex = G::TX >> k, ey = G::TY >> k, ez = G::TZ >> k;
			if (t < ex * ey * ez) produce(k, (k & 1u) ? buf1 : buf0, (k & 1u) ? buf0 : buf1, t % ex, (t / ex) % ey, t / (ex * ey), t);
		}
	}
}

// ------------------------------------------------------------------------------------------------------
// persistent TMA tile kernel (flmip_ptile2d_*): any 2D image whose source rows are 16-byte multiples.
//
// The CTA structure of the single-pass kernel (TMA ring, producer warp with the dynamic scheduler, 8 consumer warps that hold
// a 512 B x 64 row tile in registers and produce levels 1 and 2 from it, a pool of finisher warps for the rest of the tile,
// the last tile of a layer finishes the chain in the same launch) with the reference's GENERAL sampler arithmetic
// (host_image.hpp:869-894): along each axis the fetch for destination texel g reads texels {2g, 2g + 1} of the previous level
// in the roles A (neighbour) and B (active texel) with a weight t = 0.5 +- eps of B, out = (B - A) * t + A, x first, then y.
// Roles and weights come from the image's sampler table (flmip_ptile_params), computed once per image on the host with the
// reference's float32 operations.  Partial tiles at the image border: TMA fills what lies outside with zeros, stores are masked.
// Levels 1 and 2 (in registers) assume A, B in {2g, 2g + 1}; the host only selects this kernel when the texel-2 fetch of the
// reference (see axis_fetch) cannot occur there.  From level 3 on texels are addressed by index and the fetch is honoured.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float wt_weight(uint32_t e) { return __uint_as_float(e & 0x3FFFFFFFu); }
__device__ __forceinline__ uint32_t wt_swap(uint32_t e) { return e >> 31; }

// (b - a) * t + a on pairs, product and sum rounded separately (t is not 0.5: the product is inexact)
__device__ __forceinline__ f32x2_t lerp2_t(f32x2_t a, f32x2_t b, f32x2_t t) { return add2_rn(mul2_sep(sub2_rn(b, a), t), a); }

// Puts texel A of every x pair of a 16-byte chunk into the even slot and texel B into the odd one (e[i]: table entry of pair i).
template <int BPP> __device__ __forceinline__ void preswap_chunk(uint32_t (&w)[4], const uint32_t (&e)[8 / BPP]) {
	static_assert(BPP <= 8, "16-byte texels swap whole chunks");
	if constexpr (BPP == 8) {
		const bool s = (int)e[0] < 0;
		const uint32_t a0 = s ? w[2] : w[0], a1 = s ? w[3] : w[1], b0 = s ? w[0] : w[2], b1 = s ? w[1] : w[3];
		w[0] = a0; w[1] = a1; w[2] = b0; w[3] = b1;
	} else if constexpr (BPP == 4) {
#pragma unroll
		for (int p = 0; p < 2; ++p) {
			const bool s = (int)e[p] < 0;
			const uint32_t a = s ? w[2 * p + 1] : w[2 * p], b = s ? w[2 * p] : w[2 * p + 1];
			w[2 * p] = a; w[2 * p + 1] = b;
		}
	} else if constexpr (BPP == 2) {
#pragma unroll
		for (int p = 0; p < 4; ++p) w[p] = __byte_perm(w[p], 0u, (int)e[p] < 0 ? 0x1032u : 0x3210u);
	} else {
#pragma unroll
		for (int p = 0; p < 4; ++p) {
			const uint32_t selector = 0x3210u ^ ((int)e[2 * p] < 0 ? 0x0011u : 0u) ^ ((int)e[2 * p + 1] < 0 ? 0x1100u : 0u);
			w[p] = __byte_perm(w[p], 0u, selector);
		}
	}
}

// rows rA / rB (NW words of whole texels, A in the even slot of every x pair) -> NW / 2 words of the next level;
// tx[j] = weight of output texel j along x, ty = weight along y
template <uint32_t EK, int CH, int NW>
__device__ __forceinline__ void reduce_rows_2d_t(const uint32_t (&rA)[NW], const uint32_t (&rB)[NW],
												 const float (&tx)[(NW * 2 / Codec<EK>::BYTES) / CH > 0 ? (NW * 2 / Codec<EK>::BYTES) / CH : 1], float ty,
												 uint32_t (&out)[NW / 2], uint32_t no_double) {
	using C = Codec<EK>;
	constexpr int NO = NW * 2 / C::BYTES;
	static_assert(NO >= CH && (NO % CH) == 0, "row must hold at least one x pair");
	if constexpr (!C::IS_INT) {
		static_assert((NO % 2) == 0, "pairs of output elements");
		f32x2_t v[NO / 2];
		const f32x2_t ty2 = splat2(ty);
#pragma unroll
		for (int p = 0; p < NO / 2; ++p) {
			const int e0 = 2 * p, e1 = 2 * p + 1;
			const int ea0 = (2 * (e0 / CH)) * CH + (e0 % CH), eb0 = ea0 + CH, ea1 = (2 * (e1 / CH)) * CH + (e1 % CH), eb1 = ea1 + CH;
			const f32x2_t tx2 = pk2(__float_as_uint(tx[e0 / CH]), __float_as_uint(tx[e1 / CH]));
			const f32x2_t x0 = lerp2_t(C::dec2_at(rA, ea0, ea1), C::dec2_at(rA, eb0, eb1), tx2);
			const f32x2_t x1 = lerp2_t(C::dec2_at(rB, ea0, ea1), C::dec2_at(rB, eb0, eb1), tx2);
			v[p] = lerp2_t(x0, x1, ty2);
		}
		C::template enc_pack2<NO / 2>(v, out, no_double);
	} else {
		uint32_t v[NO];
#pragma unroll
		for (int e = 0; e < NO; ++e) {
			const int ea = (2 * (e / CH)) * CH + (e % CH), eb = ea + CH;
			const uint32_t x0 = C::lerp_t(C::dec_at(rA, ea), C::dec_at(rA, eb), tx[e / CH]);
			const uint32_t x1 = C::lerp_t(C::dec_at(rB, ea), C::dec_at(rB, eb), tx[e / CH]);
			v[e] = C::lerp_t(x0, x1, ty);
		}
		C::template enc_pack<NO>(v, out, no_double);
	}
}

struct PRegion {
	uint32_t w, h;   // nominal texels of the region at level `lvl` (what lies outside the image is garbage and never read)
	uint32_t ox, oy; // origin in level-`lvl` texel coordinates
	uint32_t lvl;
};

template <int BPP> __device__ __forceinline__ uint8_t* plevel_layer_ptr(const flmip_ptile_params& P, uint32_t level, uint32_t layer) {
	return reinterpret_cast<uint8_t*>(P.base) + P.level_off[level] + (uint64_t)layer * ((uint64_t)P.dim[level][0] * P.dim[level][1] * BPP);
}

// One full warp: reduces the dense region `src` (row pitch R.w texels) level by level up to `stop_level`, or until the region
// cannot be halved any more; every level is written to global memory and re-read from its stored bits.  Texels are fetched by
// index, so the texel-2 fetch of the reference is honoured (the host guarantees it stays inside the region).
// The sampler entries are staged in `scratch` first (a dependent round trip to L2 per texel would serialise the warp):
// TILE_STAGE (region of a tile, power-of-two extents): the entries of ALL levels with one batch of loads -- needs R.w + R.h words;
// otherwise (whole level of a layer, any extents): level by level, as far as `cap` words reach.
template <uint32_t EK, int CH, bool TILE_STAGE>
__device__ __forceinline__ void pcascade_warp(uint8_t*& src, uint8_t*& dst, PRegion& R, const flmip_ptile_params& P, uint32_t layer, uint32_t lane,
											  uint32_t stop_level, uint32_t* scratch, uint32_t cap) {
	using C = Codec<EK>;
	constexpr int BPP = C::BYTES * CH;
	using IO = TexelIO<BPP>;
	constexpr int NW = IO::NW;
	const uint32_t* const wtab = reinterpret_cast<const uint32_t*>(P.wtab);
	if constexpr (TILE_STAGE) {
		uint32_t w = R.w, h = R.h, lvl = R.lvl, ox = R.ox, oy = R.oy, bx = 0, by = R.w;
		while (lvl < stop_level && w >= 2 && h >= 2) {
			w >>= 1; h >>= 1; ox >>= 1; oy >>= 1; ++lvl;
			// (entries past the level's extent exist: every table segment is padded by more than a tile)
#if defined(FLMIP_PT_EXP_FIN) && FLMIP_PT_EXP_FIN == 2
			for (uint32_t i = lane; i < w; i += 32) scratch[bx + i] = 0x3F000000u;
			for (uint32_t i = lane; i < h; i += 32) scratch[by + i] = 0x3F000000u;
#else
			for (uint32_t i = lane; i < w; i += 32) scratch[bx + i] = __ldg(wtab + P.wtab_off[lvl][0] + ox + i);
			for (uint32_t i = lane; i < h; i += 32) scratch[by + i] = __ldg(wtab + P.wtab_off[lvl][1] + oy + i);
#endif
			bx += w; by += h;
		}
		__syncwarp();
	}
	uint32_t bx = 0, by = R.w;
	while (R.lvl < stop_level && R.w >= 2 && R.h >= 2) {
		const uint32_t L = R.lvl + 1;
		const uint32_t dw = R.w >> 1, dh = R.h >> 1;
		const uint32_t LW = P.dim[L][0], LH = P.dim[L][1];
		const uint32_t ox = R.ox >> 1, oy = R.oy >> 1;
		uint8_t* const gdst = plevel_layer_ptr<BPP>(P, L, layer);
		const uint32_t* const tabx = wtab + P.wtab_off[L][0] + ox;
		const uint32_t* const taby = wtab + P.wtab_off[L][1] + oy;
		bool staged = TILE_STAGE;
		if constexpr (!TILE_STAGE) {
			staged = dw + dh <= cap;
			if (staged) {
				bx = 0; by = dw;
				for (uint32_t i = lane; i < dw; i += 32) scratch[i] = __ldg(tabx + i);
				for (uint32_t i = lane; i < dh; i += 32) scratch[dw + i] = __ldg(taby + i);
				__syncwarp();
			}
		}
		const uint32_t sw = 31u - __clz(dw);
		for (uint32_t i = lane; i < dw * dh; i += 32) {
			uint32_t x, y;
			if constexpr (TILE_STAGE) { x = i & (dw - 1u); y = i >> sw; }
			else { y = i / dw; x = i - y * dw; }
			const uint32_t gx = ox + x, gy = oy + y;
			if (gx < LW && gy < LH) {
				const uint32_t ex = staged ? scratch[bx + x] : __ldg(tabx + x), ey = staged ? scratch[by + y] : __ldg(taby + y);
				// texel indices inside the region: A / B = {2g, 2g + 1} in either order, or (2, 0) for the texel-2 fetch at g = 0
				const uint32_t lx[2] = { 2u * x + ((ex & FLMIP_WTAB_TEXEL2) ? 2u : wt_swap(ex)), 2u * x + ((ex & FLMIP_WTAB_TEXEL2) ? 0u : 1u - wt_swap(ex)) };
				const uint32_t ly[2] = { 2u * y + ((ey & FLMIP_WTAB_TEXEL2) ? 2u : wt_swap(ey)), 2u * y + ((ey & FLMIP_WTAB_TEXEL2) ? 0u : 1u - wt_swap(ey)) };
				uint32_t raw[4][NW];
#pragma unroll
				for (int b = 0; b < 4; ++b) IO::template load<false>(src + ((size_t)ly[b >> 1] * R.w + lx[b & 1]) * BPP, raw[b]);
				axis_f af[2];
				af[0].t = wt_weight(ex);
				af[1].t = wt_weight(ey);
				uint32_t out[NW];
				tile_reduce_block<EK, CH, 2, NW>(raw, af, P.no_double, out);
				IO::store(dst + (size_t)i * BPP, out);
				IO::store(gdst + ((uint64_t)gy * LW + gx) * BPP, out);
			}
		}
		__syncwarp();
		uint8_t* t = src; src = dst; dst = t;
		R.w = dw; R.h = dh; R.ox = ox; R.oy = oy; R.lvl = L;
		if constexpr (TILE_STAGE) { bx += dw; by += dh; }
	}
}

// tile coordinates that ride along with a stage / a cascade slot
struct PTile {
	uint32_t x, y, layer, pad;
};

template <uint32_t EK, int CH>
__device__ __forceinline__ void ptile_body(const CUtensorMap& tmap, const flmip_ptile_params& P) {
	using C = Codec<EK>;
	constexpr int BPP = C::BYTES * CH;
	using TL = flmip_tiling<BPP, 2>;
	using IO = TexelIO<BPP>;
	constexpr bool WIDE = (BPP == 16);
	constexpr int ROW_BYTES = TL::TILE_BYTES_X;
	constexpr int NPAIR = WIDE ? 1 : 8 / BPP;          // x pairs per 16-byte chunk (16-byte texels: one pair per thread)
	static_assert(TL::THREADS == FLMIP_CONSUMER_THREADS, "one consumer thread per 32 B x 4 rows");

	// [stages x tile][FLMIP_FINISHER_WARPS cascade slots (buf_a, buf_b)][patch (gathered level of one layer)][patch / 4]
	extern __shared__ __align__(128) uint8_t smem_raw[];
	__shared__ uint64_t full_bar[FLMIP_MAX_STAGES], empty_bar[FLMIP_MAX_STAGES];
	__shared__ uint64_t slot_full[FLMIP_FINISHER_WARPS], slot_empty[FLMIP_FINISHER_WARPS];
	__shared__ uint32_t finisher_ticket, patch_lock, unit_count[2];
	__shared__ uint32_t stage_tile[FLMIP_MAX_STAGES], slot_tile[FLMIP_FINISHER_WARPS];
	__shared__ __align__(16) PTile stage_pos[FLMIP_MAX_STAGES], slot_pos[FLMIP_FINISHER_WARPS];
	__shared__ __align__(16) uint64_t stage_org[FLMIP_MAX_STAGES][2]; // the tile's part of level src + 1 / src + 2 in global memory
	// sampler entries staged by the finisher warps: tile stage (TX / 4 + TY / 4 <= 144 words per warp), last-tile stage (one level at a time)
	__shared__ uint32_t fin_tab[FLMIP_FINISHER_WARPS][160];
	__shared__ uint32_t tail_tab[FLMIP_PTILE_TAIL_TAB];
	static_assert(TL::TX / 4 + TL::TY / 4 <= 160, "tile-stage sampler entries");
	const uint32_t stages = P.stages;
	uint8_t* const cascade_base = smem_raw + (size_t)stages * TL::TILE_BYTES;
	constexpr uint32_t SLOT_BYTES = TL::CASCADE_BYTES + TL::CASCADE_BYTES / 4u;

	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t S = P.src_level;

	if (tid == 0) {
		finisher_ticket = 0;
		patch_lock = 0;
		unit_count[0] = unit_count[1] = 0;
		for (uint32_t s = 0; s < stages; ++s) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], FLMIP_CONSUMER_WARPS);
		}
		for (uint32_t f = 0; f < FLMIP_FINISHER_WARPS; ++f) {
			mbar_init(&slot_full[f], FLMIP_CONSUMER_WARPS);
			mbar_init(&slot_empty[f], 1);
		}
		fence_mbar_init();
	}
	__syncthreads();
	pdl_start(P.late_wait);
	if (tid == 0) FLMIP_STAMP(P, 0); // CTA may touch global memory from here

	// the finisher pool is idle when the consumers' in-register levels end the launch
#ifdef FLMIP_PT_EXP_NOFIN
	const bool need_finish = false; // tuning experiment: the chain stops after the consumers' two levels
#else
	const bool need_finish = P.last_level > S + 2u;
#endif
	const uint32_t tiles_per_layer = P.tiles[0] * P.tiles[1];

	if (warp >= FLMIP_FINISHER_WARP0) {
		// ---- finishers: levels src + 3 .. tile_last of a tile, then (last tile of a layer) the rest of the chain ----------
		if (!need_finish) return;
		uint8_t* const patch_a = cascade_base + FLMIP_FINISHER_WARPS * SLOT_BYTES;
		uint8_t* const patch_b = patch_a + FLMIP_PTILE_PATCH_BYTES;
		for (;;) {
			uint32_t n = 0;
			if (lane == 0) n = atomicAdd(&finisher_ticket, 1u);
			n = __shfl_sync(0xFFFFFFFFu, n, 0);
			const uint32_t slot = n % FLMIP_FINISHER_WARPS, use = n / FLMIP_FINISHER_WARPS;
			uint8_t* const buf_a = cascade_base + slot * SLOT_BYTES;
#ifdef FLMIP_FINISHER_SLEEP_WAIT
			mbar_wait_sleep(&slot_full[slot], use & 1u);
#else
			mbar_wait(&slot_full[slot], use & 1u);
#endif
			const uint32_t t = slot_tile[slot];
			if (t == FLMIP_NO_TILE) break;
			const PTile tc = slot_pos[slot];
			PRegion R;
			R.lvl = S + 2u;
			R.w = TL::TX >> 2; R.h = TL::TY >> 2;
			R.ox = tc.x * R.w; R.oy = tc.y * R.h;
			uint8_t *src = buf_a, *dst = buf_a + TL::CASCADE_BYTES;
#if !defined(FLMIP_PT_EXP_FIN) || FLMIP_PT_EXP_FIN != 1
			pcascade_warp<EK, CH, true>(src, dst, R, P, tc.layer, lane, P.tile_last, fin_tab[warp - FLMIP_FINISHER_WARP0], 160u);
#endif
			__syncwarp();
#ifdef FLMIP_PT_EXP_FIN
			// tuning experiments: 1 = the slot is handed back at once, 2 / 3 = tile stage only (constant / real sampler entries), no publish
			if (lane == 0) mbar_arrive(&slot_empty[slot]);
			continue;
#endif
			// Publishing a tile costs a release fence + an atomic round trip (microseconds under load): with many tiles per CTA the pool
			// cannot keep up (N2: 0.32 ms against 0.20 ms without finishers).  The tiles of a unit (4 consecutive tile indices, always
			// worked through by one CTA in ring order) therefore meet at a counter in shared memory and whoever brings the last one
			// publishes all of them at once.  Two counters suffice: a slot is only released after its tile has been counted, and
			// FLMIP_FINISHER_WARPS <= tiles per unit (see fast_body).
			bool publish = true;
			uint32_t t0 = t, n_pub = 1u;
			if (P.unit_shift != 0u && P.last_level > P.tile_last) {
				const uint32_t U = 1u << P.unit_shift;
				t0 = t & ~(U - 1u);
				n_pub = min(U, P.total_tiles - t0);
				const uint32_t q = (n >> P.unit_shift) & 1u;
				uint32_t last = 0;
				if (lane == 0) {
					__threadfence_block();
					last = (atomicAdd(&unit_count[q], 1u) == n_pub - 1u);
					if (last) {
						unit_count[q] = 0u;
						__threadfence_block();
					}
				}
				publish = __shfl_sync(0xFFFFFFFFu, last, 0) != 0;
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&slot_empty[slot]); // the slot is free again before the slow part starts
			if (P.last_level <= P.tile_last || !publish) continue;
			// the last tile of a layer gathers level tile_last of the layer and finishes the chain (a unit may span layers)
			const uint32_t layer_a = t0 / tiles_per_layer, layer_b = (t0 + n_pub - 1u) / tiles_per_layer;
			for (uint32_t layer = layer_a; layer <= layer_b; ++layer) {
				const uint32_t lo = max(t0, layer * tiles_per_layer), hi = min(t0 + n_pub, (layer + 1u) * tiles_per_layer);
				if (!arrive_last_n(reinterpret_cast<uint32_t*>(P.counters) + layer, hi - lo, tiles_per_layer, lane)) continue;
				if (lane == 0) {
					while (atomicCAS(&patch_lock, 0u, 1u) != 0u) __nanosleep(64);
				}
				__syncwarp();
				PRegion G;
				G.lvl = P.tile_last;
				G.w = P.dim[G.lvl][0]; G.h = P.dim[G.lvl][1];
				G.ox = G.oy = 0;
				// one layer of a level is contiguous in global memory: a linear copy, loads batched (a round trip to L2 costs microseconds
				// under load): 16 bytes per lane and load where the layer starts on a 16-byte boundary, 8 loads in flight
				{
					constexpr uint32_t U = (IO::NW >= 4 ? 2u : (IO::NW == 2 ? 4u : 8u));
					const uint8_t* const g = plevel_layer_ptr<BPP>(P, G.lvl, layer);
					const uint32_t ntex = G.w * G.h;
					uint32_t first = 0; // texels already copied by the vector part
					if ((reinterpret_cast<uint64_t>(g) & 15u) == 0u) {
						const uint32_t n16 = (uint32_t)(((uint64_t)ntex * BPP) >> 4);
						for (uint32_t b0 = 0; b0 < n16; b0 += 32u * 8u) {
							uint4 v[8];
#pragma unroll
							for (uint32_t u = 0; u < 8u; ++u) {
								const uint32_t i = b0 + u * 32u + lane;
								if (i < n16) v[u] = __ldcg(reinterpret_cast<const uint4*>(g) + i);
							}
#pragma unroll
							for (uint32_t u = 0; u < 8u; ++u) {
								const uint32_t i = b0 + u * 32u + lane;
								if (i < n16) reinterpret_cast<uint4*>(patch_a)[i] = v[u];
							}
						}
						first = (uint32_t)(((uint64_t)n16 << 4) / BPP);
					}
					for (uint32_t b0 = first; b0 < ntex; b0 += 32u * U) {
						uint32_t t[U][IO::NW];
#pragma unroll
						for (uint32_t u = 0; u < U; ++u) {
							const uint32_t i = b0 + u * 32u + lane;
							if (i < ntex) IO::template load<true>(g + (size_t)i * BPP, t[u]);
						}
#pragma unroll
						for (uint32_t u = 0; u < U; ++u) {
							const uint32_t i = b0 + u * 32u + lane;
							if (i < ntex) IO::store(patch_a + (size_t)i * BPP, t[u]);
						}
					}
					__syncwarp();
				}
				uint8_t *ps = patch_a, *pd = patch_b;
				pcascade_warp<EK, CH, false>(ps, pd, G, P, layer, lane, P.last_level, tail_tab, FLMIP_PTILE_TAIL_TAB);
				__syncwarp();
				if (lane == 0) {
					__threadfence_block();
					atomicExch(&patch_lock, 0u);
				}
			}
		}
		if (lane == 0) FLMIP_STAMP_MAX(P, 3); // last finisher warp of the CTA done
		return;
	}
	if (warp == FLMIP_PRODUCER_WARP) {
		// ---- producer: dynamic tile scheduler + TMA issue (see fast_body) -------------------------------------------------
		const uint32_t PF = FLMIP_SCHED_PREFETCH;
		uint32_t* const sched = reinterpret_cast<uint32_t*>(P.sched);
		uint32_t pf = FLMIP_NO_TILE;
		if (lane == 0) pf = blockIdx.x;
		else if (lane < PF) pf = gridDim.x + atomicAdd(&sched[0], 1u);
		uint32_t s = 0, use = 0, dead = 0;
		const uint64_t p1 = (uint64_t)P.dim[S + 1u][0] * BPP;
		const uint64_t p2 = (P.last_level >= S + 2u) ? (uint64_t)P.dim[S + 2u][0] * BPP : 0ull;
		const uint32_t U = 1u << P.unit_shift, total_units = (P.total_tiles + U - 1u) >> P.unit_shift;
		for (uint32_t it = 0; dead < PF; ++it) {
			const uint32_t u = __shfl_sync(0xFFFFFFFFu, pf, it % PF);
			if (u >= total_units) { ++dead; continue; }
			dead = 0;
			if (lane == it % PF) pf = gridDim.x + atomicAdd(&sched[0], 1u);
			for (uint32_t k = 0; k < U; ++k) {
			const uint32_t t = (u << P.unit_shift) + k;
			if (t >= P.total_tiles) break;
			if (lane == 0) {
				if (use > 0) mbar_wait(&empty_bar[s], (use - 1u) & 1u);
				PTile tc;
				tc.x = t % P.tiles[0];
				const uint32_t r = t / P.tiles[0];
				tc.y = r % P.tiles[1];
				tc.layer = r / P.tiles[1];
				tc.pad = 0;
				mbar_expect_tx(&full_bar[s], TL::TILE_BYTES);
				tma_load_3d(smem_raw + (size_t)s * TL::TILE_BYTES, &tmap, &full_bar[s], (int)(tc.x * (ROW_BYTES / 4)), (int)(tc.y * TL::TY), (int)tc.layer);
				stage_tile[s] = t;
				stage_pos[s] = tc;
				stage_org[s][0] = reinterpret_cast<uint64_t>(plevel_layer_ptr<BPP>(P, S + 1u, tc.layer)) + (uint64_t)tc.y * (TL::TY / 2) * p1 + (uint64_t)tc.x * (ROW_BYTES / 2);
				stage_org[s][1] = (P.last_level >= S + 2u)
									  ? reinterpret_cast<uint64_t>(plevel_layer_ptr<BPP>(P, S + 2u, tc.layer)) + (uint64_t)tc.y * (TL::TY / 4) * p2 + (uint64_t)tc.x * (ROW_BYTES / 4)
									  : 0ull;
				mbar_arrive(&full_bar[s]);
			}
			if (++s == stages) { s = 0; ++use; }
			}
			__syncwarp();
		}
		if (lane == 0) {
			FLMIP_STAMP(P, 1); // the scheduler ran dry for this CTA: all of its loads are issued
			if (use > 0) mbar_wait(&empty_bar[s], (use - 1u) & 1u);
			stage_tile[s] = FLMIP_NO_TILE;
			mbar_arrive(&full_bar[s]);
			__threadfence();
			if (atomicAdd(&sched[1], 1u) == gridDim.x - 1u) {
				sched[0] = 0u;
				sched[1] = 0u;
			}
		}
		return;
	}

	// ---- consumers: thread = 2 chunks (32 B) x 4 rows of the tile -> 16 B x 2 rows of level 1 -> 8 B of level 2 -----------
	const uint32_t* const wtab = reinterpret_cast<const uint32_t*>(P.wtab);
	const uint32_t* const tab1x = wtab + P.wtab_off[S + 1u][0];
	const uint32_t* const tab1y = wtab + P.wtab_off[S + 1u][1];
	const bool has_l2 = P.last_level >= S + 2u;
	const uint32_t* const tab2x = wtab + P.wtab_off[has_l2 ? S + 2u : S + 1u][0];
	const uint32_t* const tab2y = wtab + P.wtab_off[has_l2 ? S + 2u : S + 1u][1];
	const uint32_t W1 = P.dim[S + 1u][0], H1 = P.dim[S + 1u][1];
	const uint32_t W2 = has_l2 ? P.dim[S + 2u][0] : 0u, H2 = has_l2 ? P.dim[S + 2u][1] : 0u;
	const uint64_t pitch1 = (uint64_t)W1 * BPP, pitch2 = (uint64_t)W2 * BPP;
	const bool vec1 = P.vec1 != 0u; // 16-byte stores of level 1 (else two of 8 bytes: rows start on 8-byte boundaries)
	const bool vec2 = P.vec2 != 0u; // 8-byte stores of level 2 (else two of 4 bytes)
	const uint32_t sel = (lane >> 2) & 1u; // quarter-warps read conflict-free by swapping the chunk order on lane bit 2
	const uint32_t tx = tid % TL::THREADS_X, ty = tid / TL::THREADS_X;
	uint32_t s = 0, use = 0, slot = 0, slot_use = 0;
	for (;;) {
		mbar_wait(&full_bar[s], use & 1u);
		const uint32_t t = stage_tile[s];
		if (t == FLMIP_NO_TILE) {
			if (tid == 0) FLMIP_STAMP(P, 2); // consumers saw the sentinel
			break;
		}
		const uint8_t* const tile = smem_raw + (size_t)s * TL::TILE_BYTES;
		uint8_t* const buf_a = cascade_base + slot * SLOT_BYTES;
		const PTile tc = stage_pos[s];
		const ulonglong2 org = *reinterpret_cast<const ulonglong2*>(stage_org[s]);
		uint8_t* const g1 = reinterpret_cast<uint8_t*>(org.x);
		uint8_t* const g2 = reinterpret_cast<uint8_t*>(org.y);

		// sampler entries of this thread's texels (L1-resident: every thread of a tile column / row reads the same ones).
		// Level 1: rows 2 ty, 2 ty + 1; columns of the two chunks the thread reads (physical order k ^ sel).  Level 2: row ty.
		uint32_t ey1[2], ex1[2][NPAIR], ey2 = 0, ex2[NPAIR];
#ifdef FLMIP_PT_EXP_NOTAB
		// tuning experiment: constant entries instead of the table
		ey1[0] = ey1[1] = ey2 = 0x3F000000u;
#pragma unroll
		for (int i = 0; i < NPAIR; ++i) ex1[0][i] = ex1[1][i] = ex2[i] = 0x3F000000u;
		if (false)
#endif
		{
			const uint32_t gy1 = tc.y * (TL::TY / 2) + 2u * ty;
			ey1[0] = __ldg(tab1y + gy1);
			ey1[1] = __ldg(tab1y + gy1 + 1u);
			if constexpr (!WIDE) {
#pragma unroll
				for (int k = 0; k < 2; ++k)
#pragma unroll
					for (int i = 0; i < NPAIR; ++i) ex1[k][i] = __ldg(tab1x + tc.x * (TL::TX / 2) + (2u * tx + ((uint32_t)k ^ sel)) * NPAIR + i);
#pragma unroll
				for (int i = 0; i < NPAIR; ++i) ex2[i] = has_l2 ? __ldg(tab2x + tc.x * (TL::TX / 4) + tx * NPAIR + i) : 0u;
			} else {
				ex1[0][0] = ex1[1][0] = __ldg(tab1x + tc.x * (TL::TX / 2) + tx);
				ex2[0] = has_l2 ? __ldg(tab2x + tc.x * (TL::TX / 4) + (tx >> 1)) : 0u;
			}
			if (has_l2) ey2 = __ldg(tab2y + tc.y * (TL::TY / 4) + ty);
		}

		// rows in their roles: level-1 row j reads tile rows 4 ty + 2 j + {0, 1}; A is the second one when the roles are swapped
		uint32_t rawA[2][2][4], rawB[2][2][4]; // [level-1 row][chunk][word]
#pragma unroll
		for (int j = 0; j < 2; ++j) {
			const uint32_t sy = wt_swap(ey1[j]);
			const uint8_t* const rowA = tile + (4u * ty + 2u * j + sy) * ROW_BYTES;
			const uint8_t* const rowB = tile + (4u * ty + 2u * j + 1u - sy) * ROW_BYTES;
#pragma unroll
			for (int k = 0; k < 2; ++k) {
				const uint4 a = *reinterpret_cast<const uint4*>(rowA + (2 * tx + (k ^ sel)) * 16);
				const uint4 b = *reinterpret_cast<const uint4*>(rowB + (2 * tx + (k ^ sel)) * 16);
				rawA[j][k][0] = a.x; rawA[j][k][1] = a.y; rawA[j][k][2] = a.z; rawA[j][k][3] = a.w;
				rawB[j][k][0] = b.x; rawB[j][k][1] = b.y; rawB[j][k][2] = b.z; rawB[j][k][3] = b.w;
			}
		}
		__syncwarp();
		if (lane == 0) mbar_arrive(&empty_bar[s]); // the tile lives in registers: hand the stage back

		uint32_t l1[2][4]; // two level-1 rows of 16 bytes, logical (left, right) order
		if constexpr (!WIDE) {
			float tx1[2][NPAIR];
			uint32_t any1 = 0;
#pragma unroll
			for (int k = 0; k < 2; ++k)
#pragma unroll
				for (int i = 0; i < NPAIR; ++i) {
					tx1[k][i] = wt_weight(ex1[k][i]);
					any1 |= ex1[k][i];
				}
			// most tiles of most sizes have no swapped x roles at all (1920: none, 1080: every fifth column): skip the selects then
#ifdef FLMIP_PT_NO_SELSKIP
			any1 = 0x80000000u;
#endif
			if (__any_sync(0xFFFFFFFFu, (int)any1 < 0)) {
#pragma unroll
				for (int j = 0; j < 2; ++j)
#pragma unroll
					for (int k = 0; k < 2; ++k) {
						preswap_chunk<BPP>(rawA[j][k], ex1[k]);
						preswap_chunk<BPP>(rawB[j][k], ex1[k]);
					}
			}
#pragma unroll
			for (int j = 0; j < 2; ++j) {
				uint32_t o[2][2];
#pragma unroll
				for (int k = 0; k < 2; ++k) {
					reduce_rows_2d_t<EK, CH, 4>(rawA[j][k], rawB[j][k], tx1[k], wt_weight(ey1[j]), o[k], P.no_double);
				}
				l1[j][0] = sel ? o[1][0] : o[0][0]; l1[j][1] = sel ? o[1][1] : o[0][1];
				l1[j][2] = sel ? o[0][0] : o[1][0]; l1[j][3] = sel ? o[0][1] : o[1][1];
			}
		} else {
			// 16-byte texels: the thread's two chunks are texels 2 tx and 2 tx + 1; chunk k holds texel k ^ sel, A is texel `swap`
			const uint32_t pick = sel ^ wt_swap(ex1[0][0]);
			const float tx1[1] = { wt_weight(ex1[0][0]) };
#pragma unroll
			for (int j = 0; j < 2; ++j) {
				uint32_t ra[8], rb[8];
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					ra[i] = pick ? rawA[j][1][i] : rawA[j][0][i];
					ra[4 + i] = pick ? rawA[j][0][i] : rawA[j][1][i];
					rb[i] = pick ? rawB[j][1][i] : rawB[j][0][i];
					rb[4 + i] = pick ? rawB[j][0][i] : rawB[j][1][i];
				}
				reduce_rows_2d_t<EK, CH, 8>(ra, rb, tx1, wt_weight(ey1[j]), l1[j], P.no_double);
			}
		}
		// level 1: 16 bytes per row, masked at the image border
		{
			const uint32_t xb = tc.x * (ROW_BYTES / 2) + tx * 16u; // byte offset in the level-1 row
			const uint32_t gy = tc.y * (TL::TY / 2) + 2u * ty;
#pragma unroll
			for (int j = 0; j < 2; ++j) {
				if (gy + j < H1 && xb < pitch1) {
					uint8_t* const p = g1 + (uint64_t)(2u * ty + j) * pitch1 + tx * 16u;
					if (vec1) {
						*reinterpret_cast<uint4*>(p) = make_uint4(l1[j][0], l1[j][1], l1[j][2], l1[j][3]);
					} else if constexpr (!WIDE) {
						*reinterpret_cast<uint2*>(p) = make_uint2(l1[j][0], l1[j][1]);
						if (xb + 8u < pitch1) *reinterpret_cast<uint2*>(p + 8) = make_uint2(l1[j][2], l1[j][3]);
					}
				}
			}
		}
		if (has_l2) {
			const float ty2 = wt_weight(ey2);
			const uint32_t sy2 = wt_swap(ey2);
			if constexpr (!WIDE) {
				uint32_t ra[4], rb[4], l2[2];
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					ra[i] = sy2 ? l1[1][i] : l1[0][i];
					rb[i] = sy2 ? l1[0][i] : l1[1][i];
				}
				float tx2[NPAIR];
				uint32_t any2 = 0;
#pragma unroll
				for (int i = 0; i < NPAIR; ++i) {
					tx2[i] = wt_weight(ex2[i]);
					any2 |= ex2[i];
				}
#ifdef FLMIP_PT_NO_SELSKIP
				any2 = 0x80000000u;
#endif
				if (__any_sync(0xFFFFFFFFu, (int)any2 < 0)) {
					preswap_chunk<BPP>(ra, ex2);
					preswap_chunk<BPP>(rb, ex2);
				}
				reduce_rows_2d_t<EK, CH, 4>(ra, rb, tx2, ty2, l2, P.no_double);
				const uint32_t xb = tc.x * (ROW_BYTES / 4) + tx * 8u;
				if (tc.y * (TL::TY / 4) + ty < H2 && xb < pitch2) {
					uint8_t* const p = g2 + (uint64_t)ty * pitch2 + tx * 8u;
					if (vec2) {
						*reinterpret_cast<uint2*>(p) = make_uint2(l2[0], l2[1]);
					} else {
						*reinterpret_cast<uint32_t*>(p) = l2[0];
						if (xb + 4u < pitch2) *reinterpret_cast<uint32_t*>(p + 4) = l2[1];
					}
				}
				if (need_finish) {
					if (slot_use > 0) mbar_wait(&slot_empty[slot], (slot_use - 1u) & 1u);
					*reinterpret_cast<uint2*>(buf_a + ty * (ROW_BYTES / 4) + tx * 8) = make_uint2(l2[0], l2[1]);
				}
			} else {
				// a thread holds one level-1 texel per row, its x neighbour lives in lane ^ 1; even lanes produce the level-2 texel
				uint32_t own[2][4], nb[2][4];
#pragma unroll
				for (int j = 0; j < 2; ++j)
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						own[j][i] = l1[j][i];
						nb[j][i] = __shfl_xor_sync(0xFFFFFFFFu, l1[j][i], 1);
					}
				if (!(lane & 1u)) {
					const uint32_t sx2 = wt_swap(ex2[0]);
					uint32_t ra[8], rb[8], l2[4];
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						// row roles first (A row = level-1 row sy2), then texel roles (A texel = the pair's texel sx2)
						const uint32_t a_own = sy2 ? own[1][i] : own[0][i], a_nb = sy2 ? nb[1][i] : nb[0][i];
						const uint32_t b_own = sy2 ? own[0][i] : own[1][i], b_nb = sy2 ? nb[0][i] : nb[1][i];
						ra[i] = sx2 ? a_nb : a_own; ra[4 + i] = sx2 ? a_own : a_nb;
						rb[i] = sx2 ? b_nb : b_own; rb[4 + i] = sx2 ? b_own : b_nb;
					}
					const float tx2[1] = { wt_weight(ex2[0]) };
					reduce_rows_2d_t<EK, CH, 8>(ra, rb, tx2, ty2, l2, P.no_double);
					const uint32_t xb = tc.x * (ROW_BYTES / 4) + (tx >> 1) * 16u;
					if (tc.y * (TL::TY / 4) + ty < H2 && xb < pitch2)
						*reinterpret_cast<uint4*>(g2 + (uint64_t)ty * pitch2 + (tx >> 1) * 16u) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
					if (need_finish) {
						if (slot_use > 0) mbar_wait(&slot_empty[slot], (slot_use - 1u) & 1u);
						*reinterpret_cast<uint4*>(buf_a + ty * (ROW_BYTES / 4) + (tx >> 1) * 16) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
					}
				}
			}
		}

		// this warp's part of buf_a is written: hand the slot to the finisher pool (arrive = release)
		if (need_finish) {
			if (tid == 0) {
				slot_tile[slot] = t; // only after this thread's wait on slot_empty above
				slot_pos[slot] = tc;
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&slot_full[slot]);
		}
		if (++s == stages) { s = 0; ++use; }
		if (++slot == FLMIP_FINISHER_WARPS) { slot = 0; ++slot_use; }
	}
	if (need_finish) {
		for (uint32_t k = 0; k < FLMIP_FINISHER_WARPS; ++k) {
			if (slot_use > 0) mbar_wait(&slot_empty[slot], (slot_use - 1u) & 1u);
			if (tid == 0) slot_tile[slot] = FLMIP_NO_TILE;
			__syncwarp();
			if (lane == 0) mbar_arrive(&slot_full[slot]);
			if (++slot == FLMIP_FINISHER_WARPS) { slot = 0; ++slot_use; }
		}
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
		case FLMIP_EK_UNORM4: generic_texel<FLMIP_EK_UNORM4>(P, idx); break;
		case FLMIP_EK_SNORM4: generic_texel<FLMIP_EK_SNORM4>(P, idx); break;
		case FLMIP_EK_UNORM2: generic_texel<FLMIP_EK_UNORM2>(P, idx); break;
		case FLMIP_EK_SNORM2: generic_texel<FLMIP_EK_SNORM2>(P, idx); break;
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
	extern "C" __global__ void __launch_bounds__(FLMIP_BLOCK_THREADS, 2) flmip_fast##D##d_k##K##_c##CHN(const __grid_constant__ CUtensorMap tmap,    \
																					   const __grid_constant__ flmip_fast_params P) { \
		fast_body<K, CHN, D>(tmap, P);                                                                                          \
		if (P.late_wait == 1u && blockIdx.x == 0 && threadIdx.x == 0) pdl_late_wait();                                             \
	}
#define FLMIP_FAST_KERNELS_FOR_KIND(K) \
	FLMIP_FAST_KERNEL(2, K, 1) FLMIP_FAST_KERNEL(2, K, 2) FLMIP_FAST_KERNEL(2, K, 4) FLMIP_FAST_KERNEL(3, K, 1) FLMIP_FAST_KERNEL(3, K, 2) FLMIP_FAST_KERNEL(3, K, 4)

// -DFLMIP_DEV_ONLY: a handful of instantiations for quick SASS inspection (never shipped)
#ifdef FLMIP_DEV_ONLY
FLMIP_FAST_KERNEL(2, 2, 4) FLMIP_FAST_KERNEL(2, 1, 4) FLMIP_FAST_KERNEL(3, 0, 1) FLMIP_FAST_KERNEL(2, 0, 4) FLMIP_FAST_KERNEL(2, 4, 4)
#else
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
#endif

#define FLMIP_PTILE_KERNEL(K, CHN)                                                                                                  \
	extern "C" __global__ void __launch_bounds__(FLMIP_BLOCK_THREADS, 2) flmip_ptile2d_k##K##_c##CHN(const __grid_constant__ CUtensorMap tmap,     \
																							   const __grid_constant__ flmip_ptile_params P) { \
		ptile_body<K, CHN>(tmap, P);                                                                                                \
		if (P.late_wait == 1u && blockIdx.x == 0 && threadIdx.x == 0) pdl_late_wait();                                                 \
	}
#define FLMIP_PTILE_KERNELS_FOR_KIND(K) FLMIP_PTILE_KERNEL(K, 1) FLMIP_PTILE_KERNEL(K, 2) FLMIP_PTILE_KERNEL(K, 4)
#ifdef FLMIP_DEV_ONLY
FLMIP_PTILE_KERNEL(2, 4) FLMIP_PTILE_KERNEL(1, 4) FLMIP_PTILE_KERNEL(0, 4) FLMIP_PTILE_KERNEL(2, 1) FLMIP_PTILE_KERNEL(10, 2)
#else
FLMIP_PTILE_KERNELS_FOR_KIND(0)
FLMIP_PTILE_KERNELS_FOR_KIND(1)
FLMIP_PTILE_KERNELS_FOR_KIND(2)
FLMIP_PTILE_KERNELS_FOR_KIND(3)
FLMIP_PTILE_KERNELS_FOR_KIND(4)
FLMIP_PTILE_KERNELS_FOR_KIND(5)
FLMIP_PTILE_KERNELS_FOR_KIND(6)
FLMIP_PTILE_KERNELS_FOR_KIND(7)
FLMIP_PTILE_KERNELS_FOR_KIND(8)
FLMIP_PTILE_KERNELS_FOR_KIND(9)
FLMIP_PTILE_KERNELS_FOR_KIND(10)
FLMIP_PTILE_KERNELS_FOR_KIND(11)
#endif

// 2D, texels below 16 bytes: 4 CTAs per SM (64 registers, no spills).  Measured on N2 (RGBA16F) with the per-CTA sampler table:
// 3 CTAs 5 175 GB/s, 4 CTAs 5 645, 5 CTAs (48 registers, 16 bytes of spills) 5 315, 6 CTAs 5 354 (profiles/r1/15_tile_sampler_table_ab.txt;
// before the table the kernel executed twice the instructions and 5 CTAs were best).  The same 4 CTAs (64 registers, no spills) for 16-byte
// texels (ptxas' own choice: 72 registers, 3 CTAs; NPOT RGBA32F layers +2 %) and for volumes (ptxas: 74 .. 90 registers, 2 CTAs; 500 x 300 x 200:
// R32F +10 %, RGBA8 +22 %, RGBA16F +30 %; scripts/tile_occ_ab.py, same log).
#ifndef FLMIP_TILE2D_MIN_BLOCKS
#define FLMIP_TILE2D_MIN_BLOCKS 4
#endif
#ifndef FLMIP_TILE2D_WIDE_MIN_BLOCKS
#define FLMIP_TILE2D_WIDE_MIN_BLOCKS 4
#endif
#ifndef FLMIP_TILE3D_MIN_BLOCKS
#define FLMIP_TILE3D_MIN_BLOCKS 4
#endif
#define FLMIP_TILE_MIN_BLOCKS(D, K, CHN) ((D) == 3 ? FLMIP_TILE3D_MIN_BLOCKS : (flmip_elem_bytes(K) * (CHN) < 16 ? FLMIP_TILE2D_MIN_BLOCKS : FLMIP_TILE2D_WIDE_MIN_BLOCKS))
#define FLMIP_TILE_KERNEL(D, K, CHN)                                                                                              \
	extern "C" __global__ void __launch_bounds__(256, FLMIP_TILE_MIN_BLOCKS(D, K, CHN)) flmip_tile##D##d_k##K##_c##CHN(const __grid_constant__ flmip_tile_params P) { \
		tile_body<K, CHN, D>(P);                                                                                                 \
		if (P.late_wait == 1u && blockIdx.x == gridDim.x - 1u && threadIdx.x == 0) pdl_late_wait();                                 \
	}
#define FLMIP_TILE_KERNELS_FOR_KIND(K) \
	FLMIP_TILE_KERNEL(2, K, 1) FLMIP_TILE_KERNEL(2, K, 2) FLMIP_TILE_KERNEL(2, K, 4) FLMIP_TILE_KERNEL(3, K, 1) FLMIP_TILE_KERNEL(3, K, 2) FLMIP_TILE_KERNEL(3, K, 4) \
	FLMIP_TILE_KERNEL(2, K, 3) FLMIP_TILE_KERNEL(3, K, 3) /* 3-channel images: Host-Compute minifies them, a CUarray cannot hold them (cuda_image.cpp:173-180) */
#ifdef FLMIP_DEV_ONLY
FLMIP_TILE_KERNEL(2, 2, 4) FLMIP_TILE_KERNEL(2, 1, 4) FLMIP_TILE_KERNEL(3, 0, 1) FLMIP_TILE_KERNEL(2, 0, 4) FLMIP_TILE_KERNEL(3, 1, 4) FLMIP_TILE_KERNEL(2, 2, 3) FLMIP_TILE_KERNEL(2, 1, 3) FLMIP_TILE_KERNEL(3, 0, 3)
#else
FLMIP_TILE_KERNELS_FOR_KIND(0)
FLMIP_TILE_KERNELS_FOR_KIND(1)
FLMIP_TILE_KERNELS_FOR_KIND(2)
FLMIP_TILE_KERNELS_FOR_KIND(3)
FLMIP_TILE_KERNELS_FOR_KIND(4)
FLMIP_TILE_KERNELS_FOR_KIND(5)
FLMIP_TILE_KERNELS_FOR_KIND(6)
FLMIP_TILE_KERNELS_FOR_KIND(7)
FLMIP_TILE_KERNELS_FOR_KIND(8)
FLMIP_TILE_KERNELS_FOR_KIND(9)
FLMIP_TILE_KERNELS_FOR_KIND(10)
FLMIP_TILE_KERNELS_FOR_KIND(11)
#endif
