/*
 * minify_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of libfloor's Host-Compute mip-map minification path, used as the parity
 * oracle and as the "port" CPU baseline of bench.py.  PINNED AGAINST THE REFERENCE ITSELF: the
 * reference ships no tests / golden vectors for this path and libfloor as a whole needs
 * clang >= 19, but its Host-Compute minify kernels + software sampler + image-size helpers do
 * compile with g++ through oracle/build_ref.py (-> oracle/_ref/); tests/test_reference_pin.py
 * requires this restatement to equal them bit for bit (all formats, 2D / array / cube / 3D,
 * POT and NPOT, both encoder-scale modes, adversarial floats, BASELINE configs), and
 * tests/golden/golden_ref.json freezes reference-computed chains.  All 17 kernels of
 * FLOOR_MINIFY_IMAGE_TYPES (1D, 1D-array, 2D, 2D-array, 3D x FLOAT / INT / UINT, depth, depth-array)
 * are covered.  Evaluated in strict IEEE-754 order (-fno-fast-math -ffp-contract=off).
 *
 * Each function cites the reference file:line (relative to a2flo/floor) it follows:
 *   kernel ................. include/floor/device/backend/mip_map_minify.hpp:78-108
 *   level/layer loop ....... src/device/device_image.cpp:290-327
 *   sampler + codecs ....... include/floor/device/backend/host_image.hpp:141-174, 235-271,
 *                            333-460, 487-561, 640-667, 672-722, 801-825, 842-929
 *   wrap/fractional/lerp ... include/floor/constexpr/const_math.hpp:308-313, 859-869, 981-996
 *   level table ............ src/device/host/host_image.cpp:81-108
 *   sizes/offsets .......... include/floor/device/backend/image_types.hpp:449-809
 *   execution model ........ src/device/host/host_function.cpp:734-1015 (work-group tickets)
 *
 * Deviations (documented): byte offsets are 64-bit (reference: uint32_t, host_image.hpp:44);
 * float->integer casts go through a wide signed type so that out-of-range inputs are
 * deterministic instead of UB; signed (b - a) wraps instead of being UB on overflow.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- IMAGE_TYPE bit layout (image_types.hpp:24-236) ------------------------------------- */
#define T_FORMAT_MASK 0x3Full
#define T_COMPRESSION_MASK 0x3C0ull
#define T_DATA_TYPE_MASK 0x3000ull
#define T_INT 0x1000ull
#define T_UINT 0x2000ull
#define T_FLOAT 0x3000ull
#define T_CHANNELS_SHIFT 14
#define T_DIM_SHIFT 16
#define T_FLAG_ARRAY (1ull << 20)
#define T_FLAG_MSAA (1ull << 22)
#define T_FLAG_CUBE (1ull << 23)
#define T_FLAG_DEPTH (1ull << 24)
#define T_FLAG_STENCIL (1ull << 25)
#define T_FLAG_MIPMAPPED (1ull << 27)
#define T_FLAG_NORMALIZED (1ull << 30)
#define FMT_2 2u
#define FMT_4 4u
#define FMT_8 11u
#define FMT_16 18u
#define FMT_32 22u

#define ORACLE_MAX_LEVELS 16 /* host_limits::max_mip_levels */

enum { FLO_OK = 0, FLO_ERR_UNSUPPORTED = -1, FLO_ERR_ARG = -2 };
/* flags for flo_generate_mip_map_chain */
#define FLO_FLAG_NO_DOUBLE 1u /* FLOOR_DEVICE_NO_DOUBLE variant (host_image.hpp:398-402) */

static uint32_t dim_count(uint64_t t) { return (uint32_t)((t >> T_DIM_SHIFT) & 3u); }
static uint32_t channel_count(uint64_t t) { return (uint32_t)((t >> T_CHANNELS_SHIFT) & 3u) + 1u; }
static uint32_t format_of(uint64_t t) { return (uint32_t)(t & T_FORMAT_MASK); }

/* image_types.hpp:541-556: bits per channel for the uniform formats this path supports */
static uint32_t bits_per_channel(uint64_t t) {
	switch (format_of(t)) {
		case FMT_2: return 2; /* FORMAT_2 / FORMAT_4: normalized formats only, channels packed into 1 - 2 bytes (host_image.hpp:341-353) */
		case FMT_4: return 4;
		case FMT_8: return 8;
		case FMT_16: return 16;
		case FMT_32: return 32;
		default: return 0;
	}
}

/* image_types.hpp:641-644 */
uint32_t flo_bytes_per_pixel(uint64_t t) { return (bits_per_channel(t) * channel_count(t) + 7u) / 8u; }

/* image_types.hpp:694-712 (prev_pot: const_math.hpp:1122-1136) */
uint32_t flo_mip_level_count(const uint32_t dim[4], uint64_t t) {
	if (!(t & T_FLAG_MIPMAPPED)) return 1;
	const uint32_t dc = dim_count(t);
	uint32_t m = dim[0];
	if (dc >= 2 && dim[1] > m) m = dim[1];
	if (dc >= 3 && dim[2] > m) m = dim[2];
	if (m <= 1) return 1;
	uint32_t pot = 1;
	while ((uint64_t)pot * 2u <= m) pot *= 2u;
	return 32u - (uint32_t)__builtin_clz(pot);
}

/* image_types.hpp:716-726 */
uint32_t flo_layer_count(const uint32_t dim[4], uint64_t t) {
	const uint32_t dc = dim_count(t);
	uint32_t n = !(t & T_FLAG_ARRAY) ? 1u : (dc == 1 ? dim[1] : (dc == 2 ? dim[2] : dim[3]));
	if (t & T_FLAG_CUBE) n *= 6u;
	return n;
}

/* image_types.hpp:675-691 (uncompressed, non-MSAA) */
static uint64_t slice_size(const uint32_t d[3], uint64_t t) {
	const uint32_t dc = dim_count(t);
	uint64_t s = d[0];
	if (dc >= 2) s *= d[1];
	if (dc == 3) s *= d[2];
	return (s * (uint64_t)(bits_per_channel(t) * channel_count(t))) / 8u;
}

/* level dim = dim >> level, no max(1) (image_types.hpp:751-766, host_image.cpp:75-81) */
void flo_level_dim(const uint32_t dim[4], uint64_t t, uint32_t level, uint32_t out[3]) {
	const uint32_t dc = dim_count(t);
	out[0] = level < 32 ? dim[0] >> level : 0;
	out[1] = dc >= 2 && level < 32 ? dim[1] >> level : 0;
	out[2] = dc >= 3 && level < 32 ? dim[2] >> level : 0;
}

/* bytes of one level over all layers (host_image.cpp:84-85) */
uint64_t flo_level_size(const uint32_t dim[4], uint64_t t, uint32_t level) {
	uint32_t d[3];
	flo_level_dim(dim, t, level, d);
	return slice_size(d, t) * flo_layer_count(dim, t);
}

/* image_types.hpp:731-776: offset of `level` == size of levels [0, level) */
uint64_t flo_level_offset(const uint32_t dim[4], uint64_t t, uint32_t level) {
	uint64_t off = 0;
	for (uint32_t l = 0; l < level; ++l) off += flo_level_size(dim, t, l);
	return off;
}

/* device_image.hpp:483-488: level count incl. mip_level_limit; size of all levels */
uint32_t flo_effective_level_count(const uint32_t dim[4], uint64_t t, uint32_t mip_level_limit) {
	uint32_t n = flo_mip_level_count(dim, t);
	if (mip_level_limit > 0 && mip_level_limit < n) n = mip_level_limit;
	return n;
}
uint64_t flo_image_data_size(const uint32_t dim[4], uint64_t t, uint32_t mip_level_limit) {
	return flo_level_offset(dim, t, flo_effective_level_count(dim, t, mip_level_limit));
}

/* ---- fp16 <-> fp32, IEEE RNE, subnormals kept (soft_f16.hpp:42-90 == __fp16 semantics) --- */
static float half_to_float(uint16_t h) {
	const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
	uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
	if (exp == 0) {
		if (man == 0) {
			bits = sign;
		} else { /* subnormal: normalise */
			int e = -1;
			do { ++e; man <<= 1; } while (!(man & 0x400u));
			bits = sign | (uint32_t)(127 - 15 - e) << 23 | (man & 0x3FFu) << 13;
		}
	} else if (exp == 31) {
		bits = sign | 0x7F800000u | man << 13;
	} else {
		bits = sign | (exp + 112u) << 23 | man << 13;
	}
	float f;
	memcpy(&f, &bits, 4);
	return f;
}
static uint16_t float_to_half(float f) {
	uint32_t x;
	memcpy(&x, &f, 4);
	const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
	x &= 0x7FFFFFFFu;
	if (x >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (x > 0x7F800000u ? 0x200u | ((x >> 13) & 0x3FFu) : 0u));
	if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u); /* rounds to >= 65520 -> inf */
	if (x < 0x33000001u) return sign;                        /* <= 2^-25 -> 0 (tie to even) */
	const int32_t e = (int32_t)(x >> 23) - 127;
	uint32_t man = (x & 0x7FFFFFu) | 0x800000u;
	uint32_t shift, hexp;
	if (e < -14) { shift = (uint32_t)(13 + (-14 - e)); hexp = 0; }
	else { shift = 13; hexp = (uint32_t)(e + 15); }
	const uint32_t rem = man & ((1u << shift) - 1u), half_ulp = 1u << (shift - 1);
	uint32_t q = man >> shift;
	if (rem > half_ulp || (rem == half_ulp && (q & 1u))) ++q;
	/* q carries the implicit bit for normals: (hexp << 10) + q - 0x400; a mantissa carry bumps the exponent */
	const uint32_t out = hexp == 0 ? q : ((hexp << 10) + q - 0x400u);
	return (uint16_t)(sign | out);
}
/* exported so that the tests can check them against the compiler's _Float16 */
float flo_half_to_float(uint16_t h) { return half_to_float(h); }
uint16_t flo_float_to_half(float f) { return float_to_half(f); }

/* ---- per-level table (host_image.cpp:81-108) --------------------------------------------- */
typedef struct {
	uint32_t dim[3];
	uint64_t offset;
	float fdim[3];      /* clamp_dim_float */
	float fdim_excl[3]; /* clamp_dim_float_excl = nextafterf(dim, 0) */
} level_info;

typedef struct {
	uint8_t* data;
	uint64_t type;
	uint32_t dc, channels, bpc, bpp, is_array, layers;
	int sample_class; /* 0 = float (float/normalized/depth-float), 1 = int, 2 = uint */
	int normalized, is_float_data, is_signed, no_double;
	level_info lv[ORACLE_MAX_LEVELS];
} image;

/* ---- sampler ------------------------------------------------------------------------------ */
/* const_math.hpp:859-869 / rt_math.hpp:160-172 */
static float wrapf(float val, float max) {
	uint32_t mb;
	memcpy(&mb, &max, 4);
	mb -= 1u;
	float next_towards_zero;
	memcpy(&next_towards_zero, &mb, 4);
	if (val < 0.0f) {
		const float w = max + fmodf(val, max);
		return w < next_towards_zero ? w : next_towards_zero;
	}
	return fmodf(val, max);
}
/* const_math.hpp:308-313 */
static float fractionalf(float v) { return v - floorf(v); }
/* const_math.hpp:852-856: clamp to [0, max] */
static float clamp0f(float v, float max) { return v > max ? max : (v < 0.0f ? 0.0f : v); }

/* host_image.hpp:141-174 (float coords, clamp-to-edge, non-cube) + :235-271 (offset) */
static uint64_t texel_offset(const image* img, uint32_t lod, const float coord[3], const int off[3], uint32_t layer) {
	const level_info* li = &img->lv[lod];
	uint32_t c[3] = { 0, 0, 0 };
	for (uint32_t d = 0; d < img->dc; ++d) {
		const float scaled = coord[d] * li->fdim[d];
		const float moved = scaled + (float)off[d];
		c[d] = (uint32_t)(int64_t)clamp0f(moved, li->fdim_excl[d]);
	}
	uint64_t texel;
	if (img->dc == 1) texel = c[0];
	else if (img->dc == 2) texel = (uint64_t)(li->dim[0] * c[1] + c[0]);
	else texel = (uint64_t)(li->dim[0] * li->dim[1] * c[2] + li->dim[0] * c[1] + c[0]);
	uint64_t o = li->offset + texel * img->bpp;
	if (img->is_array) o += slice_size(li->dim, img->type) * layer;
	return o;
}

/* host_image.hpp:333-383 (extract_channels) for FORMAT_2 / FORMAT_4: channel i sits at bits 6 - 2i of byte 0 / in the high (even i) or
   low nibble of byte i / 2.  Signed formats: the reference's sign fix-up reads
       if (bpc % 8 != 0 && *((uchannel_type*)&ret[i]) & high_bits[i] != 0u) { ret[i] ^= high_bits[i]; ret[i] = channel_type(-ret[i]); }
   where `!=` binds tighter than `&`: the test is on bit 0 of the channel, not on its sign bit.  Restated as written (the
   reference's own build, oracle/_ref, behaves this way: tests/test_reference_pin.py). */
static int32_t extract_packed(const image* img, const uint8_t* p, uint32_t i) {
	uint8_t v = img->bpc == 2 ? (uint8_t)((p[0] >> (6u - 2u * i)) & 0x3u) : (uint8_t)((p[i / 2u] >> (i % 2u == 0 ? 4u : 0u)) & 0xFu);
	if (!img->is_signed) return (int32_t)v;
	int8_t sv = (int8_t)v;
	if (v & 1u) {
		sv = (int8_t)(sv ^ (int8_t)(1u << (img->bpc - 1u)));
		sv = (int8_t)(-sv);
	}
	return (int32_t)sv;
}

/* host_image.hpp:487-561: decode to float4 (only the stored channels are meaningful) */
static void decode_float(const image* img, const uint8_t* p, float out[4]) {
	for (uint32_t i = 0; i < img->channels; ++i) {
		if (img->bpc < 8) {
			const uint64_t mask = (1ull << img->bpc) - 1ull;
			const double den = (double)(img->is_signed ? (mask >> 1) : mask);
			const float factor = (float)(1.0 / (den > 1.0 ? den : 1.0));
			out[i] = (float)extract_packed(img, p, i) * factor;
			continue;
		}
		if (img->is_float_data) {
			if (img->bpc == 32) memcpy(&out[i], p + 4 * i, 4);
			else { uint16_t h; memcpy(&h, p + 2 * i, 2); out[i] = half_to_float(h); }
		} else if (!img->is_signed) {
			/* float(1.0 / (2^bpc - 1)) */
			const float factor = (float)(1.0 / (double)((1ull << img->bpc) - 1ull));
			uint32_t v = 0;
			if (img->bpc == 8) v = p[i];
			else if (img->bpc == 16) { uint16_t h; memcpy(&h, p + 2 * i, 2); v = h; }
			else memcpy(&v, p + 4 * i, 4);
			out[i] = (float)v * factor;
		} else {
			/* float(1.0 / (2^(bpc-1) - 1)) */
			const float factor = (float)(1.0 / (double)(((1ull << img->bpc) - 1ull) >> 1));
			int32_t v = 0;
			if (img->bpc == 8) v = (int8_t)p[i];
			else if (img->bpc == 16) { int16_t h; memcpy(&h, p + 2 * i, 2); v = h; }
			else memcpy(&v, p + 4 * i, 4);
			out[i] = (float)v * factor;
		}
	}
}

/* host_image.hpp:640-667: widen to 32 bit, no float involved */
static void decode_int(const image* img, const uint8_t* p, uint32_t out[4]) {
	for (uint32_t i = 0; i < img->channels; ++i) {
		if (img->bpc == 8) out[i] = img->is_signed ? (uint32_t)(int32_t)(int8_t)p[i] : p[i];
		else if (img->bpc == 16) {
			uint16_t h; memcpy(&h, p + 2 * i, 2);
			out[i] = img->is_signed ? (uint32_t)(int32_t)(int16_t)h : h;
		} else memcpy(&out[i], p + 4 * i, 4);
	}
}

/* host_image.hpp:672-722 + insert_channels :391-460 */
static void encode_float(const image* img, uint8_t* p, const float c[4]) {
	if (img->bpc < 8) {
		/* FORMAT_2 / FORMAT_4 (:421-446): channel_type is 8 bits wide, scale in float, truncate; unsigned keeps the low bits, signed
		   keeps the low bpc - 1 bits and puts (value < 0) into the channel's top bit; the texel's bytes are cleared first */
		memset(p, 0, img->bpp);
		const uint32_t scale_i = (1u << (img->bpc - (img->is_signed ? 1u : 0u))) - 1u;
		for (uint32_t i = 0; i < img->channels; ++i) {
			const float scaled = c[i] * (float)scale_i;
			uint32_t bits;
			if (!img->is_signed) {
				const uint8_t q = (uint8_t)scaled; /* channel_type(fp) with channel_type = uint8_t */
				bits = q & ((1u << img->bpc) - 1u);
			} else {
				const int8_t q = (int8_t)scaled;
				bits = ((uint32_t)q & ((1u << (img->bpc - 1u)) - 1u)) | (q < 0 ? 1u << (img->bpc - 1u) : 0u);
			}
			if (img->bpc == 2) p[0] |= (uint8_t)(bits << (6u - 2u * i));
			else p[i / 2u] |= (uint8_t)(bits << (i % 2u == 0 ? 4u : 0u));
		}
		return;
	}
	for (uint32_t i = 0; i < img->channels; ++i) {
		if (img->is_float_data) {
			if (img->bpc == 32) memcpy(p + 4 * i, &c[i], 4);
			else { const uint16_t h = float_to_half(c[i]); memcpy(p + 2 * i, &h, 2); }
			continue;
		}
		/* scale with 2^bpc - 1 (unsigned) or 2^(bpc-1) - 1 (signed); C cast == truncation */
		const uint64_t scale_i = ((1ull << (img->bpc - (img->is_signed ? 1u : 0u))) - 1ull);
		int64_t q;
		if (img->bpc <= 8 || img->no_double) q = (int64_t)(c[i] * (float)scale_i);
		else if (img->bpc <= 16) q = (int64_t)((double)c[i] * (double)scale_i);
		else q = (int64_t)((long double)c[i] * (long double)scale_i);
		if (img->bpc == 8) p[i] = (uint8_t)q;
		else if (img->bpc == 16) { const uint16_t h = (uint16_t)q; memcpy(p + 2 * i, &h, 2); }
		else { const uint32_t w = (uint32_t)q; memcpy(p + 4 * i, &w, 4); }
	}
}

/* host_image.hpp:801-825: narrowing cast per channel */
static void encode_int(const image* img, uint8_t* p, const uint32_t c[4]) {
	for (uint32_t i = 0; i < img->channels; ++i) {
		if (img->bpc == 8) p[i] = (uint8_t)c[i];
		else if (img->bpc == 16) { const uint16_t h = (uint16_t)c[i]; memcpy(p + 2 * i, &h, 2); }
		else memcpy(p + 4 * i, &c[i], 4);
	}
}

/* const_math.hpp:981-984 (Host-Compute branch) */
static float lerp_f(float a, float b, float t) {
	const float d = b - a;
	const float s = d * t;
	return s + a;
}
/* const_math.hpp:990-996: any_type(fp_type(b - a) * t) + a */
static uint32_t lerp_i(uint32_t a, uint32_t b, float t, int is_signed) {
	const uint32_t d = b - a; /* (b - a) in T */
	const float fd = is_signed ? (float)(int32_t)d : (float)d;
	const float s = fd * t;
	const uint32_t back = is_signed ? (uint32_t)(int32_t)(int64_t)s : (uint32_t)(int64_t)s;
	return back + a;
}

/* mip_map_minify.hpp:89-108 + host_image.hpp:842-929: one destination texel */
static void minify_texel(const image* img, const uint32_t g[3], const float inv_prev[3], uint32_t level, uint32_t layer) {
	const uint32_t dc = img->dc, lod = level - 1u;
	const level_info* li = &img->lv[lod];
	float coord[3] = { 0, 0, 0 }, w[3] = { 0, 0, 0 };
	int so[3] = { 0, 0, 0 };
	for (uint32_t d = 0; d < dc; ++d) {
		coord[d] = (float)(g[d] * 2u + 1u) * inv_prev[d];
		const float scaled = wrapf(coord[d], 1.0f) * li->fdim[d];
		const float frac = fractionalf(scaled);
		so[d] = frac < 0.5f ? -1 : 1;
		w[d] = frac < 0.5f ? frac + 0.5f : 1.5f - frac;
	}
	/* colors[k]: bit0 clear -> x outside texel, bit1 clear -> y outside, bit2 clear -> z outside */
	const uint32_t n = 1u << dc;
	float cf[8][4];
	uint32_t ci[8][4];
	for (uint32_t k = 0; k < n; ++k) {
		const int off[3] = { (k & 1u) ? 0 : so[0], (k & 2u) ? 0 : so[1], (k & 4u) ? 0 : so[2] };
		const uint8_t* p = img->data + texel_offset(img, lod, coord, off, layer);
		if (img->sample_class == 0) decode_float(img, p, cf[k]);
		else decode_int(img, p, ci[k]);
	}
	/* x first, then y, then z; colors[even] is the `a` (t == 0) operand */
	for (uint32_t d = 0; d < dc; ++d) {
		const uint32_t step = 1u << d;
		for (uint32_t k = 0; k < n; k += 2u * step) {
			for (uint32_t c = 0; c < img->channels; ++c) {
				if (img->sample_class == 0) cf[k][c] = lerp_f(cf[k][c], cf[k + step][c], w[d]);
				else ci[k][c] = lerp_i(ci[k][c], ci[k + step][c], w[d], img->sample_class == 1);
			}
		}
	}
	/* write_lod: integer coords are clamped to [0, dim-1] (host_image.hpp:141-170) -- always in range here */
	const level_info* lo = &img->lv[level];
	uint64_t texel;
	if (dc == 1) texel = g[0];
	else if (dc == 2) texel = (uint64_t)(lo->dim[0] * g[1] + g[0]);
	else texel = (uint64_t)(lo->dim[0] * lo->dim[1] * g[2] + lo->dim[0] * g[1] + g[0]);
	uint8_t* q = img->data + lo->offset + texel * img->bpp + (img->is_array ? slice_size(lo->dim, img->type) * layer : 0);
	if (img->sample_class == 0) encode_float(img, q, cf[0]);
	else encode_int(img, q, ci[0]);
}

static int image_init(image* img, uint8_t* data, const uint32_t dim[4], uint64_t type, uint32_t flags) {
	memset(img, 0, sizeof(*img));
	img->data = data;
	img->type = type;
	img->dc = dim_count(type);
	img->channels = channel_count(type);
	img->bpc = bits_per_channel(type);
	img->bpp = flo_bytes_per_pixel(type);
	/* cube faces are addressed as layers array_idx*6+face (host_image.hpp:263-271) */
	img->is_array = (type & (T_FLAG_ARRAY | T_FLAG_CUBE)) != 0;
	img->layers = flo_layer_count(dim, type);
	img->no_double = (flags & FLO_FLAG_NO_DOUBLE) != 0;
	if (img->dc < 1 || img->dc > 3 || img->bpc == 0) return FLO_ERR_UNSUPPORTED;
	if ((type & T_COMPRESSION_MASK) || (type & (T_FLAG_MSAA | T_FLAG_STENCIL))) return FLO_ERR_UNSUPPORTED;
	const uint64_t dt = type & T_DATA_TYPE_MASK;
	img->normalized = (type & T_FLAG_NORMALIZED) != 0;
	img->is_float_data = dt == T_FLOAT;
	img->is_signed = dt == T_INT;
	if (dt == 0) return FLO_ERR_UNSUPPORTED;
	if (img->is_float_data && img->bpc == 8) return FLO_ERR_UNSUPPORTED;
	/* FORMAT_2 / FORMAT_4 only exist normalized (host_image.hpp:1167-1186), and only where a texel is a whole number of bytes: the
	   reference sizes images by bits (image_types.hpp:675-691) but addresses texels by bytes_per_pixel = ceil(bits / 8)
	   (host_image.hpp:235-271), so R2 / RG2 / RGB2 / R4 / RGB4 images are smaller than what its own kernels touch */
	if (img->bpc < 8 && (!img->normalized || (img->bpc * img->channels) % 8u != 0u)) return FLO_ERR_UNSUPPORTED;
	if (type & T_FLAG_DEPTH) { /* only the DEPTH_FLOAT kernels exist (mip_map_minify.hpp:22-30) */
		if (!(img->is_float_data && img->bpc == 32 && img->channels == 1)) return FLO_ERR_UNSUPPORTED;
	}
	/* kernel selection (mip_map_minify.hpp:53-69): normalized -> FLOAT sample type */
	img->sample_class = (img->normalized || img->is_float_data) ? 0 : (img->is_signed ? 1 : 2);
	uint64_t off = 0;
	for (uint32_t l = 0; l < ORACLE_MAX_LEVELS; ++l) {
		level_info* li = &img->lv[l];
		flo_level_dim(dim, type, l, li->dim);
		li->offset = off;
		off += slice_size(li->dim, type) * img->layers;
		for (int d = 0; d < 3; ++d) {
			li->fdim[d] = li->dim[d] > 0 ? (float)li->dim[d] : 0.0f;
			li->fdim_excl[d] = li->dim[d] > 0 ? nextafterf((float)li->dim[d], 0.0f) : 0.0f;
		}
	}
	return FLO_OK;
}

/* ---- execution: one "launch" per (layer, level), work-groups handed out by atomic ticket --- */
typedef struct {
	const image* img;
	uint32_t level, layer, level_size[3], lsize[3], groups[3];
	float inv_prev[3];
	atomic_uint ticket;
} launch;

static void run_groups(launch* L) {
	const uint32_t total = L->groups[0] * L->groups[1] * L->groups[2];
	const uint32_t dc = L->img->dc;
	for (;;) {
		const uint32_t t = atomic_fetch_add(&L->ticket, 1u);
		if (t >= total) break;
		const uint32_t gx = t % L->groups[0], gy = (t / L->groups[0]) % L->groups[1], gz = t / (L->groups[0] * L->groups[1]);
		for (uint32_t z = 0; z < L->lsize[2]; ++z)
			for (uint32_t y = 0; y < L->lsize[1]; ++y)
				for (uint32_t x = 0; x < L->lsize[0]; ++x) {
					const uint32_t g[3] = { gx * L->lsize[0] + x, gy * L->lsize[1] + y, gz * L->lsize[2] + z };
					/* mip_map_minify.hpp:99-100 */
					if (g[0] >= L->level_size[0]) continue;
					if (dc >= 2 && g[1] >= L->level_size[1]) continue;
					if (dc >= 3 && g[2] >= L->level_size[2]) continue;
					minify_texel(L->img, g, L->inv_prev, L->level, L->layer);
				}
	}
}

typedef struct {
	pthread_barrier_t bar;
	launch* cur;
	int quit;
} pool;
typedef struct { pool* p; } worker_arg;

static void* worker(void* a) {
	pool* p = ((worker_arg*)a)->p;
	for (;;) {
		pthread_barrier_wait(&p->bar);
		if (p->quit) break;
		run_groups(p->cur);
		pthread_barrier_wait(&p->bar);
	}
	return NULL;
}

/* device_image.cpp:290-327 */
int flo_generate_mip_map_chain(uint8_t* data, const uint32_t dim[4], uint64_t type, uint32_t mip_level_limit,
							   uint32_t flags, uint32_t threads) {
	image img;
	const int rc = image_init(&img, data, dim, type, flags);
	if (rc != FLO_OK) return rc;
	if (type & T_FLAG_CUBE) {
		/* the reference kernel static_asserts cube out (mip_map_minify.hpp:97); defined here as the
		   2D-array path over layer = cube*6 + face on the same bytes */
	}
	const uint32_t levels = flo_effective_level_count(dim, type, mip_level_limit);
	if (levels > ORACLE_MAX_LEVELS) return FLO_ERR_UNSUPPORTED;
	if (threads == 0) threads = 1;

	pool P;
	pthread_t* tids = NULL;
	worker_arg wa = { &P };
	if (threads > 1) {
		P.quit = 0;
		P.cur = NULL;
		pthread_barrier_init(&P.bar, NULL, threads);
		tids = (pthread_t*)malloc(sizeof(pthread_t) * (threads - 1));
		for (uint32_t i = 0; i + 1 < threads; ++i) pthread_create(&tids[i], NULL, worker, &wa);
	}

	launch L;
	L.img = &img;
	/* max_total_local_size of the host device is 1024 -> 1024 / 32x32 / 32x16x2 */
	if (img.dc == 1) { L.lsize[0] = 1024; L.lsize[1] = 1; L.lsize[2] = 1; }
	else if (img.dc == 2) { L.lsize[0] = 32; L.lsize[1] = 32; L.lsize[2] = 1; }
	else { L.lsize[0] = 32; L.lsize[1] = 16; L.lsize[2] = 2; }

	for (uint32_t layer = 0; layer < img.layers; ++layer) {
		uint32_t level_size[3] = { dim[0], img.dc >= 2 ? dim[1] : 0u, img.dc >= 3 ? dim[2] : 0u };
		float inv_prev[3] = { 0, 0, 0 };
		for (uint32_t level = 0; level < levels; ++level) {
			if (level > 0) {
				L.level = level;
				L.layer = layer;
				uint32_t empty = 0;
				for (int d = 0; d < 3; ++d) {
					L.level_size[d] = level_size[d];
					L.inv_prev[d] = inv_prev[d];
					/* global size = level_size rounded up to lsize; a zero dim means zero work-items */
					L.groups[d] = (uint32_t)d < img.dc ? (level_size[d] + L.lsize[d] - 1u) / L.lsize[d] : 1u;
					if ((uint32_t)d < img.dc && level_size[d] == 0) empty = 1;
				}
				if (!empty) {
					atomic_store(&L.ticket, 0u);
					if (threads > 1) {
						P.cur = &L;
						pthread_barrier_wait(&P.bar);
						run_groups(&L);
						pthread_barrier_wait(&P.bar);
					} else {
						run_groups(&L);
					}
				}
			}
			/* loop increment: inv_prev = 1.0f / float3(level_size); level_size >>= 1 */
			for (int d = 0; d < 3; ++d) {
				inv_prev[d] = 1.0f / (float)level_size[d];
				level_size[d] >>= 1;
			}
		}
	}

	if (threads > 1) {
		P.quit = 1;
		pthread_barrier_wait(&P.bar);
		for (uint32_t i = 0; i + 1 < threads; ++i) pthread_join(tids[i], NULL);
		pthread_barrier_destroy(&P.bar);
		free(tids);
	}
	return FLO_OK;
}

/* ---- counter-based synthetic input; shared definition with the CUDA fill kernel ---- */
static uint64_t splitmix64(uint64_t x) {
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}
#define FLO_SEED 0x9E3779B97F4A7C15ull

/* storage bits of channel element `elem` (= texel_index*channels + c) of `layer` */
uint32_t flo_synth_element(uint64_t type, uint64_t config_id, uint64_t layer, uint64_t elem) {
	const uint64_t r = splitmix64(splitmix64(splitmix64(FLO_SEED + config_id) + layer) + elem);
	const uint32_t bpc = bits_per_channel(type);
	const uint64_t dt = type & T_DATA_TYPE_MASK;
	if (dt == T_FLOAT) {
		if (bpc == 16) { /* finite normals in +-[2^-14, 2^5): exponent field 1..19 */
			const uint32_t sign = (uint32_t)(r >> 63), ex = 1u + (uint32_t)((r >> 32) % 19u), man = (uint32_t)(r & 0x3FFu);
			return sign << 15 | ex << 10 | man;
		}
		const float f = (float)(r >> 40) * 0x1p-24f; /* uniform [0, 1) */
		uint32_t b;
		memcpy(&b, &f, 4);
		return b;
	}
	if (bpc == 32 && dt == T_INT && !(type & T_FLAG_NORMALIZED)) return (uint32_t)((int32_t)(uint32_t)r >> 1); /* keep b - a in range */
	return bpc == 32 ? (uint32_t)r : (uint32_t)(r & ((1ull << bpc) - 1ull));
}

/* fills level 0 of layers [layer_begin, layer_begin+layer_num) placed at `dst` (layer-major), with global
   layer ids starting at `layer_id0` so that a per-GPU shard reproduces the bytes of the whole image */
void flo_fill_synthetic(uint8_t* dst, const uint32_t dim[4], uint64_t type, uint64_t config_id, uint64_t layer_id0,
						uint32_t layer_num) {
	uint32_t d0[3];
	flo_level_dim(dim, type, 0, d0);
	uint32_t bpc = bits_per_channel(type);
	const uint64_t slice_bytes = slice_size(d0, type);
	if (bpc < 8) { /* FORMAT_2 / FORMAT_4: the pattern is per storage byte (that of an 8-bit single-channel unorm image of the same bytes) */
		type = (type & ~(uint64_t)0x3Full & ~(uint64_t)0xC000ull & ~T_DATA_TYPE_MASK) | FMT_8 | T_UINT;
		bpc = 8;
	}
	const uint64_t elems = slice_bytes / (bpc / 8u);
	for (uint32_t l = 0; l < layer_num; ++l) {
		uint8_t* p = dst + (uint64_t)l * elems * (bpc / 8u);
		for (uint64_t e = 0; e < elems; ++e) {
			const uint32_t v = flo_synth_element(type, config_id, layer_id0 + l, e);
			if (bpc == 8) p[e] = (uint8_t)v;
			else if (bpc == 16) { const uint16_t h = (uint16_t)v; memcpy(p + 2 * e, &h, 2); }
			else memcpy(p + 4 * e, &v, 4);
		}
	}
}
