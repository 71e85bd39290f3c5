// BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py): runs the reference's OWN GPU implementation of the mip-chain path on this
// box, so that bench.py can put a GPU-over-GPU number beside the GPU-over-CPU one ("the bar on the same box is this
// sm_50 PTX JIT'd on B200 via floor's per-level launch loop").  Nothing here is product code and nothing of the reference is
// copied: the kernels are the PTX text that oracle/build_incumbent.py extracts from the reference's etc/mip_map_minify/mmm.fubar
// into oracle/_ref/ at build time; this file restates, through the plain driver API, the HOST side the reference wraps around them:
//   * JIT: cuModuleLoadDataEx with TARGET = device sm, MAX_REGISTERS = toolchain.cuda.max_registers (32), OPTIMIZATION_LEVEL = 4
//     (src/device/cuda/cuda_context.cpp:606-633, src/floor/floor.cpp:431-433);
//   * storage: one CUmipmappedArray per image (LAYERED for arrays, SURFACE_LDST), format LUT of cuda_image.cpp:197-207, the texture
//     object of sampler index 3 = normalized coordinates | linear filter | clamp-to-edge | no compare (backend/cuda_sampler.hpp:25-95;
//     the minify kernels read exactly this one of the 96, `ld.param.u32 [param_0 + 12]`), one surface object per level and the device
//     buffer with the per-level surface handles (cuda_image.cpp:404-526);
//   * argument block: 96 x u32 texture ids, u64 surface of level 0, u64 pointer to the surface-LOD buffer, u64 run-time image type
//     (408 bytes), then uint3 level_size, float3 inv_prev_level_size, u32 level, u32 layer (src/device/cuda/cuda_function.cpp:119-166);
//   * the launch loop: layer outermost, one launch per level, block 32x32 / 32x16x2 / 1024, cuStreamSynchronize after EVERY
//     launch (src/device/device_image.cpp:290-327, cuda_function.cpp:225-228).
// Two timings per chain: `blocking` = the loop as the reference runs it (host wall clock, what an application pays), and `enqueued`
// = the same launches without the per-launch sync, between CUDA events (what the kernels alone cost: generous to the reference).
// The results are read back in floor's host layout so the caller can look at them; they are NOT the parity target (the texture
// unit filters with 9-bit fixed-point weights).
#include <cuda.h>
#include <dlfcn.h>

#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

namespace {
#define FN_LIST(F)                                                                                                                          \
	F(cuInit) F(cuDeviceGet) F(cuDevicePrimaryCtxRetain) F(cuCtxPushCurrent) F(cuCtxPopCurrent) F(cuGetErrorName) F(cuDeviceGetAttribute)   \
	F(cuModuleLoadDataEx) F(cuModuleUnload) F(cuModuleGetFunction) F(cuMipmappedArrayCreate) F(cuMipmappedArrayGetLevel)                   \
	F(cuMipmappedArrayDestroy) F(cuTexObjectCreate) F(cuTexObjectDestroy) F(cuSurfObjectCreate) F(cuSurfObjectDestroy) F(cuMemAlloc)       \
	F(cuMemFree) F(cuMemcpyHtoD) F(cuMemcpy3D) F(cuStreamCreate) F(cuStreamDestroy) F(cuStreamSynchronize) F(cuLaunchKernelEx)              \
	F(cuEventCreate) F(cuEventRecord) F(cuEventSynchronize) F(cuEventElapsedTime) F(cuEventDestroy)
struct api {
#define D(n) decltype(&n) n##_ = nullptr;
	FN_LIST(D)
#undef D
} cu;
std::string g_err;
bool load_api() {
	static bool done = false, ok = false;
	if (done) return ok;
	done = true;
	void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h) { g_err = "libcuda not found"; return false; }
	// cuda.h maps most names to versioned symbols (cuMemcpy3D -> cuMemcpy3D_v2): stringify AFTER macro expansion
#define STR2(x) #x
#define STR(x) STR2(x)
#define L(n)                                                       \
	cu.n##_ = reinterpret_cast<decltype(&n)>(dlsym(h, STR(n)));    \
	if (!cu.n##_) { g_err = std::string("libcuda lacks ") + STR(n); return false; }
	FN_LIST(L)
#undef L
	ok = cu.cuInit_(0) == CUDA_SUCCESS;
	if (!ok) g_err = "cuInit failed";
	return ok;
}
bool fail(const char* fmt, ...) {
	char b[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(b, sizeof(b), fmt, ap);
	va_end(ap);
	g_err = b;
	return false;
}
#define CK(call)                                                          \
	do {                                                                  \
		const CUresult r_ = (call);                                       \
		if (r_ != CUDA_SUCCESS) {                                         \
			const char* n_ = nullptr;                                     \
			cu.cuGetErrorName_(r_, &n_);                                  \
			return fail("%s failed: %s", #call, n_ ? n_ : "?");           \
		}                                                                 \
	} while (0)

// IMAGE_TYPE bits (include/floor/device/backend/image_types.hpp:24-236)
constexpr uint64_t T_FORMAT_MASK = 0x3F, T_DATA_MASK = 0x3000, T_INT = 0x1000, T_UINT = 0x2000, T_FLOAT = 0x3000;
constexpr uint64_t T_ARRAY = 1ull << 20, T_CUBE = 1ull << 23, T_DEPTH = 1ull << 24, T_NORMALIZED = 1ull << 30;

struct state {
	CUcontext ctx = nullptr;
	CUmodule mod = nullptr;
	CUmipmappedArray arr = nullptr;
	std::vector<CUarray> levels;
	CUtexObject tex = 0;
	std::vector<CUsurfObject> surfs;
	CUdeviceptr surf_lod = 0;
	CUstream stream = nullptr;
	~state() {
		if (!ctx) return;
		if (stream) cu.cuStreamDestroy_(stream);
		if (surf_lod) cu.cuMemFree_(surf_lod);
		for (auto s : surfs) cu.cuSurfObjectDestroy_(s);
		if (tex) cu.cuTexObjectDestroy_(tex);
		if (arr) cu.cuMipmappedArrayDestroy_(arr);
		if (mod) cu.cuModuleUnload_(mod);
		CUcontext old = nullptr;
		cu.cuCtxPopCurrent_(&old);
	}
};

bool run(const char* ptx_path, int device, uint64_t type, const uint32_t dim[4], const void* level0, void* out_all, uint32_t warmup, uint32_t steps,
		 double* blocking_ms, double* enqueued_ms, uint64_t* launches_per_chain) {
	if (!load_api()) return false;
	std::ifstream f(ptx_path, std::ios::binary);
	std::string ptx((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	if (ptx.empty()) return fail("cannot read %s", ptx_path);
	ptx.push_back('\0');

	const uint32_t dc = (uint32_t)((type >> 16) & 3), channels = (uint32_t)((type >> 14) & 3) + 1, fmt = (uint32_t)(type & T_FORMAT_MASK);
	const uint32_t bpc = fmt == 11 ? 8 : fmt == 18 ? 16 : fmt == 22 ? 32 : 0;
	const uint64_t dt = type & T_DATA_MASK;
	if (dc < 1 || dc > 3 || bpc == 0 || channels == 3) return fail("image type %#llx is not supported by the reference's CUDA backend", (unsigned long long)type);
	if (type & T_CUBE) return fail("the reference has no minify kernel for cube images (mip_map_minify.hpp:95-97, device_image.cpp:278-283)");
	const bool is_array = (type & T_ARRAY) != 0;
	const uint32_t layers = !is_array ? 1u : (dc == 1 ? dim[1] : dim[2]);
	const uint32_t bpp = bpc / 8 * channels;
	uint32_t m = dim[0];
	if (dc >= 2 && dim[1] > m) m = dim[1];
	if (dc >= 3 && dim[2] > m) m = dim[2];
	uint32_t level_count = 0;
	while (m) { ++level_count; m >>= 1; }

	state S;
	CUdevice dev;
	CK(cu.cuDeviceGet_(&dev, device));
	CK(cu.cuDevicePrimaryCtxRetain_(&S.ctx, dev));
	CK(cu.cuCtxPushCurrent_(S.ctx));
	int sm_major = 0, sm_minor = 0, max_threads = 0;
	cu.cuDeviceGetAttribute_(&sm_major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, dev);
	cu.cuDeviceGetAttribute_(&sm_minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, dev);
	cu.cuDeviceGetAttribute_(&max_threads, CU_DEVICE_ATTRIBUTE_MAX_THREADS_PER_BLOCK, dev);

	// ---- JIT (cuda_context.cpp:606-633) ----
	CUjit_option opts[] = { CU_JIT_TARGET, CU_JIT_GENERATE_LINE_INFO, CU_JIT_GENERATE_DEBUG_INFO, CU_JIT_MAX_REGISTERS, CU_JIT_OPTIMIZATION_LEVEL };
	void* vals[] = { (void*)(size_t)(sm_major * 10 + sm_minor), (void*)(size_t)0, (void*)(size_t)0, (void*)(size_t)32, (void*)(size_t)4 };
	CK(cu.cuModuleLoadDataEx_(&S.mod, ptx.data(), 5, opts, vals));
	std::string name = "libfloor_mip_map_minify_";
	if (type & T_DEPTH) name += is_array ? "IMAGE_DEPTH_ARRAY" : "IMAGE_DEPTH";
	else name += dc == 1 ? (is_array ? "IMAGE_1D_ARRAY" : "IMAGE_1D") : dc == 2 ? (is_array ? "IMAGE_2D_ARRAY" : "IMAGE_2D") : "IMAGE_3D";
	// normalized formats are sampled as FLOAT (minify_image_base_type, mip_map_minify.hpp:51-69)
	name += (dt == T_FLOAT || (type & T_NORMALIZED)) ? "_FLOAT" : (dt == T_INT ? "_INT" : "_UINT");
	CUfunction fn = nullptr;
	CK(cu.cuModuleGetFunction_(&fn, S.mod, name.c_str()));

	// ---- storage (cuda_image.cpp:158-330) ----
	CUarray_format af;
	if (dt == T_FLOAT) af = bpc == 16 ? CU_AD_FORMAT_HALF : CU_AD_FORMAT_FLOAT;
	else if (dt == T_INT) af = bpc == 8 ? CU_AD_FORMAT_SIGNED_INT8 : bpc == 16 ? CU_AD_FORMAT_SIGNED_INT16 : CU_AD_FORMAT_SIGNED_INT32;
	else af = bpc == 8 ? CU_AD_FORMAT_UNSIGNED_INT8 : bpc == 16 ? CU_AD_FORMAT_UNSIGNED_INT16 : CU_AD_FORMAT_UNSIGNED_INT32;
	CUDA_ARRAY3D_DESCRIPTOR ad;
	memset(&ad, 0, sizeof(ad));
	ad.Width = dim[0];
	ad.Height = dc >= 2 ? dim[1] : 0;
	ad.Depth = dc >= 3 ? dim[2] : (is_array ? layers : 0);
	ad.Format = af;
	ad.NumChannels = channels;
	ad.Flags = (is_array ? CUDA_ARRAY3D_LAYERED : 0u) | CUDA_ARRAY3D_SURFACE_LDST;
	CK(cu.cuMipmappedArrayCreate_(&S.arr, &ad, level_count));
	S.levels.resize(level_count);
	for (uint32_t l = 0; l < level_count; ++l) CK(cu.cuMipmappedArrayGetLevel_(&S.levels[l], S.arr, l));

	auto level_dims = [&](uint32_t l, uint32_t d[3]) {
		d[0] = dim[0] >> l;
		d[1] = dc >= 2 ? dim[1] >> l : 1u;
		d[2] = dc >= 3 ? dim[2] >> l : 1u;
	};
	auto copy_level = [&](uint32_t l, bool to_device, uint8_t* host) -> bool {
		uint32_t d[3];
		level_dims(l, d);
		if (d[0] == 0 || d[1] == 0 || d[2] == 0) return true;
		CUDA_MEMCPY3D c;
		memset(&c, 0, sizeof(c));
		c.WidthInBytes = (size_t)d[0] * bpp;
		c.Height = d[1];
		c.Depth = dc == 3 ? d[2] : layers;
		if (to_device) {
			c.srcMemoryType = CU_MEMORYTYPE_HOST; c.srcHost = host; c.srcPitch = c.WidthInBytes; c.srcHeight = d[1];
			c.dstMemoryType = CU_MEMORYTYPE_ARRAY; c.dstArray = S.levels[l];
		} else {
			c.srcMemoryType = CU_MEMORYTYPE_ARRAY; c.srcArray = S.levels[l];
			c.dstMemoryType = CU_MEMORYTYPE_HOST; c.dstHost = host; c.dstPitch = c.WidthInBytes; c.dstHeight = d[1];
		}
		CK(cu.cuMemcpy3D_(&c));
		return true;
	};
	if (level0 && !copy_level(0, true, (uint8_t*)const_cast<void*>(level0))) return false;

	// ---- texture object #3 and the per-level surfaces (cuda_image.cpp:404-526) ----
	CUDA_RESOURCE_DESC rd;
	memset(&rd, 0, sizeof(rd));
	rd.resType = CU_RESOURCE_TYPE_MIPMAPPED_ARRAY;
	rd.res.mipmap.hMipmappedArray = S.arr;
	CUDA_TEXTURE_DESC td;
	memset(&td, 0, sizeof(td));
	td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = CU_TR_ADDRESS_MODE_CLAMP;
	td.filterMode = CU_TR_FILTER_MODE_LINEAR;
	td.mipmapFilterMode = CU_TR_FILTER_MODE_LINEAR;
	td.flags = CU_TRSF_NORMALIZED_COORDINATES; // no READ_AS_INTEGER: 8 / 16-bit integer storage is sampled as normalized float
	td.maxAnisotropy = 1;
	td.minMipmapLevelClamp = 0.0f;
	td.maxMipmapLevelClamp = 16.0f;
	CK(cu.cuTexObjectCreate_(&S.tex, &rd, &td, nullptr));
	S.surfs.resize(level_count);
	for (uint32_t l = 0; l < level_count; ++l) {
		CUDA_RESOURCE_DESC sd;
		memset(&sd, 0, sizeof(sd));
		sd.resType = CU_RESOURCE_TYPE_ARRAY;
		sd.res.array.hArray = S.levels[l];
		CK(cu.cuSurfObjectCreate_(&S.surfs[l], &sd));
	}
	CK(cu.cuMemAlloc_(&S.surf_lod, level_count * sizeof(CUsurfObject)));
	CK(cu.cuMemcpyHtoD_(S.surf_lod, S.surfs.data(), level_count * sizeof(CUsurfObject)));
	CK(cu.cuStreamCreate_(&S.stream, CU_STREAM_NON_BLOCKING));

	// ---- argument block (cuda_function.cpp:119-166) ----
	struct image_arg {
		uint32_t textures[96];
		uint64_t surface0, surf_lod, image_type;
	} ia;
	static_assert(sizeof(image_arg) == 408, "the kernels declare param_0[408]");
	memset(&ia, 0, sizeof(ia));
	if (S.tex >> 32) return fail("texture object id does not fit 32 bits");
	for (int i = 0; i < 96; ++i) ia.textures[i] = (uint32_t)S.tex; // the kernels read slot 3 only
	ia.surface0 = S.surfs[0];
	ia.surf_lod = S.surf_lod;
	ia.image_type = type;

	// ---- the launch loop (device_image.cpp:290-327) ----
	uint32_t ls[3];
	if (dc == 1) { ls[0] = (uint32_t)max_threads; ls[1] = 1; ls[2] = 1; }
	else if (dc == 2) { ls[0] = max_threads > 256 ? 32 : 16; ls[1] = max_threads > 512 ? 32 : 16; ls[2] = 1; }
	else { ls[0] = max_threads > 512 ? 32 : 16; ls[1] = max_threads > 256 ? 16 : 8; ls[2] = 2; }
	uint64_t launches = 0;
	auto chain = [&](bool blocking) -> bool {
		launches = 0;
		for (uint32_t layer = 0; layer < layers; ++layer) {
			uint32_t size[3] = { dim[0], dc >= 2 ? dim[1] : 0u, dc >= 3 ? dim[2] : 0u };
			float inv_prev[3] = { 0, 0, 0 };
			for (uint32_t level = 0; level < level_count; ++level) {
				if (level != 0) {
					uint32_t grid[3];
					for (int d = 0; d < 3; ++d) {
						grid[d] = (size[d] + ls[d] - 1) / ls[d];
						if (grid[d] == 0) grid[d] = 1;
					}
					uint32_t lvl = level, lay = layer;
					void* args[] = { &ia, size, inv_prev, &lvl, &lay };
					CUlaunchConfig cfg;
					memset(&cfg, 0, sizeof(cfg));
					cfg.gridDimX = grid[0]; cfg.gridDimY = grid[1]; cfg.gridDimZ = grid[2];
					cfg.blockDimX = ls[0]; cfg.blockDimY = ls[1]; cfg.blockDimZ = ls[2];
					cfg.hStream = S.stream;
					CK(cu.cuLaunchKernelEx_(&cfg, fn, args, nullptr));
					++launches;
					if (blocking) CK(cu.cuStreamSynchronize_(S.stream)); // wait_until_completion = true
				}
				for (int d = 0; d < 3; ++d) {
					inv_prev[d] = 1.0f / (float)size[d];
					size[d] >>= 1;
				}
			}
		}
		return true;
	};
	for (uint32_t i = 0; i < (warmup ? warmup : 1u); ++i)
		if (!chain(true)) return false;
	CK(cu.cuStreamSynchronize_(S.stream));
	if (steps) {
		const auto t0 = std::chrono::steady_clock::now();
		for (uint32_t i = 0; i < steps; ++i)
			if (!chain(true)) return false;
		CK(cu.cuStreamSynchronize_(S.stream));
		*blocking_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / steps;
		CUevent e0, e1;
		CK(cu.cuEventCreate_(&e0, CU_EVENT_DEFAULT));
		CK(cu.cuEventCreate_(&e1, CU_EVENT_DEFAULT));
		CK(cu.cuEventRecord_(e0, S.stream));
		for (uint32_t i = 0; i < steps; ++i)
			if (!chain(false)) return false;
		CK(cu.cuEventRecord_(e1, S.stream));
		CK(cu.cuEventSynchronize_(e1));
		float ms = 0;
		CK(cu.cuEventElapsedTime_(&ms, e0, e1));
		*enqueued_ms = (double)ms / steps;
		cu.cuEventDestroy_(e0);
		cu.cuEventDestroy_(e1);
	}
	*launches_per_chain = launches;
	if (out_all) {
		uint8_t* cur = (uint8_t*)out_all;
		for (uint32_t l = 0; l < level_count; ++l) {
			uint32_t d[3];
			level_dims(l, d);
			if (!copy_level(l, false, cur)) return false;
			cur += (size_t)d[0] * d[1] * d[2] * bpp * (dc == 3 ? 1u : layers);
		}
	}
	return true;
}
} // namespace

extern "C" {
// returns 0 on success; on failure the message is in flinc_last_error()
int flinc_run(const char* ptx_path, int device, uint64_t image_type, const uint32_t image_dim[4], const void* level0, void* out_all_levels, uint32_t warmup,
			  uint32_t steps, double* blocking_ms_per_chain, double* enqueued_ms_per_chain, uint64_t* launches_per_chain) {
	double a = 0, b = 0;
	uint64_t n = 0;
	const bool ok = run(ptx_path, device, image_type, image_dim, level0, out_all_levels, warmup, steps, &a, &b, &n);
	if (blocking_ms_per_chain) *blocking_ms_per_chain = a;
	if (enqueued_ms_per_chain) *enqueued_ms_per_chain = b;
	if (launches_per_chain) *launches_per_chain = n;
	return ok ? 0 : -1;
}
const char* flinc_last_error(void) { return g_err.c_str(); }
}
