// Host side of libfloor_b200_mip.so: a thin C-ABI layer over the CUDA *driver* API (dlopen'd, like floor's
// src/device/cuda/cuda_api.cpp:46-110) that owns linear image allocations, TMA tensor maps and the launch of
// the sm_100a kernels embedded as a cubin (mip_kernels.cu).  No CUDA runtime, no CPU fallback.
#include "../../include/floor_b200_mip.h"

#include <cuda.h> // types and prototypes only -- libcuda is resolved at run time with dlopen/dlsym
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "mip_params.h"
#include "mip_tiling.h"

// the cubin produced from mip_kernels.cu, embedded by cubin_blob.S (cf. `#embed` of mmm.fubar, device_image.cpp:133-139)
extern "C" const unsigned char flmip_cubin_begin[];
extern "C" const unsigned char flmip_cubin_end[];

namespace {

// ------------------------------------------------------------------------------------------------------
// error reporting
// ------------------------------------------------------------------------------------------------------
thread_local std::string tl_error;
int fail(int code, const char* fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	tl_error = buf;
	return code;
}

// ------------------------------------------------------------------------------------------------------
// driver function table
// ------------------------------------------------------------------------------------------------------
#define FL_STR2(x) #x
#define FL_STR(x) FL_STR2(x)
#define FL_DRIVER_FUNCTIONS(F)                                                                                                 \
	F(cuInit) F(cuDriverGetVersion) F(cuGetErrorName) F(cuGetErrorString) F(cuDeviceGetCount) F(cuDeviceGet) F(cuDeviceGetName)   \
	F(cuDeviceGetAttribute) F(cuDeviceTotalMem) F(cuDevicePrimaryCtxRetain) F(cuDevicePrimaryCtxRelease) F(cuCtxPushCurrent)     \
	F(cuCtxPopCurrent) F(cuMemAlloc) F(cuMemFree) F(cuMemsetD8Async) F(cuMemsetD32Async) F(cuMemcpyHtoDAsync) F(cuMemcpyDtoHAsync) \
	F(cuMemcpy3DAsync) F(cuMemHostAlloc) F(cuMemFreeHost) F(cuStreamCreate) F(cuStreamDestroy) F(cuStreamSynchronize)            \
	F(cuEventCreate) F(cuEventRecord) F(cuEventSynchronize) F(cuEventElapsedTime) F(cuEventDestroy) F(cuModuleLoadData)          \
	F(cuModuleGetFunction) F(cuFuncSetAttribute) F(cuLaunchKernel) F(cuLaunchKernelEx) F(cuTensorMapEncodeTiled) F(cuMemcpyDtoDAsync)                 \
	F(cuMipmappedArrayCreate) F(cuMipmappedArrayGetLevel) F(cuMipmappedArrayDestroy) F(cuGraphCreate) F(cuGraphAddKernelNode)           \
	F(cuGraphInstantiate) F(cuGraphLaunch) F(cuGraphExecDestroy) F(cuGraphDestroy) F(cuStreamWaitEvent) F(cuCtxGetDevice) F(cuMemcpyHtoD)

struct driver_api {
#define FL_DECL(name) decltype(&name) p_##name = nullptr;
	FL_DRIVER_FUNCTIONS(FL_DECL)
#undef FL_DECL
	void* handle = nullptr;
};
driver_api cu;

struct device_state {
	CUdevice dev = 0;
	std::atomic<CUcontext> ctx { nullptr }; // retained on first use (double-checked under mtx)
	CUmodule module = nullptr;
	std::mutex mtx;
	std::unordered_map<std::string, CUfunction> functions;
	flmip_device_info info {};
	uint32_t smem_per_sm = 0, smem_per_block_optin = 0;
	bool ctx_attached = false;      // the context was handed in by flmip_device_attach_context (not retained, never released)
	CUstream util_stream = nullptr; // private non-blocking stream for creation-time memsets (never the legacy stream)
	std::mutex images_mtx;
	std::unordered_set<flmip_image_s*> images; // live images of this device (flmip_stream_destroy hands their chains over)
};

std::once_flag init_once;
int init_status = FLMIP_ERR_NO_CUDA;
std::string init_error = "flmip_init() has not been called";
std::vector<device_state*> devices;
std::atomic<uint64_t> launch_counter { 0 };

int cu_fail(CUresult r, const char* what) {
	const char *name = nullptr, *desc = nullptr;
	if (cu.p_cuGetErrorName) cu.p_cuGetErrorName(r, &name);
	if (cu.p_cuGetErrorString) cu.p_cuGetErrorString(r, &desc);
	return fail(r == CUDA_ERROR_OUT_OF_MEMORY ? FLMIP_ERR_OUT_OF_MEMORY : FLMIP_ERR_DRIVER, "%s: %s (%s)", what, name ? name : "?",
				desc ? desc : "?");
}
#define CU_TRY(call, what)                              \
	do {                                                \
		const CUresult cu_try_r_ = (call);              \
		if (cu_try_r_ != CUDA_SUCCESS) return cu_fail(cu_try_r_, what); \
	} while (0)

void do_init() {
	const char* names[] = { "libcuda.so.1", "libcuda.so" };
	for (const char* n : names) {
		cu.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (cu.handle) break;
	}
	if (!cu.handle) {
		init_error = std::string("failed to load libcuda: ") + dlerror();
		return;
	}
#define FL_LOAD(name)                                                                        \
	cu.p_##name = reinterpret_cast<decltype(&name)>(dlsym(cu.handle, FL_STR(name)));         \
	if (!cu.p_##name) {                                                                      \
		init_error = std::string("libcuda lacks ") + FL_STR(name) + " (driver too old?)";    \
		return;                                                                              \
	}
	FL_DRIVER_FUNCTIONS(FL_LOAD)
#undef FL_LOAD
	CUresult r = cu.p_cuInit(0);
	if (r != CUDA_SUCCESS) {
		const char* name = nullptr;
		cu.p_cuGetErrorName(r, &name);
		init_error = std::string("cuInit failed: ") + (name ? name : "?");
		return;
	}
	int version = 0;
	cu.p_cuDriverGetVersion(&version);
	int count = 0;
	if (cu.p_cuDeviceGetCount(&count) != CUDA_SUCCESS || count <= 0) {
		init_error = "no CUDA device";
		return;
	}
	for (int i = 0; i < count; ++i) {
		auto* ds = new device_state;
		if (cu.p_cuDeviceGet(&ds->dev, i) != CUDA_SUCCESS) { init_error = "cuDeviceGet failed"; return; }
		auto attr = [&](CUdevice_attribute a) {
			int v = 0;
			cu.p_cuDeviceGetAttribute(&v, a, ds->dev);
			return (uint32_t)v;
		};
		cu.p_cuDeviceGetName(ds->info.name, sizeof(ds->info.name), ds->dev);
		size_t mem = 0;
		cu.p_cuDeviceTotalMem(&mem, ds->dev);
		ds->info.global_mem_size = mem;
		ds->info.sm_major = attr(CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR);
		ds->info.sm_minor = attr(CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR);
		ds->info.units = attr(CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT);
		ds->info.max_total_local_size = attr(CU_DEVICE_ATTRIBUTE_MAX_THREADS_PER_BLOCK);
		ds->info.max_image_2d_dim[0] = attr(CU_DEVICE_ATTRIBUTE_MAXIMUM_TEXTURE2D_WIDTH);
		ds->info.max_image_2d_dim[1] = attr(CU_DEVICE_ATTRIBUTE_MAXIMUM_TEXTURE2D_HEIGHT);
		ds->info.max_image_3d_dim[0] = attr(CU_DEVICE_ATTRIBUTE_MAXIMUM_TEXTURE3D_WIDTH);
		ds->info.max_image_3d_dim[1] = attr(CU_DEVICE_ATTRIBUTE_MAXIMUM_TEXTURE3D_HEIGHT);
		ds->info.max_image_3d_dim[2] = attr(CU_DEVICE_ATTRIBUTE_MAXIMUM_TEXTURE3D_DEPTH);
		ds->info.max_mip_levels = FLMIP_MAX_LEVELS;
		ds->smem_per_sm = attr(CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_MULTIPROCESSOR);
		ds->smem_per_block_optin = attr(CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN);
		ds->info.driver_version = (uint32_t)version;
		ds->info.clock_mhz = attr(CU_DEVICE_ATTRIBUTE_CLOCK_RATE) / 1000u;
		ds->info.mem_clock_mhz = attr(CU_DEVICE_ATTRIBUTE_MEMORY_CLOCK_RATE) / 1000u;
		ds->info.mem_bus_width = attr(CU_DEVICE_ATTRIBUTE_GLOBAL_MEMORY_BUS_WIDTH);
		ds->info.l2_cache_size = attr(CU_DEVICE_ATTRIBUTE_L2_CACHE_SIZE);
		devices.push_back(ds);
	}
	init_status = FLMIP_OK;
	init_error.clear();
}

int ensure_init() {
	std::call_once(init_once, do_init);
	if (init_status != FLMIP_OK) return fail(init_status, "%s", init_error.c_str());
	return FLMIP_OK;
}

int get_device(int device, device_state** out) {
	const int rc = ensure_init();
	if (rc != FLMIP_OK) return rc;
	if (device < 0 || (size_t)device >= devices.size()) return fail(FLMIP_ERR_INVALID, "invalid device index %d (have %zu)", device, devices.size());
	device_state* ds = devices[(size_t)device];
	if (!ds->ctx.load(std::memory_order_acquire)) {
		std::lock_guard<std::mutex> lock(ds->mtx);
		if (!ds->ctx.load(std::memory_order_relaxed)) {
			// primary context: shared with any CUDA-runtime user in the process (floor creates its own: cuda_context.cpp:117-124)
			CUcontext ctx = nullptr;
			CU_TRY(cu.p_cuDevicePrimaryCtxRetain(&ctx, ds->dev), "cuDevicePrimaryCtxRetain");
			ds->ctx.store(ctx, std::memory_order_release);
		}
	}
	*out = ds;
	return FLMIP_OK;
}

// every entry point makes the device's context current for the calling thread (cf. cuda_function.cpp:83-87)
struct ctx_guard {
	bool pushed = false;
	int push(device_state* ds) {
		CU_TRY(cu.p_cuCtxPushCurrent(ds->ctx.load(std::memory_order_acquire)), "cuCtxPushCurrent");
		pushed = true;
		return FLMIP_OK;
	}
	~ctx_guard() {
		if (pushed) {
			CUcontext old = nullptr;
			cu.p_cuCtxPopCurrent(&old);
		}
	}
};
#define WITH_DEVICE(device)                                           \
	device_state* ds = nullptr;                                       \
	{                                                                 \
		const int wd_rc_ = get_device((device), &ds);                 \
		if (wd_rc_ != FLMIP_OK) return wd_rc_;                        \
	}                                                                 \
	ctx_guard guard;                                                  \
	{                                                                 \
		const int wd_rc_ = guard.push(ds);                            \
		if (wd_rc_ != FLMIP_OK) return wd_rc_;                        \
	}

int get_function(device_state* ds, const std::string& name, uint32_t dynamic_smem, CUfunction* out) {
	std::lock_guard<std::mutex> lock(ds->mtx);
	if (!ds->module) {
		if (ds->info.sm_major != 10) {
			return fail(FLMIP_ERR_UNSUPPORTED, "device '%s' is sm_%u%u; this library only carries sm_100a code", ds->info.name, ds->info.sm_major,
						ds->info.sm_minor);
		}
		CU_TRY(cu.p_cuModuleLoadData(&ds->module, flmip_cubin_begin), "cuModuleLoadData(embedded cubin)");
	}
	auto it = ds->functions.find(name);
	if (it == ds->functions.end()) {
		CUfunction fn = nullptr;
		CU_TRY(cu.p_cuModuleGetFunction(&fn, ds->module, name.c_str()), ("cuModuleGetFunction " + name).c_str());
		if (dynamic_smem > 48u * 1024u) {
			CU_TRY(cu.p_cuFuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dynamic_smem), "cuFuncSetAttribute(smem)");
		}
		it = ds->functions.emplace(name, fn).first;
	}
	*out = it->second;
	return FLMIP_OK;
}

// `dependent` = programmatic dependent launch: the kernel may become resident while its predecessor in the stream drains (its
// CTAs run their prologue, then block in griddepcontrol.wait until the predecessor has completed and flushed) -- only for
// kernels that execute griddepcontrol.wait before their first global access (flmip_fast*, flmip_tile*).  FLMIP_PDL=0 disables it.
bool pdl_enabled() {
	static const bool on = [] {
		const char* v = getenv("FLMIP_PDL");
		return !(v && v[0] == '0');
	}();
	return on;
}

// ---- overlap of chains on independent images (flmip_stream_set_chain_overlap) ---------------------------------------------------
// An "open run" of a stream = the images of the chain kernels enqueued on it since the last kernel that waited for its predecessor
// at its START (pdl_start in mip_kernels.cu).  Only kernels of the open run can still be running when the next kernel's CTAs become
// resident: a late-waiting kernel starts once the CTAs of the kernel in front of it have all started, and so on back to the kernel
// that opened the run, whose start implied that everything before it had completed and flushed.  So the first kernel of a chain may
// skip the wait at its start iff its image is not in the open run; it then waits at its end (completion keeps stream order).
// Anything else this library enqueues on the stream closes the run (run_close), as does every later kernel of a multi-kernel chain.
// Opt-in per stream, because work the caller enqueues on the stream behind the library's back cannot be seen here (the caller
// announces it with flmip_stream_fence).
struct stream_run {
	// held from the decision about a chain's first kernel until the chain is noted (flmip_mip_chain_generate_from), and by whatever closes
	// the run: the bookkeeping must see chains and closers of one stream in ONE order even when several host threads enqueue on it.  A
	// closer takes it BEFORE it enqueues its own work, so its work lands behind every chain noted so far; chains decided after the
	// close find an empty run and wait at their start, wherever they land relative to the closer's work.
	std::mutex mtx;
	bool enabled = false;
	std::vector<const void*> images; // compared by address only
};
std::mutex runs_mtx; // guards the map only
std::unordered_map<CUstream, std::shared_ptr<stream_run>> runs;
std::atomic<uint32_t> overlap_streams { 0 }; // streams that have opted in (0: every hook below returns at once)
constexpr size_t MAX_RUN_IMAGES = 64;
#ifndef FLMIP_BATCH_LANES_DEFAULT
#define FLMIP_BATCH_LANES_DEFAULT 16u // chains of a batch graph that run side by side (0: all of them)
#endif

std::shared_ptr<stream_run> run_of(CUstream stream, bool create = false) {
	if (!create && overlap_streams.load(std::memory_order_relaxed) == 0) return nullptr;
	std::lock_guard<std::mutex> lock(runs_mtx);
	auto it = runs.find(stream);
	if (it != runs.end()) return it->second;
	if (!create) return nullptr;
	return runs.emplace(stream, std::make_shared<stream_run>()).first->second;
}
void run_close(CUstream stream) {
	if (auto r = run_of(stream)) {
		std::lock_guard<std::mutex> lock(r->mtx);
		r->images.clear();
	}
}
// may the first kernel of a chain on `img` start without waiting for the kernel in front of it?  (r->mtx held)
bool run_allows_late_head(const stream_run& r, const void* img) {
	if (!r.enabled) return false;
	const std::vector<const void*>& v = r.images;
	if (v.empty() || v.size() >= MAX_RUN_IMAGES) return false; // nothing of ours in front / bound the bookkeeping
	return std::find(v.begin(), v.end(), img) == v.end();
}
// after a chain of `kernels` launches on `img`, the first of which started late (or not)  (r->mtx held).  Only a first kernel that
// waited at its start opens a new run: the later kernels of a chain on an overlapping queue release their dependents BEFORE they wait
// (mode 2 of pdl_start), so they say nothing about what has completed by the time the next kernel starts.
void run_note_chain(stream_run& r, const void* img, uint32_t kernels, bool late_head) {
	if (!r.enabled || kernels == 0) return;
	if (!late_head) r.images.clear();
	r.images.push_back(img);
}
// The chain's launch modes travel to the launchers through the calling thread: the first kernel of a chain takes the head's mode (0 or 1),
// every later one mode 2 on an overlapping queue and 0 otherwise (see pdl_start in mip_kernels.cu).
thread_local bool tl_late_head = false, tl_overlap_chain = false, tl_head_taken = false;
thread_local uint32_t tl_chain_launches = 0; // kernels the calling thread has enqueued since flmip_mip_chain_generate_from reset it
uint32_t take_launch_mode() {
	if (!tl_head_taken) {
		tl_head_taken = true;
		return tl_late_head ? 1u : 0u;
	}
	return tl_overlap_chain ? 2u : 0u;
}

// While a batch is being built (flmip_batch_create) the launches of the calling thread are recorded as kernel nodes of a CUDA
// graph instead of being issued: the nodes of one image form a chain, the chains of different images have no edges between them.
struct graph_recorder {
	CUgraph graph = nullptr;
	CUgraphNode last = nullptr; // previous node of the image being recorded
	bool has_last = false;
	uint32_t nodes = 0;
};
thread_local graph_recorder* tl_recorder = nullptr;

int launch(CUfunction fn, uint64_t grid, uint32_t block, uint32_t smem, CUstream stream, void** args, bool dependent = false) {
	if (grid == 0) return FLMIP_OK;
	if (grid > 0x7FFFFFFFull) return fail(FLMIP_ERR_INVALID, "grid of %llu blocks exceeds the launch limit", (unsigned long long)grid);
	if (tl_recorder) {
		CUDA_KERNEL_NODE_PARAMS np;
		memset(&np, 0, sizeof(np));
		np.func = fn;
		np.gridDimX = (unsigned)grid; np.gridDimY = 1; np.gridDimZ = 1;
		np.blockDimX = block; np.blockDimY = 1; np.blockDimZ = 1;
		np.sharedMemBytes = smem;
		np.kernelParams = args; // copied by the driver now
		CUgraphNode node = nullptr;
		CU_TRY(cu.p_cuGraphAddKernelNode(&node, tl_recorder->graph, tl_recorder->has_last ? &tl_recorder->last : nullptr, tl_recorder->has_last ? 1 : 0, &np),
			   "cuGraphAddKernelNode");
		tl_recorder->last = node;
		tl_recorder->has_last = true;
		++tl_recorder->nodes;
		return FLMIP_OK;
	}
	if (dependent && pdl_enabled()) {
		CUlaunchAttribute attr;
		memset(&attr, 0, sizeof(attr));
		attr.id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
		attr.value.programmaticStreamSerializationAllowed = 1;
		CUlaunchConfig cfg;
		memset(&cfg, 0, sizeof(cfg));
		cfg.gridDimX = (unsigned)grid; cfg.gridDimY = 1; cfg.gridDimZ = 1;
		cfg.blockDimX = block; cfg.blockDimY = 1; cfg.blockDimZ = 1;
		cfg.sharedMemBytes = smem;
		cfg.hStream = stream;
		cfg.attrs = &attr;
		cfg.numAttrs = 1;
		CU_TRY(cu.p_cuLaunchKernelEx(&cfg, fn, args, nullptr), "cuLaunchKernelEx");
	} else {
		CU_TRY(cu.p_cuLaunchKernel(fn, (unsigned)grid, 1, 1, block, 1, 1, smem, stream, args, nullptr), "cuLaunchKernel");
	}
	launch_counter.fetch_add(1, std::memory_order_relaxed);
	++tl_chain_launches;
	return FLMIP_OK;
}

// ------------------------------------------------------------------------------------------------------
// IMAGE_TYPE decoding (bit layout: include/floor/device/backend/image_types.hpp:24-236)
// ------------------------------------------------------------------------------------------------------
constexpr uint64_t T_FORMAT_MASK = 0x3Full, T_COMPRESSION_MASK = 0x3C0ull, T_DATA_TYPE_MASK = 0x3000ull;
constexpr uint64_t T_INT = 0x1000ull, T_UINT = 0x2000ull, T_FLOAT = 0x3000ull;
constexpr uint64_t T_FLAG_ARRAY = 1ull << 20, T_FLAG_MSAA = 1ull << 22, T_FLAG_CUBE = 1ull << 23, T_FLAG_DEPTH = 1ull << 24,
				   T_FLAG_STENCIL = 1ull << 25, T_FLAG_MIPMAPPED = 1ull << 27, T_FLAG_NORMALIZED = 1ull << 30;
constexpr uint32_t FMT_2 = 2, FMT_4 = 4, FMT_8 = 11, FMT_16 = 18, FMT_32 = 22;

bool is_pot(uint32_t v) { return v != 0 && (v & (v - 1)) == 0; }

} // namespace

struct flmip_image_s {
	int device = 0;
	uint64_t type = 0;
	uint32_t dim[4] = { 0, 0, 0, 0 };
	uint32_t dc = 0, channels = 0, bpc = 0, bpp = 0, layers = 0, level_count = 0, elem_kind = 0, no_double = 0, sm_count = 0;
	flmip_level_info levels[FLMIP_MAX_LEVELS] {};
	uint64_t total_size = 0;
	CUdeviceptr mem = 0, counters = 0;
	// single-pass plan
	bool fast = false;
	uint32_t fast_level_count = 0; // levels [0, fast_level_count) are produced by the single-pass launch
	flmip_fast_params fast_params {};
	flmip_tiling_rt tiling {};
	uint32_t fast_smem = 0, fast_grid = 0; // dynamic shared memory and persistent grid of the single-pass launch
	alignas(64) CUtensorMap tmap {};
	std::string fast_name;
	// levels the single-pass launch does not produce: multi-level tile kernel (2D / 3D) or one generic launch per level
	bool tiled = false;
	std::string tile_name;
	// persistent TMA tile kernel (flmip_ptile2d_*): sampler table, per-layer counters + scheduler words, one tensor map per source level
	bool ptile = false;
	uint32_t ptile_flags = 0; // FLMIP_IMAGE_TMA_TILES_* tuning overrides
	std::string ptile_name;
	CUdeviceptr wtab = 0, pcounters = 0;
	uint32_t wtab_off[FLMIP_MAX_LEVELS][2] = {};
	bool texel2[FLMIP_MAX_LEVELS][2] = {}; // destination level / axis whose texel 0 takes the reference's texel-2 fetch
	uint32_t ptile_smem = 0;
	CUtensorMap ptile_map[FLMIP_MAX_LEVELS];
	bool ptile_map_valid[FLMIP_MAX_LEVELS] = {};
	bool external_mem = false; // `mem` belongs to the caller (flmip_image_create_external): never freed here
	// One chain per image may be in flight at a time: the group / layer / scheduler counters beside the image are shared by every
	// launch on it.  Chains on ONE stream are ordered by the stream; when a chain is enqueued on a different stream than the
	// previous one, that stream is first made to wait for everything enqueued on the old stream so far (event hand-over), so
	// overlapping chains on one image serialise instead of corrupting each other.
	std::mutex gen_mtx;
	CUstream last_stream = nullptr;
	bool has_last = false, pending_handover = false;
	CUevent handover = nullptr;
};

namespace {

// image_types.hpp:694-712
uint32_t mip_level_count_for(const uint32_t dim[4], uint64_t type, uint32_t dc) {
	if (!(type & T_FLAG_MIPMAPPED)) return 1;
	uint32_t m = dim[0];
	if (dc >= 2 && dim[1] > m) m = dim[1];
	if (dc >= 3 && dim[2] > m) m = dim[2];
	if (m <= 1) return 1;
	uint32_t n = 0;
	while (m) { ++n; m >>= 1; }
	return n; // 32 - clz(prev_pot(max_dim))
}

int decode_type(flmip_image_s& im) {
	const uint64_t t = im.type;
	im.dc = (uint32_t)((t >> 16) & 3u);
	im.channels = (uint32_t)((t >> 14) & 3u) + 1u;
	const uint32_t fmt = (uint32_t)(t & T_FORMAT_MASK);
	im.bpc = fmt == FMT_2 ? 2 : fmt == FMT_4 ? 4 : fmt == FMT_8 ? 8 : fmt == FMT_16 ? 16 : fmt == FMT_32 ? 32 : 0;
	if (im.dc < 1 || im.dc > 3) return fail(FLMIP_ERR_INVALID, "invalid image dimensionality in type %#llx", (unsigned long long)t);
	if (t & T_COMPRESSION_MASK) return fail(FLMIP_ERR_UNSUPPORTED, "compressed images cannot be minified (device_image.hpp:519-525)");
	if (t & T_FLAG_MSAA) return fail(FLMIP_ERR_UNSUPPORTED, "msaa is not supported (mip_map_minify.hpp:95)");
	if (t & T_FLAG_STENCIL) return fail(FLMIP_ERR_UNSUPPORTED, "stencil images have no minification kernel");
	if (im.bpc == 0) return fail(FLMIP_ERR_UNSUPPORTED, "unsupported image format %u (2/4/8/16/32 bit per channel only)", fmt);
	// 3-channel images: the reference's CUDA backend rejects them because a CUarray cannot hold them (cuda_image.cpp:173-180); its
	// Host-Compute backend minifies them, and so does this path: linear memory has no such restriction (literal kernel).
	if (im.dc == 3 && (t & (T_FLAG_ARRAY | T_FLAG_CUBE))) return fail(FLMIP_ERR_UNSUPPORTED, "3D array images have no minification kernel");
	const uint64_t dt = t & T_DATA_TYPE_MASK;
	const bool norm = (t & T_FLAG_NORMALIZED) != 0;
	if (im.bpc < 8) {
		// FORMAT_2 / FORMAT_4 (host_image.hpp:341-353, 419-446, dispatched at :1167-1186): normalized only, and only where a texel is
		// a whole number of bytes -- the reference sizes images by bits (image_types.hpp:675-691) but addresses texels by
		// ceil(bits / 8) bytes (host_image.hpp:235-271), so R2 / RG2 / RGB2 / R4 / RGB4 images are smaller than what its own kernels touch
		if (!norm || (dt != T_UINT && dt != T_INT)) return fail(FLMIP_ERR_UNSUPPORTED, "2 / 4-bit formats only exist as normalized integers");
		if ((im.bpc * im.channels) % 8u) return fail(FLMIP_ERR_UNSUPPORTED, "%u x %u-bit texels are no whole number of bytes (the reference's own size helpers and kernels disagree on such images)", im.channels, im.bpc);
		if (t & T_FLAG_DEPTH) return fail(FLMIP_ERR_UNSUPPORTED, "only D32F depth images can be minified");
		im.elem_kind = im.bpc == 4 ? (dt == T_INT ? FLMIP_EK_SNORM4 : FLMIP_EK_UNORM4) : (dt == T_INT ? FLMIP_EK_SNORM2 : FLMIP_EK_UNORM2);
		im.bpp = im.bpc * im.channels / 8u;
		return FLMIP_OK;
	}
	if (dt == T_FLOAT) {
		if (im.bpc == 32) im.elem_kind = FLMIP_EK_F32;
		else if (im.bpc == 16) im.elem_kind = FLMIP_EK_F16;
		else return fail(FLMIP_ERR_UNSUPPORTED, "8-bit float formats do not exist");
	} else if (dt == T_UINT || dt == T_INT) {
		const bool s = dt == T_INT;
		if (norm) {
			if (im.bpc == 8) im.elem_kind = s ? FLMIP_EK_SNORM8 : FLMIP_EK_UNORM8;
			else if (im.bpc == 16) im.elem_kind = s ? FLMIP_EK_SNORM16 : FLMIP_EK_UNORM16;
			else return fail(FLMIP_ERR_UNSUPPORTED, "32-bit normalized formats are not supported on CUDA");
		} else {
			im.elem_kind = im.bpc == 8 ? (s ? FLMIP_EK_I8 : FLMIP_EK_U8) : im.bpc == 16 ? (s ? FLMIP_EK_I16 : FLMIP_EK_U16) : (s ? FLMIP_EK_I32 : FLMIP_EK_U32);
		}
	} else {
		return fail(FLMIP_ERR_INVALID, "image type %#llx has no data type", (unsigned long long)t);
	}
	if (t & T_FLAG_DEPTH) {
		// only libfloor_mip_map_minify_IMAGE_DEPTH[_ARRAY]_FLOAT exist (mip_map_minify.hpp:22-30)
		if (!(im.elem_kind == FLMIP_EK_F32 && im.channels == 1)) return fail(FLMIP_ERR_UNSUPPORTED, "only D32F depth images can be minified");
	}
	im.bpp = im.bpc / 8u * im.channels;
	return FLMIP_OK;
}

// levels a cascade adds when it starts with a region of (w, h, d) texels at `lvl`
uint32_t simulate_cascade(uint32_t w, uint32_t h, uint32_t d, bool is3d, uint32_t lvl, uint32_t level_count) {
	while (lvl + 1 < level_count && w >= 2 && h >= 2 && (!is3d || d >= 2)) {
		w >>= 1; h >>= 1; if (is3d) d >>= 1;
		++lvl;
	}
	return lvl;
}

bool next_level_has_texels(const flmip_image_s& im, uint32_t lvl) {
	const uint32_t n = lvl + 1;
	if (n >= im.level_count) return false;
	const flmip_level_info& li = im.levels[n];
	return li.dim[0] != 0 && (im.dc < 2 || li.dim[1] != 0) && (im.dc < 3 || li.dim[2] != 0);
}

uint32_t env_u32(const char* name, uint32_t def) {
	const char* v = getenv(name);
	return v && *v ? (uint32_t)strtoul(v, nullptr, 10) : def;
}

// decides whether the single-pass kernel applies and how far it gets; mirrors fast_body() in mip_kernels.cu
int plan_fast(flmip_image_s& im, device_state* ds, uint32_t flags) {
	im.fast = false;
	if (flags & FLMIP_IMAGE_FORCE_GENERIC) return FLMIP_OK;
	if (im.level_count < 2 || im.dc < 2) return FLMIP_OK;
	if (im.channels == 3 || im.elem_kind >= FLMIP_EK_COUNT) return FLMIP_OK; // literal kernel only
	const bool is3d = im.dc == 3;
	const uint32_t W = im.dim[0], H = im.dim[1], D = is3d ? im.dim[2] : 1u;
	if (!is_pot(W) || !is_pot(H) || !is_pot(D)) return FLMIP_OK;
	const flmip_tiling_rt tl = flmip_tiling_lookup(im.bpp, im.dc);
	if (W < tl.tx || H < tl.ty || D < tl.tz) return FLMIP_OK;
	if ((uint64_t)W * im.bpp >= (1ull << 32) * 4ull) return FLMIP_OK;

	flmip_fast_params& P = im.fast_params;
	memset(&P, 0, sizeof(P));
	P.base = im.mem;
	for (uint32_t l = 0; l < FLMIP_MAX_LEVELS; ++l) P.level_off[l] = l < im.level_count ? im.levels[l].offset : 0;
	P.dim[0] = W; P.dim[1] = H; P.dim[2] = D;
	P.tiles[0] = W / tl.tx; P.tiles[1] = H / tl.ty; P.tiles[2] = D / tl.tz;
	// units of 2 x 2 (x 2) tiles when every dim has at least two tiles and there are enough tiles to keep every
	// resident CTA busy with whole units (small images stay latency-bound: one tile per CTA then)
	const uint64_t all_tiles = (uint64_t)P.tiles[0] * P.tiles[1] * P.tiles[2] * im.layers;
	const bool units_possible = P.tiles[0] >= 2 && P.tiles[1] >= 2 && (!is3d || P.tiles[2] >= 2);
	// Units cut the publishes (one acq_rel atomic round trip each) by 4 - 8x, but they also coarsen the work the persistent CTAs
	// draw from the scheduler.  Measured over image sizes (scripts/units_ab.py, profiles/r1/07_units_sweep*.txt): with the
	// scheduler prefetching only 2 units they are as good as or better than single tiles from about one unit per three CTAs on
	// (below that single tiles use more SMs), except where there is more than one wave and the last one is mostly empty: with
	// U work items for R resident CTAs the wave efficiency is (U / R) / ceil(U / R); single tiles are taken when theirs is
	// more than 15 % better.
	const uint64_t n_units = all_tiles >> im.dc, resident_ctas = 2ull * ds->info.units;
	auto wave_efficiency = [&](uint64_t items) {
		const uint64_t waves = (items + resident_ctas - 1u) / resident_ctas;
		return waves ? (double)items / (double)(waves * resident_ctas) : 0.0;
	};
	const bool units_pay = n_units * 3ull >= resident_ctas && (n_units <= resident_ctas || wave_efficiency(n_units) * 1.15 >= wave_efficiency(all_tiles));
	const bool units_wanted = (flags & FLMIP_IMAGE_UNITS_ALWAYS) || (!(flags & FLMIP_IMAGE_UNITS_NEVER) && units_pay);
	P.unit_shift = (units_possible && units_wanted) ? 1u : 0u;
	for (int i = 0; i < 3; ++i) {
		P.units[i] = (i == 2 && !is3d) ? 1u : P.tiles[i] >> P.unit_shift;
		P.unit_cshift[i] = (uint32_t)flmip_ilog2(P.units[i]);
		P.groups[i] = (P.units[i] + tl.group - 1) / tl.group;
	}
	const uint64_t total_tiles = (uint64_t)P.tiles[0] * P.tiles[1] * P.tiles[2] * im.layers;
	if (total_tiles > 0x7FFFFFFFull) return FLMIP_OK; // general path
	P.total_units = (uint32_t)(total_tiles >> (P.unit_shift * im.dc));
	P.layers = im.layers;
	P.no_double = im.no_double;

	// tile stage
	uint32_t lvl = simulate_cascade(tl.tx, tl.ty, tl.tz, is3d, 0, im.level_count);
	uint32_t covered = lvl;
	uint32_t rw = tl.tx >> lvl, rh = tl.ty >> lvl, rd = tl.tz >> lvl; // remainder of one tile, then of one unit
	bool more = next_level_has_texels(im, lvl);
	if (more && P.unit_shift) {
		// unit stage
		rw *= 2u; rh *= 2u; if (is3d) rd *= 2u;
		const uint32_t l2 = simulate_cascade(rw, rh, rd, is3d, lvl, im.level_count);
		rw >>= (l2 - lvl); rh >>= (l2 - lvl); if (is3d) rd >>= (l2 - lvl);
		covered = lvl = l2;
		more = next_level_has_texels(im, lvl);
	}
	if (more) {
		// group stage
		const uint32_t ntx = P.units[0] < tl.group ? P.units[0] : tl.group, nty = P.units[1] < tl.group ? P.units[1] : tl.group,
					   ntz = P.units[2] < tl.group ? P.units[2] : tl.group;
		lvl = simulate_cascade(rw * ntx, rh * nty, rd * ntz, is3d, lvl, im.level_count);
		covered = lvl;
		if (next_level_has_texels(im, lvl)) {
			// layer stage: the whole level must fit into the cascade scratch, else the general path finishes the chain
			const uint64_t patch = (uint64_t)(W >> lvl) * (H >> lvl) * (is3d ? (D >> lvl) : 1u) * im.bpp;
			if (patch <= tl.cascade_bytes) covered = simulate_cascade(W >> lvl, H >> lvl, is3d ? D >> lvl : 1u, is3d, lvl, im.level_count);
		}
	}
	im.fast_level_count = covered + 1;
	P.level_count = im.fast_level_count;
	im.tiling = tl;

	// counters: one per tile group and one per layer, zeroed once; the kernel resets what it uses
	const uint64_t n_groups = (uint64_t)im.layers * P.groups[0] * P.groups[1] * P.groups[2];
#ifdef FLMIP_TIMELINE
	const uint64_t n_counters = n_groups + im.layers + 2u + 4u + 2u * 8u * 1024u; // + 8 x u64 per CTA behind the scheduler words (tuning builds)
#else
	const uint64_t n_counters = n_groups + im.layers + 2u /* scheduler */;
#endif
	CU_TRY(cu.p_cuMemAlloc(&im.counters, n_counters * sizeof(uint32_t)), "cuMemAlloc(counters)");
	{
		std::lock_guard<std::mutex> lock(ds->mtx);
		if (!ds->util_stream) CU_TRY(cu.p_cuStreamCreate(&ds->util_stream, CU_STREAM_NON_BLOCKING), "cuStreamCreate(util)");
		CU_TRY(cu.p_cuMemsetD32Async(im.counters, 0, n_counters, ds->util_stream), "cuMemsetD32Async(counters)");
		CU_TRY(cu.p_cuStreamSynchronize(ds->util_stream), "cuStreamSynchronize(util)");
	}
	P.counters = im.counters;
	P.sched = im.counters + (n_groups + im.layers) * sizeof(uint32_t);

	// TMA descriptor over level 0: rank 3 in uint32 units -- (x, y, layer) for 2D / array / cube, (x, y, z) for volumes
	const cuuint64_t row_bytes = (cuuint64_t)W * im.bpp;
	cuuint64_t gdim[3] = { row_bytes / 4u, H, is3d ? D : im.layers };
	cuuint64_t gstride[2] = { row_bytes, row_bytes * H };
	cuuint32_t box[3] = { tl.tile_bytes_x / 4u, tl.ty, is3d ? tl.tz : 1u };
	cuuint32_t estride[3] = { 1, 1, 1 };
	CU_TRY(cu.p_cuTensorMapEncodeTiled(&im.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, reinterpret_cast<void*>(im.mem), gdim, gstride, box, estride,
									   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
									   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE),
		   "cuTensorMapEncodeTiled");

	// persistent launch shape: CTAs per SM x ring depth that fit the SM's shared memory (1 KiB is reserved per CTA).
	// Defaults: 2 CTAs per SM x 2 stages (128 KiB of loads in flight per SM: more only adds queueing latency to every
	// fence / atomic round trip of the finishers); FLMIP_CTAS_PER_SM / FLMIP_STAGES override for tuning.
	// (the overrides are read once per process)
	static const uint32_t env_ctas_per_sm = env_u32("FLMIP_CTAS_PER_SM", 2), env_stages = env_u32("FLMIP_STAGES", 2);
	uint32_t ctas_per_sm = env_ctas_per_sm;
	if (ctas_per_sm < 1) ctas_per_sm = 1;
	const uint32_t per_cta = ds->smem_per_sm / ctas_per_sm - 1024u - 2048u /* static: mbarriers, ticket, lock, unit patches */;
	const uint32_t budget = per_cta < ds->smem_per_block_optin - 2048u ? per_cta : ds->smem_per_block_optin - 2048u;
	if (budget < tl.cascade_smem_bytes + tl.tile_bytes) return fail(FLMIP_ERR_INVALID, "FLMIP_CTAS_PER_SM=%u leaves no room for a tile", ctas_per_sm);
	uint32_t stages = (budget - tl.cascade_smem_bytes) / tl.tile_bytes;
	const uint32_t want = env_stages;
	if (stages > want) stages = want;
	if (stages > FLMIP_MAX_STAGES) stages = FLMIP_MAX_STAGES;
	if (stages < 1) stages = 1;
	P.stages = stages;
	im.fast_smem = stages * tl.tile_bytes + tl.cascade_smem_bytes;
	const uint64_t resident = (uint64_t)ds->info.units * ctas_per_sm;
	im.fast_grid = (uint32_t)(P.total_units < resident ? P.total_units : resident);

	char name[64];
	snprintf(name, sizeof(name), "flmip_fast%ud_k%u_c%u", im.dc, im.elem_kind, im.channels);
	im.fast_name = name;
	CUfunction fn = nullptr;
	const int rc = get_function(ds, im.fast_name, im.fast_smem, &fn); // resolve now: fail at creation, not at first use
	if (rc != FLMIP_OK) return rc;
	im.fast = true;
	return FLMIP_OK;
}

int launch_generic_level(flmip_image_s& im, device_state* ds, uint32_t level, CUstream stream) {
	const flmip_level_info &src = im.levels[level - 1], &dst = im.levels[level];
	flmip_generic_params G;
	memset(&G, 0, sizeof(G));
	uint64_t texels = 1;
	for (uint32_t d = 0; d < im.dc; ++d) {
		if (dst.dim[d] == 0) return FLMIP_OK; // empty level (zero dim quirk): the reference launches zero work-items
		texels *= dst.dim[d];
	}
	G.base = im.mem;
	G.src_off = src.offset; G.dst_off = dst.offset;
	G.src_slice = src.slice_size; G.dst_slice = dst.slice_size;
	G.total = texels * im.layers;
	for (uint32_t d = 0; d < 3; ++d) {
		G.src_dim[d] = src.dim[d]; G.dst_dim[d] = dst.dim[d];
		// device_image.cpp:311-312 and host_image.cpp:96-107, evaluated in IEEE fp32 on the host
		G.inv_prev[d] = 1.0f / (float)src.dim[d];
		G.fdim[d] = src.dim[d] > 0 ? (float)src.dim[d] : 0.0f;
		G.fdim_excl[d] = src.dim[d] > 0 ? nextafterf((float)src.dim[d], 0.0f) : 0.0f;
	}
	G.dc = im.dc; G.layers = im.layers; G.elem_kind = im.elem_kind; G.channels = im.channels; G.no_double = im.no_double;
	CUfunction fn = nullptr;
	const int rc = get_function(ds, "flmip_generic", 0, &fn);
	if (rc != FLMIP_OK) return rc;
	void* args[] = { &G };
	return launch(fn, (G.total + 255u) / 256u, 256, 0, stream, args);
}

// first level without texels (zero dim quirk, image_types.hpp:751-766): every later level is empty as well
uint32_t populated_levels(const flmip_image_s& im) {
	uint32_t n = 0;
	while (n < im.level_count && im.levels[n].size != 0) ++n;
	return n;
}

// The reference's sampler quirk the tile kernel has to honour (see axis_fetch in mip_kernels.cu): for destination texel 0
// of a source level of N texels with fl(fl(1/N) * N) == pred(1.0f), the neighbour texel is 2, not 1.  Same IEEE fp32
// operations as the device code (host_image.hpp:141-174, 869-894 with g = 0).
bool axis_reads_texel_2(uint32_t n) {
	if (n < 3) return false;
	const float fn = (float)n;
	volatile float coord = 1.0f * (1.0f / fn);
	volatile float scaled = coord * fn;
	const float frac = scaled - floorf(scaled);
	volatile float ma = scaled + (frac < 0.5f ? -1.0f : 1.0f);
	return ma >= 2.0f;
}

// levels one tile-kernel launch can produce from source level s: up to 6 (2D) / 4 (3D), cut before a level whose
// texel-2 fetch would leave the 2-texel-wide remainder of a tile (produced level k reads level k - 1 of the tile,
// which is tile >> (k - 1) texels wide)
uint32_t tile_levels_from(const flmip_image_s& im, uint32_t s, uint32_t pop) {
	const uint32_t tile[3] = { im.dc == 3 ? FLMIP_TILE3D_X : FLMIP_TILE2D_X, im.dc == 3 ? FLMIP_TILE3D_Y : FLMIP_TILE2D_Y, FLMIP_TILE3D_Z };
	const uint32_t step = im.dc == 3 ? FLMIP_TILE3D_MAX_LEVELS : FLMIP_TILE_MAX_LEVELS;
	uint32_t n = pop - 1u - s < step ? pop - 1u - s : step;
	for (uint32_t k = 2; k <= n; ++k) {
		for (uint32_t d = 0; d < im.dc; ++d) {
			if ((tile[d] >> (k - 1u)) < 3u && axis_reads_texel_2(im.levels[s + k - 1u].dim[d])) return k - 1u;
		}
	}
	return n;
}

// number of tile-kernel launches that produce levels (src_level, populated)
// what one launch of the persistent tile kernel produces from source level `src`
struct ptile_step {
	bool ok = false;
	uint32_t tile_last = 0, last = 0;
};
ptile_step plan_ptile_step(const flmip_image_s& im, uint32_t src, uint32_t pop);
int launch_ptile(flmip_image_s& im, device_state* ds, uint32_t src, const ptile_step& st, CUstream stream);
uint32_t tile_launch_count(const flmip_image_s& im, uint32_t src_level, uint32_t* tma_launches);

// multi-level tile kernel: levels src_level + 1 ... (stream-ordered launches of up to 6 (2D) / 4 (3D) levels each)
int launch_tile_levels(flmip_image_s& im, device_state* ds, uint32_t src_level, CUstream stream) {
	const uint32_t pop = populated_levels(im);
	CUfunction fn = nullptr;
	uint32_t step = 0;
	for (uint32_t s = src_level; s + 1 < pop; s += step) {
		// the persistent TMA kernel where the level qualifies as a source (it may finish the whole chain) ...
		const ptile_step pst = plan_ptile_step(im, s, pop);
		if (pst.ok) {
			const int rc = launch_ptile(im, ds, s, pst, stream);
			if (rc != FLMIP_OK) return rc;
			step = pst.last - s;
			continue;
		}
		// ... else the LDG tile kernel: any size, up to 6 (2D) / 4 (3D) levels per launch
		step = tile_levels_from(im, s, pop);
		if (!fn) {
			const int rc = get_function(ds, im.tile_name, 0, &fn);
			if (rc != FLMIP_OK) return rc;
		}
		flmip_tile_params T;
		memset(&T, 0, sizeof(T));
		T.base = im.mem;
		T.nlev = step;
		for (uint32_t k = 0; k <= T.nlev; ++k) {
			const flmip_level_info& li = im.levels[s + k];
			T.level_off[k] = li.offset;
			T.slice[k] = li.slice_size;
			T.dim[k][0] = li.dim[0]; T.dim[k][1] = li.dim[1]; T.dim[k][2] = im.dc == 3 ? li.dim[2] : 1u;
			if (k < T.nlev) {
				for (uint32_t d = 0; d < im.dc; ++d) {
					// device_image.cpp:311-312 and host_image.cpp:96-107, evaluated in IEEE fp32 on the host
					T.inv_prev[k][d] = 1.0f / (float)li.dim[d];
					T.fdim[k][d] = (float)li.dim[d];
					T.fdim_excl[k][d] = nextafterf((float)li.dim[d], 0.0f);
				}
			}
		}
		const uint32_t tx = im.dc == 3 ? FLMIP_TILE3D_X : FLMIP_TILE2D_X, ty = im.dc == 3 ? FLMIP_TILE3D_Y : FLMIP_TILE2D_Y;
		T.tiles[0] = (T.dim[0][0] + tx - 1u) / tx;
		T.tiles[1] = (T.dim[0][1] + ty - 1u) / ty;
		T.tiles[2] = im.dc == 3 ? (T.dim[0][2] + FLMIP_TILE3D_Z - 1u) / FLMIP_TILE3D_Z : 1u;
		T.layers = im.layers;
		T.no_double = im.no_double;
		T.late_wait = take_launch_mode();
		for (uint32_t k = 2; k <= T.nlev; ++k)
			for (uint32_t d = 0; d < im.dc; ++d)
				if (axis_reads_texel_2(im.levels[s + k - 1u].dim[d])) T.block_sync = 1u;
		void* args[] = { &T };
		const int rc = launch(fn, (uint64_t)T.tiles[0] * T.tiles[1] * T.tiles[2] * im.layers, 256, 0, stream, args, true);
		if (rc != FLMIP_OK) return rc;
	}
	return FLMIP_OK;
}

// ---- persistent TMA tile kernel: sampler table and launch plan --------------------------------------------------------
// One table entry: the reference's linear fetch for destination texel g along one axis of a source level of n texels
// (mip_map_minify.hpp:106, host_image.hpp:141-174, 869-894), evaluated with the same IEEE fp32 operations as axis_fetch() /
// generic_texel() in mip_kernels.cu (this file is built with -fno-fast-math -ffp-contract=off).  Returns false if the fetch is
// none of the three shapes the table can express (never observed: tests/test_npot_weights.py).
bool sampler_entry(uint32_t g, uint32_t n, uint32_t* out) {
	const volatile float fdim = (float)n;
	const volatile float inv_prev = 1.0f / fdim;                     // device_image.cpp:311-312
	const volatile float fdim_excl = nextafterf(fdim, 0.0f);           // host_image.cpp:102-107
	const volatile float coord = (float)(g * 2u + 1u) * inv_prev;
	if (!(coord >= 0.0f && coord < 1.0f)) return false;              // wrap(coord, 1) would not be the identity
	const volatile float m = coord * fdim;
	const volatile float frac = m - floorf(m);
	const bool lo = frac < 0.5f;
	volatile float t = lo ? frac + 0.5f : 1.5f - frac;
	volatile float ma = m + (lo ? -1.0f : 1.0f);
	const float mb = m > fdim_excl ? fdim_excl : m;
	const float mac = ma > fdim_excl ? fdim_excl : (ma < 0.0f ? 0.0f : ma);
	const uint32_t b = (uint32_t)(long long)mb, a = (uint32_t)(long long)mac;
	uint32_t bits;
	const float tv = t;
	memcpy(&bits, &tv, 4);
	if (bits == 0u || (bits & 0xC0000000u)) return false;           // 0 < t < 2 keeps the two flag bits free
	if (a == 2u * g && b == 2u * g + 1u) *out = bits;
	else if (a == 2u * g + 1u && b == 2u * g) *out = bits | FLMIP_WTAB_SWAP;
	else if (g == 0u && a == 2u && b == 0u) *out = bits | FLMIP_WTAB_TEXEL2;
	else return false;
	return true;
}

constexpr uint32_t WTAB_PAD = 512u; // consumers of a border tile read entries past the level's extent (results masked)

int build_sampler_table(flmip_image_s& im, device_state* ds) {
	(void)ds;
	std::vector<uint32_t> tab;
	const uint32_t pop = populated_levels(im);
	for (uint32_t L = 1; L < pop; ++L) {
		for (uint32_t d = 0; d < 2; ++d) {
			while (tab.size() % 4u) tab.push_back(0x3F000000u); // segments start on 16-byte boundaries
			im.wtab_off[L][d] = (uint32_t)tab.size();
			const uint32_t n_dst = im.levels[L].dim[d], n_src = im.levels[L - 1].dim[d];
			for (uint32_t g = 0; g < n_dst; ++g) {
				uint32_t e = 0;
				if (!sampler_entry(g, n_src, &e)) return fail(FLMIP_ERR_UNSUPPORTED, "irregular sampler fetch at level %u axis %u texel %u", L, d, g);
				if (e & FLMIP_WTAB_TEXEL2) im.texel2[L][d] = true;
				tab.push_back(e);
			}
			for (uint32_t i = 0; i < WTAB_PAD; ++i) tab.push_back(0x3F000000u);
		}
	}
	if (tab.empty()) return fail(FLMIP_ERR_INVALID, "no levels to generate");
	CU_TRY(cu.p_cuMemAlloc(&im.wtab, tab.size() * sizeof(uint32_t)), "cuMemAlloc(sampler table)");
	CU_TRY(cu.p_cuMemcpyHtoD(im.wtab, tab.data(), tab.size() * sizeof(uint32_t)), "cuMemcpyHtoD(sampler table)");
	return FLMIP_OK;
}

ptile_step plan_ptile_step(const flmip_image_s& im, uint32_t src, uint32_t pop) {
	ptile_step st;
	if (!im.ptile || src + 1u >= pop) return st;
	const flmip_level_info& ls = im.levels[src];
	const uint64_t pitch = (uint64_t)ls.dim[0] * im.bpp;
	// TMA: 16-byte aligned base and row pitch; a whole tile row / column must exist (smaller images are latency-bound anyway)
	if (((im.mem + ls.offset) & 15u) || (pitch & 15u) || pitch < im.tiling.tile_bytes_x || ls.dim[1] < im.tiling.ty) return st;
	if (pitch >= (1ull << 32) * 4ull) return st;
	// levels src + 1 and src + 2 are produced in registers from texels {2g, 2g + 1}: no texel-2 fetch there
	for (uint32_t k = src + 1u; k <= src + 2u && k < pop; ++k)
		if (im.texel2[k][0] || im.texel2[k][1]) return st;
	const uint64_t tiles = (uint64_t)((ls.dim[0] + im.tiling.tx - 1u) / im.tiling.tx) * ((ls.dim[1] + im.tiling.ty - 1u) / im.tiling.ty) * im.layers;
	if (tiles > 0x7FFFFFFFull) return st;
	// When does the persistent kernel pay off?  Measured against the LDG tile kernel over sizes and formats (scripts/ptile_sweep.py,
	// profiles/r2/05_ptile_plan_sweep.txt), per resident CTA (2 per SM):
	//  * its launch has a fixed cost of ~15 us (ring ramp-up, 3.4 tiles per CTA already take 13 us: scripts/timeline_ptile.py) against
	//    ~5 us of the LDG kernel, so a level with few tiles per CTA is better off there (3840 x 2160 RGBA8: 21 us against 31 .. 38 us);
	//  * streaming, it beats the LDG kernel the more the narrower the texels are: R8 2 216 against 1 008 GB/s, RGBA8 3 765 against
	//    2 794 GB/s, RGBA16F 6 950 against 5 650 GB/s (levels 1 - 2 of 64 x 1920 x 1080), while 16-byte texels already run at the copy
	//    peak with plain loads (7 078 GB/s) -- never taken for those.
	const uint64_t resident_ctas = 2ull * im.sm_count;
	static const uint32_t env_min = env_u32("FLMIP_PTILE_MIN_TILES_PER_CTA", 0);
	const uint32_t min_tiles = env_min ? env_min : (im.bpp == 1 ? 3u : im.bpp == 2 ? 6u : im.bpp == 4 ? 12u : im.bpp == 8 ? 48u : 0xFFFFFFFFu);
	if (!(im.ptile_flags & FLMIP_IMAGE_TMA_TILES_ALWAYS) && (min_tiles == 0xFFFFFFFFu || tiles < (uint64_t)min_tiles * resident_ctas)) return st;
	uint32_t tile_last = src + im.tiling.tile_levels < pop - 1u ? src + im.tiling.tile_levels : pop - 1u;
	// a texel-2 fetch needs 3 texels of the previous level inside the tile's part of it
	for (uint32_t k = src + 3u; k <= tile_last; ++k) {
		const uint32_t w = im.tiling.tx >> (k - 1u - src), h = im.tiling.ty >> (k - 1u - src);
		if ((w < 3u && im.texel2[k][0]) || (h < 3u && im.texel2[k][1])) { tile_last = k - 1u; break; }
	}
	st.ok = true;
	// By default a launch streams levels src + 1 and src + 2 only: the consumers produce both in registers, the finisher pool stays
	// idle and the kernel runs at the roofline (N2: 0.204 ms for 98.4 % of the bytes; with the pool reducing every tile further and
	// the last tile of a layer finishing the chain it is 0.27 ms, finisher-bound).  The rest of the chain, 1 / 16 of the texels, goes
	// through this planner again from level src + 2 (usually: the LDG kernel).  FLMIP_IMAGE_TMA_TILES_NO_SPLIT selects the
	// one-launch form (validation, tuning).
	const bool split = !(im.ptile_flags & FLMIP_IMAGE_TMA_TILES_NO_SPLIT);
	if (split && src + 2u < pop - 1u) {
		st.tile_last = st.last = src + 2u;
		// (Also producing level src + 3 in the consumers -- shuffles between the four threads that hold its source texels -- was built and
		// measured for 8-byte texels: bit-exact, but N2 0.2305 -> 0.2400 ms: the quarter-occupied warps cost more than the 4x smaller
		// read of the next launch saves; profiles/r2/04_ptile_experiments.txt.)
		return st;
	}
	st.tile_last = tile_last;
	st.last = tile_last;
	if (tile_last < pop - 1u) {
		const flmip_level_info& lt = im.levels[tile_last];
		if (lt.slice_size <= FLMIP_PTILE_PATCH_BYTES) st.last = pop - 1u; // the last tile of a layer finishes the chain in the same launch
	}
	return st;
}

int launch_ptile(flmip_image_s& im, device_state* ds, uint32_t src, const ptile_step& st, CUstream stream) {
	const flmip_level_info& ls = im.levels[src];
	if (!im.ptile_map_valid[src]) {
		const cuuint64_t row_bytes = (cuuint64_t)ls.dim[0] * im.bpp;
		cuuint64_t gdim[3] = { row_bytes / 4u, ls.dim[1], im.layers };
		cuuint64_t gstride[2] = { row_bytes, row_bytes * ls.dim[1] };
		cuuint32_t box[3] = { im.tiling.tile_bytes_x / 4u, im.tiling.ty, 1u };
		cuuint32_t estride[3] = { 1, 1, 1 };
		// what lies outside the level is filled with zeros (partial tiles at the right / bottom border)
		CU_TRY(cu.p_cuTensorMapEncodeTiled(&im.ptile_map[src], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, reinterpret_cast<void*>(im.mem + ls.offset), gdim, gstride, box,
										   estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
										   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE),
			   "cuTensorMapEncodeTiled(tile kernel)");
		im.ptile_map_valid[src] = true;
	}
	flmip_ptile_params P;
	memset(&P, 0, sizeof(P));
	P.base = im.mem;
	for (uint32_t l = 0; l < im.level_count; ++l) {
		P.level_off[l] = im.levels[l].offset;
		P.dim[l][0] = im.levels[l].dim[0];
		P.dim[l][1] = im.levels[l].dim[1];
		P.wtab_off[l][0] = im.wtab_off[l][0];
		P.wtab_off[l][1] = im.wtab_off[l][1];
	}
	P.wtab = im.wtab;
	P.counters = im.pcounters;
	P.sched = im.pcounters + (uint64_t)im.layers * sizeof(uint32_t);
	P.src_level = src;
	P.tile_last = st.tile_last;
	P.last_level = st.last;
	P.tiles[0] = (ls.dim[0] + im.tiling.tx - 1u) / im.tiling.tx;
	P.tiles[1] = (ls.dim[1] + im.tiling.ty - 1u) / im.tiling.ty;
	P.layers = im.layers;
	P.total_tiles = P.tiles[0] * P.tiles[1] * im.layers;
	P.stages = 2;
	P.no_double = im.no_double;
	P.late_wait = take_launch_mode();
	// units of 4 tiles (one publish per unit) once every resident CTA has many tiles to work through; below that the pool keeps up
	// with one publish per tile, and units would only coarsen what the scheduler can balance
	const uint64_t resident_ctas = 2ull * ds->info.units;
	static const uint32_t env_unit_tiles = env_u32("FLMIP_PTILE_UNIT_MIN_TILES_PER_CTA", 8);
	P.unit_shift = (st.last > st.tile_last && P.total_tiles >= env_unit_tiles * resident_ctas) ? 2u : 0u;
	// full-width vector stores need every row of the level to start on a 16- / 8-byte boundary
	const flmip_level_info& l1 = im.levels[src + 1u];
	const uint64_t pitch1 = (uint64_t)l1.dim[0] * im.bpp;
	P.vec1 = (((im.mem + l1.offset) | l1.slice_size | pitch1) & 15u) == 0u;
	if (src + 2u < im.level_count) {
		const flmip_level_info& l2 = im.levels[src + 2u];
		const uint64_t pitch2 = (uint64_t)l2.dim[0] * im.bpp, mask = im.bpp == 16 ? 15u : 7u;
		P.vec2 = (((im.mem + l2.offset) | l2.slice_size | pitch2) & mask) == 0u;
	}
	CUfunction fn = nullptr;
	const int rc = get_function(ds, im.ptile_name, im.ptile_smem, &fn);
	if (rc != FLMIP_OK) return rc;
	const uint64_t units = ((uint64_t)P.total_tiles + (1u << P.unit_shift) - 1u) >> P.unit_shift;
	const uint64_t grid = units < resident_ctas ? units : resident_ctas;
	void* args[] = { &im.ptile_map[src], &P };
	return launch(fn, grid, FLMIP_BLOCK_THREADS, im.ptile_smem, stream, args, true);
}

// number of launches that produce levels (src_level, populated): persistent TMA tile kernel where a level qualifies as its source
uint32_t tile_launch_count(const flmip_image_s& im, uint32_t src_level, uint32_t* tma_launches) {
	const uint32_t pop = populated_levels(im);
	uint32_t n = 0, tma = 0;
	for (uint32_t s = src_level; s + 1 < pop;) {
		const ptile_step pst = plan_ptile_step(im, s, pop);
		if (pst.ok) { s = pst.last; ++tma; }
		else s += tile_levels_from(im, s, pop);
		++n;
	}
	if (tma_launches) *tma_launches = tma;
	return n;
}

int check_image(flmip_image img) {
	if (!img) return fail(FLMIP_ERR_INVALID, "null image handle");
	return FLMIP_OK;
}

} // namespace

// ------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------
extern "C" {

int flmip_init(void) { return ensure_init(); }

int flmip_device_count(void) {
	if (ensure_init() != FLMIP_OK) return 0;
	return (int)devices.size();
}

int flmip_get_device_info(int device, flmip_device_info* out) {
	if (!out) return fail(FLMIP_ERR_INVALID, "null output");
	const int rc = ensure_init();
	if (rc != FLMIP_OK) return rc;
	if (device < 0 || (size_t)device >= devices.size()) return fail(FLMIP_ERR_INVALID, "invalid device index %d", device);
	*out = devices[(size_t)device]->info;
	return FLMIP_OK;
}

const char* flmip_last_error_string(void) { return tl_error.c_str(); }
uint64_t flmip_launch_count(void) { return launch_counter.load(std::memory_order_relaxed); }

int flmip_stream_create(int device, flmip_stream* out) {
	if (!out) return fail(FLMIP_ERR_INVALID, "null output");
	WITH_DEVICE(device)
	CUstream s = nullptr;
	CU_TRY(cu.p_cuStreamCreate(&s, CU_STREAM_NON_BLOCKING), "cuStreamCreate"); // cuda_context.cpp:418-437
	*out = s;
	return FLMIP_OK;
}
int flmip_stream_destroy(int device, flmip_stream stream) {
	WITH_DEVICE(device)
	{
		// images whose last chain was enqueued on this stream: leave an event behind that the next chain on them waits for
		std::lock_guard<std::mutex> lock(ds->images_mtx);
		for (flmip_image_s* im : ds->images) {
			std::lock_guard<std::mutex> g(im->gen_mtx);
			if (!im->has_last || im->last_stream != (CUstream)stream) continue;
			if (!im->handover && cu.p_cuEventCreate(&im->handover, CU_EVENT_DISABLE_TIMING) != CUDA_SUCCESS) continue;
			if (cu.p_cuEventRecord(im->handover, (CUstream)stream) == CUDA_SUCCESS) im->pending_handover = true;
			im->has_last = false;
			im->last_stream = nullptr;
		}
	}
	flmip_overlap_bookkeeping(reinterpret_cast<uint64_t>(stream), 0, 4, 0);
	CU_TRY(cu.p_cuStreamDestroy((CUstream)stream), "cuStreamDestroy");
	return FLMIP_OK;
}
int flmip_stream_sync(int device, flmip_stream stream) {
	WITH_DEVICE(device)
	CU_TRY(cu.p_cuStreamSynchronize((CUstream)stream), "cuStreamSynchronize");
	return FLMIP_OK;
}
int flmip_stream_set_chain_overlap(int device, flmip_stream stream, int enable) {
	WITH_DEVICE(device)
	(void)ds;
	return flmip_overlap_bookkeeping(reinterpret_cast<uint64_t>(stream), 0, 0, enable != 0);
}
// The bookkeeping above is plain host logic: this hook lets the CPU tests drive it without a GPU.  op 0: opt `stream` in / out (arg = 0 / 1);
// op 1: a chain of `arg` kernels on `image` whose first kernel is a PDL kernel -- returns 1 if that kernel would start late; op 2: a chain
// whose first kernel is the literal kernel (never late); op 3: anything else enqueued on the stream; op 4: forget the stream.
int flmip_overlap_bookkeeping(uint64_t stream, uint64_t image, uint32_t op, uint32_t arg) {
	const CUstream s = reinterpret_cast<CUstream>(stream);
	const void* img = reinterpret_cast<const void*>(image);
	switch (op) {
		case 0: {
			const std::shared_ptr<stream_run> r = run_of(s, true);
			std::lock_guard<std::mutex> lock(r->mtx);
			if ((arg != 0) != r->enabled) {
				r->enabled = arg != 0;
				if (r->enabled) overlap_streams.fetch_add(1, std::memory_order_relaxed);
				else overlap_streams.fetch_sub(1, std::memory_order_relaxed);
			}
			r->images.clear();
			return 0;
		}
		case 1:
		case 2: {
			const std::shared_ptr<stream_run> r = run_of(s);
			if (!r) return 0;
			std::lock_guard<std::mutex> lock(r->mtx);
			const bool late = op == 1 && run_allows_late_head(*r, img);
			run_note_chain(*r, img, arg, late);
			return late ? 1 : 0;
		}
		case 3: run_close(s); return 0;
		case 4: {
			std::lock_guard<std::mutex> lock(runs_mtx);
			auto it = runs.find(s);
			if (it != runs.end()) {
				std::lock_guard<std::mutex> rl(it->second->mtx);
				if (it->second->enabled) overlap_streams.fetch_sub(1, std::memory_order_relaxed);
				it->second->enabled = false;
				it->second->images.clear();
				runs.erase(it);
			}
			return 0;
		}
		default: return fail(FLMIP_ERR_INVALID, "unknown bookkeeping op %u", op);
	}
}
int flmip_stream_fence(int device, flmip_stream stream) {
	WITH_DEVICE(device)
	(void)ds;
	run_close((CUstream)stream);
	return FLMIP_OK;
}
int flmip_event_create(int device, flmip_event* out) {
	if (!out) return fail(FLMIP_ERR_INVALID, "null output");
	WITH_DEVICE(device)
	CUevent e = nullptr;
	CU_TRY(cu.p_cuEventCreate(&e, CU_EVENT_DEFAULT), "cuEventCreate");
	*out = e;
	return FLMIP_OK;
}
int flmip_event_record(int device, flmip_event ev, flmip_stream stream) {
	run_close((CUstream)stream); // chain overlap: anything but a chain kernel ends the stream's open run
	WITH_DEVICE(device)
	CU_TRY(cu.p_cuEventRecord((CUevent)ev, (CUstream)stream), "cuEventRecord");
	return FLMIP_OK;
}
int flmip_event_sync(int device, flmip_event ev) {
	WITH_DEVICE(device)
	CU_TRY(cu.p_cuEventSynchronize((CUevent)ev), "cuEventSynchronize");
	return FLMIP_OK;
}
int flmip_event_elapsed_ms(int device, flmip_event start, flmip_event stop, float* ms) {
	if (!ms) return fail(FLMIP_ERR_INVALID, "null output");
	WITH_DEVICE(device)
	CU_TRY(cu.p_cuEventElapsedTime(ms, (CUevent)start, (CUevent)stop), "cuEventElapsedTime");
	return FLMIP_OK;
}
int flmip_event_destroy(int device, flmip_event ev) {
	WITH_DEVICE(device)
	CU_TRY(cu.p_cuEventDestroy((CUevent)ev), "cuEventDestroy");
	return FLMIP_OK;
}
int flmip_host_alloc(int device, size_t size, void** out) {
	if (!out) return fail(FLMIP_ERR_INVALID, "null output");
	WITH_DEVICE(device)
	CU_TRY(cu.p_cuMemHostAlloc(out, size, CU_MEMHOSTALLOC_PORTABLE), "cuMemHostAlloc");
	return FLMIP_OK;
}
int flmip_host_alloc_ex(int device, size_t size, uint32_t flags, void** out) {
	if (!out) return fail(FLMIP_ERR_INVALID, "null output");
	WITH_DEVICE(device)
	unsigned f = CU_MEMHOSTALLOC_PORTABLE;
	if (flags & FLMIP_HOST_WRITE_COMBINED) f |= CU_MEMHOSTALLOC_WRITECOMBINED;
	CU_TRY(cu.p_cuMemHostAlloc(out, size, f), "cuMemHostAlloc");
	return FLMIP_OK;
}
int flmip_host_free(int device, void* ptr) {
	WITH_DEVICE(device)
	CU_TRY(cu.p_cuMemFreeHost(ptr), "cuMemFreeHost");
	return FLMIP_OK;
}

} // extern "C"

namespace {
int create_image_impl(int device, uint64_t image_type, const uint32_t image_dim[4], uint32_t mip_level_limit, uint32_t flags, uint64_t external_ptr,
					  uint64_t external_size, bool external, flmip_image* out) {
	if (!out || !image_dim) return fail(FLMIP_ERR_INVALID, "null argument");
	*out = nullptr;
	auto im = new flmip_image_s;
	im->device = device;
	im->type = image_type;
	memcpy(im->dim, image_dim, sizeof(im->dim));
	im->no_double = (flags & FLMIP_IMAGE_NO_DOUBLE) ? 1u : 0u;
	int rc = decode_type(*im);
	if (rc != FLMIP_OK) { delete im; return rc; }
	if (im->dim[0] == 0 || (im->dc >= 2 && im->dim[1] == 0) || (im->dc >= 3 && im->dim[2] == 0)) {
		delete im;
		return fail(FLMIP_ERR_INVALID, "image dimensions must be non-zero");
	}
	// image_types.hpp:716-726
	const bool is_array = (image_type & T_FLAG_ARRAY) != 0, is_cube = (image_type & T_FLAG_CUBE) != 0;
	im->layers = !is_array ? 1u : (im->dc == 1 ? im->dim[1] : (im->dc == 2 ? im->dim[2] : im->dim[3]));
	if (is_cube) {
		im->layers *= 6u;
		if (im->dim[0] != im->dim[1]) { delete im; return fail(FLMIP_ERR_INVALID, "cube map side width and height must be equal (cuda_image.cpp:187-191)"); }
	}
	if (im->layers == 0) { delete im; return fail(FLMIP_ERR_INVALID, "array image without layers"); }
	// device_image.hpp:483-485
	im->level_count = mip_level_count_for(im->dim, image_type, im->dc);
	if (mip_level_limit > 0 && mip_level_limit < im->level_count) im->level_count = mip_level_limit;
	if (im->level_count > FLMIP_MAX_LEVELS) { delete im; return fail(FLMIP_ERR_UNSUPPORTED, "more than %u mip levels", FLMIP_MAX_LEVELS); }
	// level table: dim >> level without max(1) (host_image.cpp:75-88, image_types.hpp:751-766), 64-bit offsets
	uint64_t off = 0;
	for (uint32_t l = 0; l < im->level_count; ++l) {
		flmip_level_info& li = im->levels[l];
		li.dim[0] = im->dim[0] >> l;
		li.dim[1] = im->dc >= 2 ? im->dim[1] >> l : 0;
		li.dim[2] = im->dc >= 3 ? im->dim[2] >> l : 0;
		uint64_t texels = li.dim[0];
		if (im->dc >= 2) texels *= li.dim[1];
		if (im->dc >= 3) texels *= li.dim[2];
		li.slice_size = texels * im->bpp;
		li.size = li.slice_size * im->layers;
		li.offset = off;
		off += li.size;
	}
	im->total_size = off;

	device_state* ds = nullptr;
	rc = get_device(device, &ds);
	if (rc != FLMIP_OK) { delete im; return rc; }
	ctx_guard guard;
	rc = guard.push(ds);
	if (rc != FLMIP_OK) { delete im; return rc; }
	if (external) {
		// caller-owned linear memory in floor's host layout (another context's buffer, an imported Vulkan / OpenCL allocation ...)
		if (external_ptr == 0 || external_size < im->total_size) {
			const unsigned long long need = im->total_size;
			delete im;
			return fail(FLMIP_ERR_INVALID, "external image memory: null or smaller than the %llu bytes of the level-major image", need);
		}
		if (external_ptr & 15u) { delete im; return fail(FLMIP_ERR_INVALID, "external image memory must be 16-byte aligned (TMA, vector stores)"); }
		im->mem = (CUdeviceptr)external_ptr;
		im->external_mem = true;
	} else {
		CUresult r = cu.p_cuMemAlloc(&im->mem, im->total_size);
		if (r != CUDA_SUCCESS) { delete im; return cu_fail(r, "cuMemAlloc(image)"); }
	}
	rc = plan_fast(*im, ds, (flags & FLMIP_IMAGE_FORCE_TILED) ? (flags | FLMIP_IMAGE_FORCE_GENERIC) : flags);
	if (rc != FLMIP_OK && rc != FLMIP_ERR_OUT_OF_MEMORY) {
		// the single-pass plan is an optimisation: when one of its optional steps fails (tensor-map encoding, the shared-memory
		// opt-in, a tuning override that leaves no room for a tile) the tile / literal kernels still serve the image
		if (im->counters) cu.p_cuMemFree(im->counters);
		im->counters = 0;
		im->fast = false;
		rc = FLMIP_OK;
	}
	const bool literal_only = im->elem_kind >= FLMIP_EK_COUNT; // 2 / 4-bit formats (3-channel images: the LDG tile kernel, never TMA -- 3, 6 and 12-byte texels)
	if (rc == FLMIP_OK && im->dc >= 2 && !literal_only && ((flags & FLMIP_IMAGE_FORCE_TILED) || !(flags & FLMIP_IMAGE_FORCE_GENERIC))) {
		char name[64];
		snprintf(name, sizeof(name), "flmip_tile%ud_k%u_c%u", im->dc, im->elem_kind, im->channels);
		im->tile_name = name;
		im->tiled = true;
		if (im->level_count > 1) {
			CUfunction fn = nullptr;
			rc = get_function(ds, im->tile_name, 0, &fn); // resolve now: fail at creation, not at first use
		}
	}
	// persistent TMA tile kernel for what the single-pass kernel does not take (2D, incl. arrays / cubes): optional, like the
	// single-pass plan -- when any of its steps fails the LDG tile kernel serves the image
	static const bool env_no_tma_tiles = env_u32("FLMIP_NO_TMA_TILES", 0) != 0; // A/B tuning runs
	if (rc == FLMIP_OK && im->tiled && im->dc == 2 && im->channels != 3 && !(flags & FLMIP_IMAGE_NO_TMA_TILES) && !env_no_tma_tiles && populated_levels(*im) > 1u &&
		(!im->fast || im->fast_level_count < populated_levels(*im))) {
		im->tiling = flmip_tiling_lookup(im->bpp, 2);
		im->sm_count = ds->info.units;
		im->ptile_flags = flags;
		char name[64];
		snprintf(name, sizeof(name), "flmip_ptile2d_k%u_c%u", im->elem_kind, im->channels);
		im->ptile_name = name;
		im->ptile_smem = 2u * im->tiling.tile_bytes + FLMIP_FINISHER_WARPS * (im->tiling.cascade_bytes + im->tiling.cascade_bytes / 4u) +
						 FLMIP_PTILE_PATCH_BYTES + FLMIP_PTILE_PATCH_BYTES / 4u;
		CUfunction fn = nullptr;
		int prc = get_function(ds, im->ptile_name, im->ptile_smem, &fn);
		if (prc == FLMIP_OK) prc = build_sampler_table(*im, ds);
		if (prc == FLMIP_OK) {
#ifdef FLMIP_TIMELINE
			const size_t n_pc = (size_t)im->layers + 2u + 4u + 2u * 8u * 1024u; // + 4 x u64 per CTA behind the scheduler words (tuning builds)
#else
			const size_t n_pc = (size_t)im->layers + 2u;
#endif
			const CUresult r = cu.p_cuMemAlloc(&im->pcounters, n_pc * sizeof(uint32_t));
			if (r != CUDA_SUCCESS) prc = cu_fail(r, "cuMemAlloc(tile counters)");
		}
		if (prc == FLMIP_OK) {
#ifdef FLMIP_TIMELINE
			const size_t n_pc = (size_t)im->layers + 2u + 4u + 2u * 8u * 1024u;
#else
			const size_t n_pc = (size_t)im->layers + 2u;
#endif
			std::lock_guard<std::mutex> lock(ds->mtx);
			CUresult r = CUDA_SUCCESS;
			if (!ds->util_stream) r = cu.p_cuStreamCreate(&ds->util_stream, CU_STREAM_NON_BLOCKING);
			if (r == CUDA_SUCCESS) r = cu.p_cuMemsetD32Async(im->pcounters, 0, n_pc, ds->util_stream);
			if (r == CUDA_SUCCESS) r = cu.p_cuStreamSynchronize(ds->util_stream);
			if (r != CUDA_SUCCESS) prc = cu_fail(r, "zeroing the tile counters");
		}
		if (prc == FLMIP_OK) {
			im->ptile = true;
		} else if (prc == FLMIP_ERR_OUT_OF_MEMORY) {
			rc = prc;
		} else {
			if (im->wtab) cu.p_cuMemFree(im->wtab);
			if (im->pcounters) cu.p_cuMemFree(im->pcounters);
			im->wtab = im->pcounters = 0;
		}
	}
	if (rc != FLMIP_OK) {
		if (im->counters) cu.p_cuMemFree(im->counters);
		if (im->wtab) cu.p_cuMemFree(im->wtab);
		if (im->pcounters) cu.p_cuMemFree(im->pcounters);
		if (!im->external_mem) cu.p_cuMemFree(im->mem);
		delete im;
		return rc;
	}
	{
		std::lock_guard<std::mutex> lock(ds->images_mtx);
		ds->images.insert(im);
	}
	*out = im;
	return FLMIP_OK;
}

// Orders a chain about to be enqueued on `stream` after the previous chain of the image (see flmip_image_s::gen_mtx).  Called with
// gen_mtx held and the device's context current.
int order_after_previous_chain(flmip_image_s& im, CUstream stream) {
	if (tl_recorder) return FLMIP_OK; // recording a batch graph: flmip_batch_generate orders the graph launch
	if (im.pending_handover) {
		run_close(stream);
		CU_TRY(cu.p_cuStreamWaitEvent(stream, im.handover, 0), "cuStreamWaitEvent(chain hand-over)");
		im.pending_handover = false;
	} else if (im.has_last && im.last_stream != stream) {
		run_close(stream);
		run_close(im.last_stream);
		if (!im.handover) CU_TRY(cu.p_cuEventCreate(&im.handover, CU_EVENT_DISABLE_TIMING), "cuEventCreate(chain hand-over)");
		CU_TRY(cu.p_cuEventRecord(im.handover, im.last_stream), "cuEventRecord(chain hand-over)");
		CU_TRY(cu.p_cuStreamWaitEvent(stream, im.handover, 0), "cuStreamWaitEvent(chain hand-over)");
	}
	im.last_stream = stream;
	im.has_last = true;
	return FLMIP_OK;
}
} // namespace

extern "C" {

int flmip_image_create(int device, uint64_t image_type, const uint32_t image_dim[4], uint32_t mip_level_limit, uint32_t flags, flmip_image* out) {
	return create_image_impl(device, image_type, image_dim, mip_level_limit, flags, 0, 0, false, out);
}

int flmip_image_create_external(int device, uint64_t image_type, const uint32_t image_dim[4], uint32_t mip_level_limit, uint32_t flags,
								uint64_t device_ptr, uint64_t size, flmip_image* out) {
	return create_image_impl(device, image_type, image_dim, mip_level_limit, flags, device_ptr, size, true, out);
}

int flmip_device_attach_context(int device, void* cu_context) {
	if (!cu_context) return fail(FLMIP_ERR_INVALID, "null context");
	const int rc = ensure_init();
	if (rc != FLMIP_OK) return rc;
	if (device < 0 || (size_t)device >= devices.size()) return fail(FLMIP_ERR_INVALID, "invalid device index %d (have %zu)", device, devices.size());
	device_state* ds = devices[(size_t)device];
	std::lock_guard<std::mutex> lock(ds->mtx);
	const CUcontext cur = ds->ctx.load(std::memory_order_acquire);
	if (cur == (CUcontext)cu_context) return FLMIP_OK;
	if (cur) return fail(FLMIP_ERR_INVALID, "device %d already runs on another context: attach before the first use of the device", device);
	// the context must belong to this device
	CU_TRY(cu.p_cuCtxPushCurrent((CUcontext)cu_context), "cuCtxPushCurrent(attached context)");
	CUdevice d = -1;
	const CUresult r = cu.p_cuCtxGetDevice(&d);
	CUcontext old = nullptr;
	cu.p_cuCtxPopCurrent(&old);
	if (r != CUDA_SUCCESS) return cu_fail(r, "cuCtxGetDevice");
	if (d != ds->dev) return fail(FLMIP_ERR_INVALID, "the context belongs to another device");
	ds->ctx_attached = true;
	ds->ctx.store((CUcontext)cu_context, std::memory_order_release);
	return FLMIP_OK;
}

int flmip_image_destroy(flmip_image img) {
	if (!img) return FLMIP_OK;
	WITH_DEVICE(img->device)
	{
		std::lock_guard<std::mutex> lock(ds->images_mtx);
		ds->images.erase(img);
	}
	if (img->handover) cu.p_cuEventDestroy(img->handover);
	if (img->wtab) cu.p_cuMemFree(img->wtab);
	if (img->pcounters) cu.p_cuMemFree(img->pcounters);
	if (img->counters) cu.p_cuMemFree(img->counters);
	if (img->mem && !img->external_mem) cu.p_cuMemFree(img->mem);
	delete img;
	return FLMIP_OK;
}

int flmip_image_mip_level_count(flmip_image img, uint32_t* out) {
	if (check_image(img) || !out) return fail(FLMIP_ERR_INVALID, "null argument");
	*out = img->level_count;
	return FLMIP_OK;
}
int flmip_image_layer_count(flmip_image img, uint32_t* out) {
	if (check_image(img) || !out) return fail(FLMIP_ERR_INVALID, "null argument");
	*out = img->layers;
	return FLMIP_OK;
}
int flmip_image_data_size(flmip_image img, uint64_t* out) {
	if (check_image(img) || !out) return fail(FLMIP_ERR_INVALID, "null argument");
	*out = img->total_size;
	return FLMIP_OK;
}
int flmip_image_get_level_info(flmip_image img, uint32_t level, flmip_level_info* out) {
	if (check_image(img) || !out) return fail(FLMIP_ERR_INVALID, "null argument");
	if (level >= img->level_count) return fail(FLMIP_ERR_INVALID, "mip level %u out of range (%u levels)", level, img->level_count);
	*out = img->levels[level];
	return FLMIP_OK;
}
int flmip_image_device_ptr(flmip_image img, uint64_t* out) {
	if (check_image(img) || !out) return fail(FLMIP_ERR_INVALID, "null argument");
	*out = (uint64_t)img->mem;
	return FLMIP_OK;
}
int flmip_image_plan(flmip_image img, uint32_t* uses_single_pass, uint32_t* fast_levels, uint32_t* launches) {
	if (check_image(img)) return FLMIP_ERR_INVALID;
	uint32_t n = 0;
	const uint32_t first_generic = img->fast ? img->fast_level_count : 1u;
	if (img->fast) ++n;
	if (img->tiled) {
		n += tile_launch_count(*img, first_generic - 1u, nullptr);
	} else {
		for (uint32_t l = first_generic; l < img->level_count; ++l) {
			const flmip_level_info& li = img->levels[l];
			if (li.size != 0) ++n;
		}
	}
	if (uses_single_pass) *uses_single_pass = img->fast ? 1u : 0u;
	if (fast_levels) *fast_levels = img->fast ? img->fast_level_count : 0u;
	if (launches) *launches = n;
	return FLMIP_OK;
}

int flmip_sampler_table_entry(uint32_t g, uint32_t n, uint32_t* out) {
	if (!out) return fail(FLMIP_ERR_INVALID, "null output");
	if (n < 2u || g >= (n >> 1)) return fail(FLMIP_ERR_INVALID, "destination texel %u outside a level of %u >> 1 texels", g, n);
	if (!sampler_entry(g, n, out)) return fail(FLMIP_ERR_UNSUPPORTED, "irregular sampler fetch (texel %u of a %u-texel source level)", g, n);
	return FLMIP_OK;
}

int flmip_image_plan_tma_tile_launches(flmip_image img, uint32_t* out) {
	if (check_image(img) || !out) return fail(FLMIP_ERR_INVALID, "null argument");
	*out = 0;
	if (img->tiled) tile_launch_count(*img, (img->fast ? img->fast_level_count : 1u) - 1u, out);
	return FLMIP_OK;
}

int flmip_image_upload(flmip_image img, const void* src, size_t src_size, uint32_t level_first, uint32_t level_last, flmip_stream stream) {
	run_close((CUstream)stream); // chain overlap: anything but a chain kernel ends the stream's open run
	if (check_image(img)) return FLMIP_ERR_INVALID;
	if (!src) return fail(FLMIP_ERR_INVALID, "null source");
	if (level_first > level_last || level_last >= img->level_count) return fail(FLMIP_ERR_INVALID, "invalid mip level range [%u, %u]", level_first, level_last);
	const uint64_t begin = img->levels[level_first].offset, end = img->levels[level_last].offset + img->levels[level_last].size;
	if (src_size < end - begin) return fail(FLMIP_ERR_INVALID, "image upload: insufficient host data (%zu < %llu)", src_size, (unsigned long long)(end - begin));
	if (end == begin) return FLMIP_OK;
	WITH_DEVICE(img->device)
	CU_TRY(cu.p_cuMemcpyHtoDAsync(img->mem + begin, src, end - begin, (CUstream)stream), "cuMemcpyHtoDAsync");
	return FLMIP_OK;
}

int flmip_image_download(flmip_image img, void* dst, size_t dst_size, uint32_t level_first, uint32_t level_last, flmip_stream stream) {
	run_close((CUstream)stream); // chain overlap: anything but a chain kernel ends the stream's open run
	if (check_image(img)) return FLMIP_ERR_INVALID;
	if (!dst) return fail(FLMIP_ERR_INVALID, "null destination");
	if (level_first > level_last || level_last >= img->level_count) return fail(FLMIP_ERR_INVALID, "invalid mip level range [%u, %u]", level_first, level_last);
	const uint64_t begin = img->levels[level_first].offset, end = img->levels[level_last].offset + img->levels[level_last].size;
	if (dst_size < end - begin) return fail(FLMIP_ERR_INVALID, "image download: insufficient host buffer (%zu < %llu)", dst_size, (unsigned long long)(end - begin));
	if (end == begin) return FLMIP_OK;
	WITH_DEVICE(img->device)
	CU_TRY(cu.p_cuMemcpyDtoHAsync(dst, img->mem + begin, end - begin, (CUstream)stream), "cuMemcpyDtoHAsync");
	return FLMIP_OK;
}

int flmip_image_download_layers(flmip_image img, void* dst, size_t dst_size, uint32_t level_first, uint32_t level_last, uint32_t layer_first,
								uint32_t layer_count, flmip_stream stream) {
	run_close((CUstream)stream); // chain overlap: anything but a chain kernel ends the stream's open run
	if (check_image(img)) return FLMIP_ERR_INVALID;
	if (!dst) return fail(FLMIP_ERR_INVALID, "null destination");
	if (level_first > level_last || level_last >= img->level_count) return fail(FLMIP_ERR_INVALID, "invalid mip level range [%u, %u]", level_first, level_last);
	if (layer_count == 0 || (uint64_t)layer_first + layer_count > img->layers)
		return fail(FLMIP_ERR_INVALID, "invalid layer range [%u, +%u) of %u layers", layer_first, layer_count, img->layers);
	uint64_t need = 0;
	for (uint32_t l = level_first; l <= level_last; ++l) need += img->levels[l].slice_size * layer_count;
	if (dst_size < need) return fail(FLMIP_ERR_INVALID, "image download: insufficient host buffer (%zu < %llu)", dst_size, (unsigned long long)need);
	WITH_DEVICE(img->device)
	uint8_t* cur = static_cast<uint8_t*>(dst);
	for (uint32_t l = level_first; l <= level_last; ++l) {
		const flmip_level_info& li = img->levels[l];
		const uint64_t bytes = li.slice_size * layer_count;
		if (bytes == 0) continue;
		CU_TRY(cu.p_cuMemcpyDtoHAsync(cur, img->mem + li.offset + (uint64_t)layer_first * li.slice_size, bytes, (CUstream)stream), "cuMemcpyDtoHAsync(layers)");
		cur += bytes;
	}
	return FLMIP_OK;
}

int flmip_image_write(flmip_image img, const void* src, size_t src_size, const uint32_t offset[3], const uint32_t extent[3],
					  const uint32_t mip_level_range[2], const uint32_t layer_range[2], flmip_stream stream) {
	run_close((CUstream)stream); // chain overlap: anything but a chain kernel ends the stream's open run
	if (check_image(img)) return FLMIP_ERR_INVALID;
	if (!src || !offset || !extent || !mip_level_range || !layer_range) return fail(FLMIP_ERR_INVALID, "null argument");
	// write_check (device_image.cpp:503-547)
	if (mip_level_range[0] > mip_level_range[1] || mip_level_range[1] >= img->level_count) return fail(FLMIP_ERR_INVALID, "image write: invalid mip level range");
	if (layer_range[0] > layer_range[1] || layer_range[1] >= img->layers) return fail(FLMIP_ERR_INVALID, "image write: invalid layer range");
	for (uint32_t d = 0; d < img->dc; ++d) {
		if (extent[d] == 0 || (uint64_t)offset[d] + extent[d] > img->dim[d]) return fail(FLMIP_ERR_INVALID, "image write: offset + extent out of bounds in dim %u", d);
	}
	if (src_size == 0) return fail(FLMIP_ERR_INVALID, "image write: trying to write 0 bytes");
	// The per-level region (offset >> level, max(extent >> level, 1)) can leave a level whose dim is odd (dim 5, offset 4, extent 1:
	// level 1 has 2 texels, the region starts at 2).  The reference's cuMemcpy3D into the level's CUarray fails there and write()
	// returns false; on linear memory nothing would stop the copy from running into the next level, so every level of the range
	// is checked before the first copy is issued (nothing is written when the call fails).
	const uint32_t n_layers = layer_range[1] - layer_range[0] + 1u;
	uint64_t need = 0;
	for (uint32_t level = mip_level_range[0]; level <= mip_level_range[1]; ++level) {
		const flmip_level_info& li = img->levels[level];
		if (li.size == 0) continue;
		uint64_t texels = 1;
		for (uint32_t d = 0; d < img->dc; ++d) {
			const uint32_t o = offset[d] >> level, e = (extent[d] >> level) ? (extent[d] >> level) : 1u;
			if ((uint64_t)o + e > li.dim[d])
				return fail(FLMIP_ERR_INVALID, "image write: region [%u, +%u) leaves mip-level %u (%u texels) in dim %u", o, e, level, li.dim[d], d);
			texels *= e;
		}
		need += texels * img->bpp * n_layers;
	}
	if (src_size < need) return fail(FLMIP_ERR_INVALID, "image write: insufficient host data (%zu < %llu)", src_size, (unsigned long long)need);
	WITH_DEVICE(img->device)
	const uint8_t* cur = static_cast<const uint8_t*>(src);
	size_t left = src_size;
	for (uint32_t level = mip_level_range[0]; level <= mip_level_range[1]; ++level) {
		const flmip_level_info& li = img->levels[level];
		// derive the mip extent from the user-specified extent (cuda_image.cpp:609-611)
		uint32_t e[3] = { 1, 1, 1 }, o[3] = { 0, 0, 0 };
		for (uint32_t d = 0; d < img->dc; ++d) {
			e[d] = extent[d] >> level;
			if (e[d] == 0) e[d] = 1;
			o[d] = offset[d] >> level;
		}
		if (li.size == 0) continue;
		// the source holds this level of an image of size `extent` with n_layers layers, tightly packed
		const uint64_t src_row = (uint64_t)e[0] * img->bpp, src_slice = src_row * e[1] * e[2];
		const uint64_t bytes = src_slice * n_layers;
		if (left < bytes) return fail(FLMIP_ERR_INVALID, "image write: insufficient host data at mip-level %u", level);
		for (uint32_t layer = 0; layer < n_layers; ++layer) {
			CUDA_MEMCPY3D c;
			memset(&c, 0, sizeof(c));
			c.srcMemoryType = CU_MEMORYTYPE_HOST;
			c.srcHost = cur + layer * src_slice;
			c.srcPitch = src_row;
			c.srcHeight = e[1];
			c.dstMemoryType = CU_MEMORYTYPE_DEVICE;
			c.dstDevice = img->mem + li.offset + (uint64_t)(layer_range[0] + layer) * li.slice_size;
			c.dstPitch = (uint64_t)li.dim[0] * img->bpp;
			c.dstHeight = img->dc >= 2 ? li.dim[1] : 1;
			c.dstXInBytes = (uint64_t)o[0] * img->bpp;
			c.dstY = o[1];
			c.dstZ = o[2];
			c.WidthInBytes = src_row;
			c.Height = e[1];
			c.Depth = e[2];
			CU_TRY(cu.p_cuMemcpy3DAsync(&c, (CUstream)stream), "cuMemcpy3DAsync(image write)");
		}
		cur += bytes;
		left -= bytes;
	}
	return FLMIP_OK;
}

int flmip_image_zero(flmip_image img, flmip_stream stream) {
	run_close((CUstream)stream); // chain overlap: anything but a chain kernel ends the stream's open run
	if (check_image(img)) return FLMIP_ERR_INVALID;
	WITH_DEVICE(img->device)
	CU_TRY(cu.p_cuMemsetD8Async(img->mem, 0, img->total_size, (CUstream)stream), "cuMemsetD8Async");
	return FLMIP_OK;
}

// -- blit / clone support: device_image::blit (device_image.hpp:96-101; CUDA inherits the `return false` stub, so
//    clone(copy_contents = true) copies nothing there) -- on linear images it is one device-to-device copy.
int flmip_image_blit(flmip_image dst, flmip_image src, flmip_stream stream) {
	run_close((CUstream)stream); // chain overlap: anything but a chain kernel ends the stream's open run
	if (check_image(dst) || check_image(src)) return FLMIP_ERR_INVALID;
	// blit_check (device_image.cpp:470-501): identical dim, layer count, size and format; no compressed formats
	if (memcmp(dst->dim, src->dim, sizeof(dst->dim)) != 0) return fail(FLMIP_ERR_INVALID, "blit: dim mismatch");
	if (dst->layers != src->layers) return fail(FLMIP_ERR_INVALID, "blit: layer count mismatch: src %u != dst %u", src->layers, dst->layers);
	if ((dst->type & T_FORMAT_MASK) != (src->type & T_FORMAT_MASK) || dst->channels != src->channels)
		return fail(FLMIP_ERR_INVALID, "blit: format mismatch");
	if (dst->device != src->device) return fail(FLMIP_ERR_INVALID, "blit: images live on different devices");
	// levels both images have (a clone may carry a different mip_level_limit)
	const uint32_t levels = dst->level_count < src->level_count ? dst->level_count : src->level_count;
	const uint64_t bytes = src->levels[levels - 1].offset + src->levels[levels - 1].size;
	if (bytes == 0 || dst == src) return FLMIP_OK;
	WITH_DEVICE(dst->device)
	CU_TRY(cu.p_cuMemcpyDtoDAsync(dst->mem, src->mem, bytes, (CUstream)stream), "cuMemcpyDtoDAsync(blit)");
	return FLMIP_OK;
}

int flmip_device_cu_context(int device, void** out) {
	if (!out) return fail(FLMIP_ERR_INVALID, "null output");
	device_state* ds = nullptr;
	const int rc = get_device(device, &ds);
	if (rc != FLMIP_OK) return rc;
	*out = ds->ctx.load(std::memory_order_acquire);
	return FLMIP_OK;
}

// -- interop with floor's tiled CUDA images (CUmipmappedArray + texture / surface objects,
//    src/device/cuda/cuda_image.cpp:158-539): a twin array with the reference's descriptor, and copies between the
//    linear image and the array so that kernels which still sample through texture objects see the generated chain.
int flmip_image_create_tiled_twin(flmip_image img, void** out_mipmapped_array) {
	if (check_image(img) || !out_mipmapped_array) return fail(FLMIP_ERR_INVALID, "null argument");
	*out_mipmapped_array = nullptr;
	if (img->dc < 2) return fail(FLMIP_ERR_UNSUPPORTED, "tiled interop covers 2D, 2D-array, cube, cube-array and 3D images");
	if (img->channels == 3 || img->elem_kind >= FLMIP_EK_COUNT) return fail(FLMIP_ERR_UNSUPPORTED, "a CUarray holds 1, 2 or 4 channels of 8, 16 or 32 bits (cuda_image.cpp:173-248)");
	// format LUT of cuda_image.cpp:197-207 (normalization is a property of the texture object, not of the array)
	CUarray_format fmt;
	switch (img->elem_kind) {
		case FLMIP_EK_F32: fmt = CU_AD_FORMAT_FLOAT; break;
		case FLMIP_EK_F16: fmt = CU_AD_FORMAT_HALF; break;
		case FLMIP_EK_UNORM8: case FLMIP_EK_U8: fmt = CU_AD_FORMAT_UNSIGNED_INT8; break;
		case FLMIP_EK_SNORM8: case FLMIP_EK_I8: fmt = CU_AD_FORMAT_SIGNED_INT8; break;
		case FLMIP_EK_UNORM16: case FLMIP_EK_U16: fmt = CU_AD_FORMAT_UNSIGNED_INT16; break;
		case FLMIP_EK_SNORM16: case FLMIP_EK_I16: fmt = CU_AD_FORMAT_SIGNED_INT16; break;
		case FLMIP_EK_U32: fmt = CU_AD_FORMAT_UNSIGNED_INT32; break;
		default: fmt = CU_AD_FORMAT_SIGNED_INT32; break;
	}
	const bool is_array = (img->type & T_FLAG_ARRAY) != 0, is_cube = (img->type & T_FLAG_CUBE) != 0;
	CUDA_ARRAY3D_DESCRIPTOR desc;
	memset(&desc, 0, sizeof(desc));
	desc.Width = img->dim[0];
	desc.Height = img->dim[1];
	desc.Depth = img->dc == 3 ? img->dim[2] : ((is_array || is_cube) ? img->layers : 0u); // cuda_image.cpp:166-171
	desc.Format = fmt;
	desc.NumChannels = img->channels;
	desc.Flags = (is_array ? CUDA_ARRAY3D_LAYERED : 0u) | (is_cube ? CUDA_ARRAY3D_CUBEMAP : 0u) | CUDA_ARRAY3D_SURFACE_LDST;
	WITH_DEVICE(img->device)
	CUmipmappedArray arr = nullptr;
	CU_TRY(cu.p_cuMipmappedArrayCreate(&arr, &desc, img->level_count), "cuMipmappedArrayCreate");
	*out_mipmapped_array = arr;
	return FLMIP_OK;
}

int flmip_tiled_destroy(int device, void* mipmapped_array) {
	if (!mipmapped_array) return FLMIP_OK;
	WITH_DEVICE(device)
	CU_TRY(cu.p_cuMipmappedArrayDestroy((CUmipmappedArray)mipmapped_array), "cuMipmappedArrayDestroy");
	return FLMIP_OK;
}

} // extern "C"

namespace {
enum class tiled_dir { to_tiled, from_tiled, tiled_to_host };
// one cuMemcpy3DAsync per level: rows x height x (depth | layers) between the level-major linear layout and the level's CUarray
int tiled_copy(flmip_image img, void* mipmapped_array, uint32_t level_first, uint32_t level_last, tiled_dir dir, void* host, size_t host_size,
			   CUstream stream) {
	run_close(stream);
	if (check_image(img)) return FLMIP_ERR_INVALID;
	if (!mipmapped_array) return fail(FLMIP_ERR_INVALID, "null array");
	if (img->dc < 2) return fail(FLMIP_ERR_UNSUPPORTED, "tiled interop covers 2D, 2D-array, cube, cube-array and 3D images");
	if (level_first > level_last || level_last >= img->level_count) return fail(FLMIP_ERR_INVALID, "invalid mip level range [%u, %u]", level_first, level_last);
	const uint64_t begin = img->levels[level_first].offset, end = img->levels[level_last].offset + img->levels[level_last].size;
	if (dir == tiled_dir::tiled_to_host && (!host || host_size < end - begin)) return fail(FLMIP_ERR_INVALID, "tiled download: insufficient host buffer");
	WITH_DEVICE(img->device)
	for (uint32_t level = level_first; level <= level_last; ++level) {
		const flmip_level_info& li = img->levels[level];
		if (li.size == 0) continue; // empty level (zero dim quirk); the array's level has max(1, dim >> level) texels, left untouched
		CUarray arr = nullptr;
		CU_TRY(cu.p_cuMipmappedArrayGetLevel(&arr, (CUmipmappedArray)mipmapped_array, level), "cuMipmappedArrayGetLevel");
		CUDA_MEMCPY3D c;
		memset(&c, 0, sizeof(c));
		const uint64_t row = (uint64_t)li.dim[0] * img->bpp;
		c.WidthInBytes = row;
		c.Height = li.dim[1];
		c.Depth = img->dc == 3 ? li.dim[2] : img->layers;
		if (dir == tiled_dir::to_tiled) {
			c.srcMemoryType = CU_MEMORYTYPE_DEVICE; c.srcDevice = img->mem + li.offset; c.srcPitch = row; c.srcHeight = li.dim[1];
			c.dstMemoryType = CU_MEMORYTYPE_ARRAY; c.dstArray = arr;
		} else {
			c.srcMemoryType = CU_MEMORYTYPE_ARRAY; c.srcArray = arr;
			c.dstPitch = row; c.dstHeight = li.dim[1];
			if (dir == tiled_dir::from_tiled) { c.dstMemoryType = CU_MEMORYTYPE_DEVICE; c.dstDevice = img->mem + li.offset; }
			else { c.dstMemoryType = CU_MEMORYTYPE_HOST; c.dstHost = static_cast<uint8_t*>(host) + (li.offset - begin); }
		}
		CU_TRY(cu.p_cuMemcpy3DAsync(&c, stream), "cuMemcpy3DAsync(tiled interop)");
	}
	return FLMIP_OK;
}
} // namespace

extern "C" {

int flmip_image_copy_to_tiled(flmip_image img, void* mipmapped_array, uint32_t level_first, uint32_t level_last, flmip_stream stream) {
	return tiled_copy(img, mipmapped_array, level_first, level_last, tiled_dir::to_tiled, nullptr, 0, (CUstream)stream);
}
int flmip_image_copy_from_tiled(flmip_image img, void* mipmapped_array, uint32_t level_first, uint32_t level_last, flmip_stream stream) {
	return tiled_copy(img, mipmapped_array, level_first, level_last, tiled_dir::from_tiled, nullptr, 0, (CUstream)stream);
}
int flmip_tiled_download(flmip_image geometry, void* mipmapped_array, void* dst, size_t dst_size, uint32_t level_first, uint32_t level_last,
						 flmip_stream stream) {
	return tiled_copy(geometry, mipmapped_array, level_first, level_last, tiled_dir::tiled_to_host, dst, dst_size, (CUstream)stream);
}

int flmip_mip_chain_generate_from(flmip_image img, uint32_t first_level, flmip_stream stream) {
	if (check_image(img)) return FLMIP_ERR_INVALID;
	if (first_level >= img->level_count) return fail(FLMIP_ERR_INVALID, "mip level %u out of range (%u levels)", first_level, img->level_count);
	if (first_level + 1 == img->level_count) return FLMIP_OK; // nothing to generate
	WITH_DEVICE(img->device)
	std::lock_guard<std::mutex> gen_lock(img->gen_mtx);
	{
		const int rc = order_after_previous_chain(*img, (CUstream)stream);
		if (rc != FLMIP_OK) return rc;
	}
	// chain overlap (opt-in per stream): the first kernel of this chain may start without waiting for the kernel in front of it when
	// that one belongs to a chain on another image; only the PDL kernels know how (not the literal kernel, not a recorded graph)
	const bool head_is_pdl_kernel = (img->fast && first_level == 0) || img->tiled;
	const std::shared_ptr<stream_run> run = tl_recorder ? nullptr : run_of((CUstream)stream);
	std::unique_lock<std::mutex> run_lock;
	if (run) run_lock = std::unique_lock<std::mutex>(run->mtx);
	const bool late_head = run && head_is_pdl_kernel && pdl_enabled() && run_allows_late_head(*run, img);
	tl_late_head = late_head;
	tl_overlap_chain = run && run->enabled && pdl_enabled();
	tl_head_taken = false;
	tl_chain_launches = 0;
	struct chain_note { // runs before run_lock is released (declared after it)
		flmip_image img; stream_run* run; bool late;
		~chain_note() {
			tl_late_head = tl_overlap_chain = false;
			tl_head_taken = true; // launches outside a chain (fills, ...) never take a chain's mode
			if (run) run_note_chain(*run, img, tl_chain_launches, late);
		}
	} note { img, run.get(), late_head };
	uint32_t next = first_level + 1; // first level still to be produced
	if (img->fast && first_level == 0) {
		CUfunction fn = nullptr;
		int rc = get_function(ds, img->fast_name, img->fast_smem, &fn);
		if (rc != FLMIP_OK) return rc;
		flmip_fast_params P = img->fast_params;
		P.late_wait = take_launch_mode();
		void* args[] = { &img->tmap, &P };
		rc = launch(fn, img->fast_grid, FLMIP_BLOCK_THREADS, img->fast_smem, (CUstream)stream, args, true);
		if (rc != FLMIP_OK) return rc;
		next = img->fast_level_count;
	}
	// remaining levels: the multi-level tile kernel (2D / 3D, any size) ...
	if (img->tiled) return launch_tile_levels(*img, ds, next - 1u, (CUstream)stream);
	// ... or the literal general path: one launch per level, stream-ordered (the reference syncs the host after each: device_image.cpp:322)
	for (uint32_t level = next; level < img->level_count; ++level) {
		const int rc = launch_generic_level(*img, ds, level, (CUstream)stream);
		if (rc != FLMIP_OK) return rc;
	}
	return FLMIP_OK;
}

int flmip_mip_chain_generate(flmip_image img, flmip_stream stream) { return flmip_mip_chain_generate_from(img, 0, stream); }

// -- batches of independent textures: the reference loops generate_mip_map_chain over them, one blocking launch per
//    (image, layer, level).  A batch is a CUDA graph with one chain of kernel nodes per image and no edges between images, so
//    one cuGraphLaunch replaces N launches (small textures are launch-rate bound: ~6 us of host time per launch) and the GPU
//    runs the chains of different images concurrently.
struct flmip_batch_s {
	int device = 0;
	CUgraph graph = nullptr;
	CUgraphExec exec = nullptr;
	uint32_t nodes = 0, images = 0;
	std::vector<flmip_image> members;
};

int flmip_batch_create(const flmip_image* images, uint32_t count, flmip_batch* out) {
	if (!images || !out || count == 0) return fail(FLMIP_ERR_INVALID, "empty batch");
	*out = nullptr;
	for (uint32_t i = 0; i < count; ++i) {
		if (check_image(images[i])) return FLMIP_ERR_INVALID;
		if (images[i]->device != images[0]->device) return fail(FLMIP_ERR_INVALID, "batch: images live on different devices");
		for (uint32_t j = 0; j < i; ++j)
			if (images[j] == images[i]) return fail(FLMIP_ERR_INVALID, "batch: image %u appears twice (its chains would race)", i);
	}
	WITH_DEVICE(images[0]->device)
	graph_recorder rec;
	CU_TRY(cu.p_cuGraphCreate(&rec.graph, 0), "cuGraphCreate");
	int rc = FLMIP_OK;
	tl_recorder = &rec;
	// Independent chains, but not `count` of them side by side: a graph with hundreds of root nodes runs its kernels with far more
	// concurrency than the SMs have room for (512 x 1024^2 RGBA8: 16.6 us per texture, worse than one chain after the other).  The images
	// are dealt onto FLMIP_BATCH_LANES lanes; the chains of a lane run one after the other, the lanes side by side.
	static const uint32_t env_lanes = env_u32("FLMIP_BATCH_LANES", FLMIP_BATCH_LANES_DEFAULT);
	const uint32_t lanes = env_lanes ? env_lanes : count;
	std::vector<CUgraphNode> lane_last(lanes < count ? lanes : count, nullptr);
	for (uint32_t i = 0; i < count && rc == FLMIP_OK; ++i) {
		CUgraphNode& tail = lane_last[i % lane_last.size()];
		rec.has_last = tail != nullptr;
		rec.last = tail;
		rc = flmip_mip_chain_generate_from(images[i], 0, nullptr);
		if (rec.has_last) tail = rec.last;
	}
	tl_recorder = nullptr;
	CUgraphExec exec = nullptr;
	if (rc == FLMIP_OK && rec.nodes != 0) {
		const CUresult r = cu.p_cuGraphInstantiate(&exec, rec.graph, 0);
		if (r != CUDA_SUCCESS) rc = cu_fail(r, "cuGraphInstantiate");
	}
	if (rc != FLMIP_OK) {
		cu.p_cuGraphDestroy(rec.graph);
		return rc;
	}
	auto* b = new flmip_batch_s;
	b->device = images[0]->device;
	b->graph = rec.graph;
	b->exec = exec;
	b->nodes = rec.nodes;
	b->images = count;
	b->members.assign(images, images + count);
	*out = b;
	return FLMIP_OK;
}

int flmip_batch_generate(flmip_batch batch, flmip_stream stream) {
	run_close((CUstream)stream); // chain overlap: anything but a chain kernel ends the stream's open run
	if (!batch) return fail(FLMIP_ERR_INVALID, "null batch handle");
	if (!batch->exec) return FLMIP_OK; // nothing to generate (single-level images)
	WITH_DEVICE(batch->device)
	for (flmip_image im : batch->members) { // the graph runs a chain on every member: same one-chain-per-image rule
		std::lock_guard<std::mutex> g(im->gen_mtx);
		const int rc = order_after_previous_chain(*im, (CUstream)stream);
		if (rc != FLMIP_OK) return rc;
	}
	CU_TRY(cu.p_cuGraphLaunch(batch->exec, (CUstream)stream), "cuGraphLaunch");
	launch_counter.fetch_add(batch->nodes, std::memory_order_relaxed);
	return FLMIP_OK;
}

int flmip_batch_kernel_count(flmip_batch batch, uint32_t* out) {
	if (!batch || !out) return fail(FLMIP_ERR_INVALID, "null argument");
	*out = batch->nodes;
	return FLMIP_OK;
}

int flmip_batch_destroy(flmip_batch batch) {
	if (!batch) return FLMIP_OK;
	WITH_DEVICE(batch->device)
	if (batch->exec) cu.p_cuGraphExecDestroy(batch->exec);
	if (batch->graph) cu.p_cuGraphDestroy(batch->graph);
	delete batch;
	return FLMIP_OK;
}

#ifdef FLMIP_TIMELINE
// tuning builds only: the 4 time stamps (ns, %globaltimer) each CTA of the last single-pass launch left behind the scheduler words
extern "C" int flmip_debug_timeline(flmip_image img, uint64_t* out, uint32_t ctas) {
	if (check_image(img) || !out || !(img->fast || img->ptile)) return FLMIP_ERR_INVALID;
	WITH_DEVICE(img->device)
	const uint64_t sched = img->fast ? img->fast_params.sched : img->pcounters + (uint64_t)img->layers * sizeof(uint32_t);
	CU_TRY(cu.p_cuStreamSynchronize(nullptr), "cuStreamSynchronize");
	CU_TRY(cu.p_cuMemcpyDtoHAsync(out, ((sched + 15ull) & ~7ull), (size_t)ctas * 16u * sizeof(uint64_t), nullptr), "cuMemcpyDtoH(timeline)");
	CU_TRY(cu.p_cuStreamSynchronize(nullptr), "cuStreamSynchronize");
	CU_TRY(cu.p_cuMemsetD8Async(((sched + 15ull) & ~7ull), 0, (size_t)ctas * 16u * sizeof(uint64_t), nullptr), "cuMemsetD8(timeline)");
	return FLMIP_OK;
}
#endif

int flmip_image_fill_synthetic(flmip_image img, uint64_t config_id, uint64_t layer_id0, flmip_stream stream) {
	run_close((CUstream)stream); // chain overlap: anything but a chain kernel ends the stream's open run
	if (check_image(img)) return FLMIP_ERR_INVALID;
	WITH_DEVICE(img->device)
	flmip_fill_params F;
	memset(&F, 0, sizeof(F));
	F.dst = img->mem;
	F.elems_per_layer = img->levels[0].slice_size / (img->bpc >= 8u ? img->bpc / 8u : 1u); // 2 / 4-bit formats: one pattern element per storage byte
	F.config_id = config_id;
	F.layer_id0 = layer_id0;
	F.layers = img->layers;
	F.elem_kind = img->elem_kind;
	CUfunction fn = nullptr;
	const int rc = get_function(ds, "flmip_fill", 0, &fn);
	if (rc != FLMIP_OK) return rc;
	const uint64_t total = F.elems_per_layer * F.layers;
	uint64_t grid = (total + 255u) / 256u;
	const uint64_t cap = (uint64_t)ds->info.units * 32u;
	if (grid > cap) grid = cap;
	void* args[] = { &F };
	return launch(fn, grid, 256, 0, (CUstream)stream, args);
}

} // extern "C"
