// Test of the C++ drop-in (include/floor_b200/floor_b200.hpp) written the way a libfloor application uses the
// reference API: cuda_context -> device -> queue -> create_image(GENERATE_MIP_MAPS) -> write / map / unmap.
// The oracle (oracle/liboracle_minify.so, dlopen'd) is the checker.  TEST CODE: lives under tests/.
//
//   dropin_test --cpu   : enum / size arithmetic checks, the context reports "not supported" without a GPU
//   dropin_test <oracle.so> : full parity run on GPU 0
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "floor_b200/floor_b200.hpp"

using namespace fl;

typedef int (*flo_generate_fn)(void*, const uint32_t*, uint64_t, uint32_t, uint32_t, uint32_t);
typedef void (*flo_fill_fn)(void*, const uint32_t*, uint64_t, uint64_t, uint64_t, uint32_t);

static int failures = 0, current_case = -1;
#define CHECK(cond)                                                        \
	do {                                                                   \
		if (!(cond)) {                                                     \
			std::fprintf(stderr, "CHECK FAILED %s:%d (case %d): %s\n", __FILE__, __LINE__, current_case, #cond); \
			++failures;                                                    \
		}                                                                  \
	} while (0)

static void cpu_checks() {
	constexpr auto t2 = IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGBA16F | IMAGE_TYPE::FLAG_MIPMAPPED | IMAGE_TYPE::READ_WRITE;
	static_assert(image_type_bits(t2) == 0x802FC12ull);
	CHECK(image_mip_level_count({ 8192, 8192, 0, 0 }, t2) == 14);
	CHECK(image_bytes_per_pixel(t2) == 8);
	CHECK(image_data_size_from_types({ 8192, 8192, 0, 0 }, t2) == 715827880ull);
	CHECK(image_data_size_from_types({ 8192, 8192, 0, 0 }, t2, true) == 536870912ull);
	CHECK(image_mip_level_data_offset_from_types({ 8192, 8192, 0, 0 }, t2, 1) == 536870912ull);
	constexpr auto tc = IMAGE_TYPE::IMAGE_CUBE_ARRAY | IMAGE_TYPE::RGBA32F | IMAGE_TYPE::FLAG_MIPMAPPED;
	CHECK(image_layer_count({ 4096, 4096, 64, 0 }, tc) == 384);
	CHECK(image_data_size_from_types({ 4096, 4096, 64, 0 }, tc) == 137438951424ull);
	// zero-dim quirk: 64x4 has 7 levels, levels 3.. are empty (image_types.hpp:751-766)
	constexpr auto tq = IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::R8 | IMAGE_TYPE::FLAG_MIPMAPPED;
	CHECK(image_mip_level_count({ 64, 4, 0, 0 }, tq) == 7);
	CHECK(image_mip_level_data_size_from_types({ 64, 4, 0, 0 }, tq, 3) == 0);
	CHECK(device_image::infer_rw_flags(IMAGE_TYPE::READ, MEMORY_FLAG::GENERATE_MIP_MAPS) == (MEMORY_FLAG::GENERATE_MIP_MAPS | MEMORY_FLAG::READ_WRITE));
	CHECK(has_flag<IMAGE_TYPE::WRITE>(device_image::handle_image_type({ 64, 64, 0, 0 }, tq | IMAGE_TYPE::READ, MEMORY_FLAG::READ | MEMORY_FLAG::GENERATE_MIP_MAPS)));
	CHECK(!has_flag<IMAGE_TYPE::FLAG_MIPMAPPED>(device_image::handle_image_type({ 1, 1, 0, 0 }, tq, MEMORY_FLAG::READ_WRITE)));
}

int main(int argc, char** argv) {
	cpu_checks();
	cuda_context ctx;
	if (argc > 1 && !std::strcmp(argv[1], "--cpu")) {
		if (!ctx.is_supported()) {
			CHECK(ctx.get_devices().empty());
			CHECK(ctx.get_device(device::TYPE::FASTEST_GPU) == nullptr);
			std::printf("dropin_test --cpu: no CUDA device, context reports unsupported (no CPU fallback)\n");
		}
		std::printf("dropin_test --cpu: %s\n", failures ? "FAILED" : "ok");
		return failures ? 1 : 0;
	}
	if (!ctx.is_supported()) {
		std::fprintf(stderr, "dropin_test: CUDA is required\n");
		return 2;
	}
	void* oracle = dlopen(argc > 1 ? argv[1] : "oracle/liboracle_minify.so", RTLD_NOW);
	if (!oracle) {
		std::fprintf(stderr, "dropin_test: cannot load the oracle: %s\n", dlerror());
		return 2;
	}
	const auto flo_generate = (flo_generate_fn)dlsym(oracle, "flo_generate_mip_map_chain");
	const auto flo_fill = (flo_fill_fn)dlsym(oracle, "flo_fill_synthetic");

	const device* dev = ctx.get_device(device::TYPE::FASTEST_GPU);
	auto queue = ctx.create_queue(*dev);
	std::printf("device: %s (%u SMs, sm_%u%u)\n", dev->name.c_str(), dev->units, static_cast<const cuda_device*>(dev)->sm.x, static_cast<const cuda_device*>(dev)->sm.y);

	struct Case { uint4 dim; IMAGE_TYPE type; };
	const Case cases[] = {
		{ { 1024, 1024, 0, 0 }, IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGBA8 },          // BASELINE configs[0]
		{ { 2048, 1024, 0, 0 }, IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGBA16F },
		{ { 256, 256, 5, 0 }, IMAGE_TYPE::IMAGE_2D_ARRAY | IMAGE_TYPE::RGBA8 },
		{ { 128, 128, 2, 0 }, IMAGE_TYPE::IMAGE_CUBE_ARRAY | IMAGE_TYPE::RGBA32F },
		{ { 128, 64, 64, 0 }, IMAGE_TYPE::IMAGE_3D | IMAGE_TYPE::R32F },
		{ { 100, 37, 0, 0 }, IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGBA16 },            // NPOT -> general kernel
	};
	for (const auto& c : cases) {
		++current_case;
		const IMAGE_TYPE type = c.type | IMAGE_TYPE::FLAG_MIPMAPPED | IMAGE_TYPE::READ;
		const uint32_t dim[4] = { c.dim.x, c.dim.y, c.dim.z, c.dim.w };
		const size_t l0_size = image_data_size_from_types(c.dim, type, true), all_size = image_data_size_from_types(c.dim, type);
		std::vector<uint8_t> want(all_size), got(all_size), l0(l0_size);
		flo_fill(l0.data(), dim, image_type_bits(type), 9, 0, image_layer_count(c.dim, type));
		std::memcpy(want.data(), l0.data(), l0_size);
		CHECK(flo_generate(want.data(), dim, image_type_bits(type | IMAGE_TYPE::WRITE), 0, 0, 8) == 0);

		// ctor upload -> chain (GENERATE_MIP_MAPS), then look at every level
		auto img = ctx.create_image(*queue, c.dim, type, l0, MEMORY_FLAG::READ | MEMORY_FLAG::HOST_READ_WRITE | MEMORY_FLAG::GENERATE_MIP_MAPS);
		CHECK(img != nullptr);
		if (!img) continue;
		CHECK(img->get_generate_mip_maps() && img->get_image_data_size() == l0_size);
		CHECK(img->get_mip_level_count() == image_mip_level_count(c.dim, type) && img->get_layer_count() == image_layer_count(c.dim, type));
		CHECK(img->read_levels(*queue, got.data(), got.size(), 0, img->get_mip_level_count() - 1));
		CHECK(got == want);

		// map -> modify -> unmap regenerates the chain
		auto* mapped = static_cast<uint8_t*>(img->map(*queue));
		CHECK(mapped != nullptr && std::memcmp(mapped, l0.data(), l0_size) == 0);
		// new contents = another synthetic pattern (bit flips would create NaN / Inf in the float formats, which the reference's fast-math leaves undefined)
		flo_fill(mapped, dim, image_type_bits(type), 10, 0, image_layer_count(c.dim, type));
		std::vector<uint8_t> l0b(mapped, mapped + l0_size);
		CHECK(img->unmap(*queue, mapped));
		std::memcpy(want.data(), l0b.data(), l0_size);
		CHECK(flo_generate(want.data(), dim, image_type_bits(type | IMAGE_TYPE::WRITE), 0, 0, 8) == 0);
		CHECK(img->read_levels(*queue, got.data(), got.size(), 0, img->get_mip_level_count() - 1));
		CHECK(got == want);

		// write() of the whole level 0 regenerates too; an out-of-bounds write is rejected with false
		const uint3 extent { c.dim.x, image_dim_count(type) >= 2 ? c.dim.y : 1u, image_dim_count(type) >= 3 ? c.dim.z : 1u };
		CHECK(img->write(*queue, l0.data(), l0.size(), { 0, 0, 0 }, extent, { 0, 0 }, { 0, img->get_layer_count() - 1 }));
		std::memcpy(want.data(), l0.data(), l0_size);
		CHECK(flo_generate(want.data(), dim, image_type_bits(type | IMAGE_TYPE::WRITE), 0, 0, 8) == 0);
		CHECK(img->read_levels(*queue, got.data(), got.size(), 0, img->get_mip_level_count() - 1));
		CHECK(got == want);
		CHECK(!img->write(*queue, l0.data(), l0.size(), { 1, 0, 0 }, extent, { 0, 0 }, { 0, 0 }));

		// clone(copy_contents) really copies (the reference's CUDA blit is a `return false` stub), blit refuses a mismatch
		auto twin = img->clone(*queue, true);
		CHECK(twin != nullptr && twin->get_device_ptr() != img->get_device_ptr());
		if (twin) {
			std::fill(got.begin(), got.end(), uint8_t(0));
			CHECK(twin->read_levels(*queue, got.data(), got.size(), 0, twin->get_mip_level_count() - 1));
			CHECK(got == want);
		}
		auto small = ctx.create_image(*queue, { 64, 64, 0, 0 }, IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGBA8 | IMAGE_TYPE::FLAG_MIPMAPPED);
		CHECK(small != nullptr && !small->blit(*queue, *img));

		// an image that lives in floor's tiled storage (CUmipmappedArray): level 0 in, chain, generated levels out
		if (image_dim_count(type) >= 2) {
			void* arr = img->create_tiled_twin();
			CHECK(arr != nullptr);
			if (arr) {
				CHECK(img->copy_to_tiled(*queue, arr, 0, 0));
				auto scratch = img->clone(*queue, false);
				CHECK(scratch != nullptr && scratch->zero(*queue));
				CHECK(scratch && scratch->generate_mip_map_chain_for_tiled(*queue, arr));
				std::fill(got.begin(), got.end(), uint8_t(0));
				CHECK(img->read_tiled_levels(*queue, arr, got.data(), got.size(), 0, img->get_mip_level_count() - 1));
				CHECK(got == want);
				img->destroy_tiled(arr);
			}
		}
	}
	// a batch of independent textures: one graph launch, every chain equal to the oracle's
	{
		++current_case;
		const Case bc[] = { { { 512, 512, 0, 0 }, IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGBA8 }, { { 300, 200, 0, 0 }, IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGBA16F },
							{ { 64, 32, 32, 0 }, IMAGE_TYPE::IMAGE_3D | IMAGE_TYPE::R32F } };
		std::vector<std::shared_ptr<device_image>> imgs;
		std::vector<std::vector<uint8_t>> wants;
		std::vector<device_image*> raw;
		for (const auto& c : bc) {
			const IMAGE_TYPE type = c.type | IMAGE_TYPE::FLAG_MIPMAPPED | IMAGE_TYPE::READ_WRITE;
			const uint32_t dim[4] = { c.dim.x, c.dim.y, c.dim.z, c.dim.w };
			const size_t l0_size = image_data_size_from_types(c.dim, type, true), all_size = image_data_size_from_types(c.dim, type);
			std::vector<uint8_t> want(all_size), l0(l0_size);
			flo_fill(l0.data(), dim, image_type_bits(type), 11, 0, image_layer_count(c.dim, type));
			std::memcpy(want.data(), l0.data(), l0_size);
			CHECK(flo_generate(want.data(), dim, image_type_bits(type), 0, 0, 8) == 0);
			auto img = ctx.create_image(*queue, c.dim, type, MEMORY_FLAG::READ_WRITE | MEMORY_FLAG::HOST_READ_WRITE);
			CHECK(img != nullptr);
			if (!img) continue;
			const uint3 extent { c.dim.x, c.dim.y, image_dim_count(type) >= 3 ? c.dim.z : 1u };
			CHECK(img->write(*queue, l0.data(), l0.size(), { 0, 0, 0 }, extent, { 0, 0 }, { 0, 0 }));
			imgs.push_back(img);
			raw.push_back(img.get());
			wants.push_back(std::move(want));
		}
		mip_chain_batch batch { std::span<device_image* const> { raw.data(), raw.size() } };
		CHECK(batch.is_valid() && batch.generate(*queue));
		for (size_t i = 0; i < imgs.size(); ++i) {
			std::vector<uint8_t> got(wants[i].size());
			CHECK(imgs[i]->read_levels(*queue, got.data(), got.size(), 0, imgs[i]->get_mip_level_count() - 1));
			CHECK(got == wants[i]);
		}
		// the same textures as overlapping chains on one queue (device_queue::set_mip_chain_overlap): zeroed first, so every level read
		// back below was written by these chains
		queue->set_mip_chain_overlap(true);
		for (int rep = 0; rep < 3; ++rep) {
			for (size_t i = 0; i < imgs.size(); ++i) {
				const size_t l0_size = image_data_size_from_types(imgs[i]->get_image_dim(), imgs[i]->get_image_type(), true);
				CHECK(imgs[i]->zero(*queue));
				const uint4 d = imgs[i]->get_image_dim();
				const uint3 extent { d.x, d.y, image_dim_count(imgs[i]->get_image_type()) >= 3 ? d.z : 1u };
				CHECK(imgs[i]->write(*queue, wants[i].data(), l0_size, { 0, 0, 0 }, extent, { 0, 0 }, { 0, 0 }));
			}
			for (auto& img : imgs) CHECK(img->generate_mip_map_chain_async(*queue));
			for (auto& img : imgs) CHECK(img->generate_mip_map_chain_async(*queue)); // and again: the same images are still in the open run
			queue->fence();
			for (size_t i = 0; i < imgs.size(); ++i) {
				std::vector<uint8_t> got(wants[i].size());
				CHECK(imgs[i]->read_levels(*queue, got.data(), got.size(), 0, imgs[i]->get_mip_level_count() - 1));
				CHECK(got == wants[i]);
			}
		}
		queue->set_mip_chain_overlap(false);
	}
	// provide_minify_program (device_image.cpp:155-194): a context-registered program receives the chains of that context's images
	{
		++current_case;
		struct counting_program : minify_program {
			int calls = 0;
			uint32_t last_first_level = ~0u;
			bool minify(device_image& img, const device_queue& cqueue, uint32_t first_level) override {
				++calls;
				last_first_level = first_level;
				return device_image::builtin_minify_program()->minify(img, cqueue, first_level);
			}
		};
		auto prog = std::make_shared<counting_program>();
		CHECK(device_image::provide_minify_program(ctx, prog));
		const uint4 dim { 512, 256, 0, 0 };
		const uint32_t d[4] = { dim.x, dim.y, 0, 0 };
		const IMAGE_TYPE type = IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGBA8 | IMAGE_TYPE::FLAG_MIPMAPPED | IMAGE_TYPE::READ_WRITE;
		const size_t l0_size = image_data_size_from_types(dim, type, true), all_size = image_data_size_from_types(dim, type);
		std::vector<uint8_t> want(all_size), got(all_size), l0(l0_size);
		flo_fill(l0.data(), d, image_type_bits(type), 12, 0, 1);
		std::memcpy(want.data(), l0.data(), l0_size);
		CHECK(flo_generate(want.data(), d, image_type_bits(type), 0, 0, 8) == 0);
		auto img = ctx.create_image(*queue, dim, type, l0, MEMORY_FLAG::READ_WRITE | MEMORY_FLAG::HOST_READ_WRITE | MEMORY_FLAG::GENERATE_MIP_MAPS);
		CHECK(img != nullptr && prog->calls == 1 && prog->last_first_level == 0u);
		if (img) {
			CHECK(img->read_levels(*queue, got.data(), got.size(), 0, img->get_mip_level_count() - 1));
			CHECK(got == want);
			CHECK(img->generate_mip_map_chain_async(*queue, 3u) && prog->calls == 2 && prog->last_first_level == 3u);
			queue->finish();
		}
		// another context is not affected; nullptr restores the built-in program
		CHECK(device_image::provide_minify_program(ctx, nullptr));
		if (img) {
			img->generate_mip_map_chain(*queue);
			CHECK(prog->calls == 2);
		}
		// a host-read-only image refuses write() (write_check, device_image.cpp:503-547), and so does a zero-size source
		auto ro = ctx.create_image(*queue, dim, type, MEMORY_FLAG::READ_WRITE | MEMORY_FLAG::HOST_READ);
		CHECK(ro != nullptr && !ro->write(*queue, l0.data(), l0.size(), { 0, 0, 0 }, { dim.x, dim.y, 1 }, { 0, 0 }, { 0, 0 }));
		CHECK(img && !img->write(*queue, l0.data(), 0, { 0, 0, 0 }, { dim.x, dim.y, 1 }, { 0, 0 }, { 0, 0 }));
		// FASTEST_GPU: highest units x clock (cuda_context.cpp:340-395)
		const device* fastest = ctx.get_device(device::TYPE::FASTEST_GPU);
		CHECK(fastest != nullptr && fastest->clock > 0u);
		for (const device* d2 : ctx.get_devices()) CHECK(uint64_t(d2->units) * d2->clock <= uint64_t(fastest->units) * fastest->clock);
	}
	// constructor invariants throw (device_image.hpp:502-539); formats without a kernel return nullptr
	bool threw = false;
	try {
		ctx.create_image(*queue, { 64, 64, 0, 0 }, IMAGE_TYPE::IMAGE_2D_MSAA | IMAGE_TYPE::RGBA8 | IMAGE_TYPE::FLAG_MIPMAPPED);
	} catch (const std::runtime_error&) { threw = true; }
	CHECK(threw);
	// 64-bit formats have no kernel: nullptr; 3-channel images, which the reference's CUDA backend rejects (cuda_image.cpp:173-180), exist here
	CHECK(ctx.create_image(*queue, { 64, 64, 0, 0 }, IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::FORMAT_64 | IMAGE_TYPE::FLOAT | IMAGE_TYPE::CHANNELS_1 | IMAGE_TYPE::FLAG_MIPMAPPED) == nullptr);
	CHECK(ctx.create_image(*queue, { 64, 64, 0, 0 }, IMAGE_TYPE::IMAGE_2D | IMAGE_TYPE::RGB8 | IMAGE_TYPE::FLAG_MIPMAPPED) != nullptr);
	queue->start_profiling();
	CHECK(queue->stop_profiling() < 1000000u);
	std::printf("dropin_test: %s\n", failures ? "FAILED" : "ok");
	return failures ? 1 : 0;
}
