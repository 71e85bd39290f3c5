// BENCH / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): opens the reference's prebuilt mip-map-minify archive
// (etc/mip_map_minify/mmm.fubar, embedded into libfloor by src/device/device_image.cpp:133-139) and writes the PTX of its
// CUDA target to a file, so that bench.py --impl incumbent can JIT and time the reference's own GPU kernels on the B200.
//
// Built by oracle/build_incumbent.py together with the reference's src/core/bcm.cpp *where it lies* under /root/reference
// (the archive's payload is one BCM stream, universal_binary.cpp:135-148); nothing of the reference is copied into this
// repository, the extracted PTX only goes to oracle/_ref/ (git-ignored, travels to the GPU box).
//
// Archive layout (include/floor/device/universal_binary.hpp:25-56): "FUBA", u32 version, u32 count, u32 flags,
// target u64[count] (bits 0-3 version, bits 4-7 PLATFORM_TYPE: CUDA = 2), offsets u64[count], toolchain versions u32[count],
// sha-256 [count]; then the binaries: {u32 function_count, u32 function_info_size, u32 binary_size, u32 flags},
// function infos, binary data.
#include <floor/core/bcm.hpp>
#include <floor/core/logger.hpp>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <sstream>
#include <vector>

// the two logger symbols bcm.cpp references (never reached on a valid stream)
namespace fl {
bool logger::prepare_log(std::stringstream&, const LOG_TYPE&, const char*, const char*) { return true; }
void logger::log_internal(std::stringstream& s, const LOG_TYPE&, const char*) { std::fprintf(stderr, "bcm: %s\n", s.str().c_str()); }
} // namespace fl

int main(int argc, char** argv) {
	if (argc < 3) {
		std::fprintf(stderr, "usage: fubar_extract <archive.fubar> <out.ptx>\n");
		return 2;
	}
	std::ifstream f(argv[1], std::ios::binary);
	std::vector<uint8_t> ar((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	if (ar.size() < 16 || std::memcmp(ar.data(), "FUBA", 4) != 0) {
		std::fprintf(stderr, "not a FUBAR archive\n");
		return 1;
	}
	uint32_t version, count, flags;
	std::memcpy(&version, &ar[4], 4);
	std::memcpy(&count, &ar[8], 4);
	std::memcpy(&flags, &ar[12], 4);
	const size_t header = 16 + (size_t)count * (8 + 8 + 4 + 32);
	std::vector<uint64_t> targets(count), offsets(count);
	std::memcpy(targets.data(), &ar[16], count * 8);
	std::memcpy(offsets.data(), &ar[16 + count * 8], count * 8);
	std::vector<uint8_t> payload;
	if (flags & 1u) { // is_compressed
		payload = fl::bcm::bcm_decompress(std::span<const uint8_t> { ar.data() + header, ar.size() - header });
		if (payload.empty()) {
			std::fprintf(stderr, "BCM decompression failed\n");
			return 1;
		}
	} else {
		payload.assign(ar.begin() + (long)header, ar.end());
	}
	std::printf("FUBAR v%u, %u binaries, flags %#x, payload %zu bytes\n", version, count, flags, payload.size());
	for (uint32_t i = 0; i < count; ++i) {
		const uint32_t type = (uint32_t)((targets[i] >> 4) & 0xF);
		if (type != 2u) continue; // PLATFORM_TYPE::CUDA
		const size_t at = offsets[i] - header;
		uint32_t hdr[4];
		std::memcpy(hdr, &payload[at], 16);
		const size_t data = at + 16 + hdr[1];
		std::printf("binary #%u: CUDA target %#llx, %u functions, %u bytes\n", i, (unsigned long long)targets[i], hdr[0], hdr[2]);
		if (data + hdr[2] > payload.size() || std::memcmp(&payload[data], "//", 2) != 0) {
			std::fprintf(stderr, "binary #%u is not PTX text\n", i);
			return 1;
		}
		std::ofstream out(argv[2], std::ios::binary);
		out.write((const char*)&payload[data], hdr[2]);
		return 0;
	}
	std::fprintf(stderr, "no CUDA binary in the archive\n");
	return 1;
}
