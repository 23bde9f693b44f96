// container_tool -- test helper for the container layer (no GPU involved):
//   container_tool info <file>              -> one JSON line: type, layout, chunk list
//   container_tool invert <in> <out>        -> out = in with every payload byte XOR 0xFF
// `invert` exercises exactly the path process_file uses to produce its output:
// create_output() (all non-sample bytes verbatim) + write_payload().
#include <cstring>
#include <format>
#include <iostream>
#include <vector>

#include "audio_container.hpp"
#include "errors.hpp"

using namespace lowcut;

int main(int argc, char** argv)
{
	try {
		if (argc >= 3 && !std::strcmp(argv[1], "info")) {
			AudioContainer c(argv[2]);
			const PcmLayout& l = c.pcm();
			std::string chunks;
			for (const ChunkInfo& k : c.chunks()) {
				std::string id;
				for (int i = 0; i < 4; ++i) id += (k.id[i] >= 32 && k.id[i] < 127 && k.id[i] != '"' && k.id[i] != '\\') ? k.id[i] : '?';
				chunks += std::format("{}{{\"id\":\"{}\",\"offset\":{},\"size\":{}}}", chunks.empty() ? "" : ",", id,
				                      k.header_offset, k.size);
			}
			std::cout << std::format(
				"{{\"type\":\"{}\",\"channels\":{},\"bits\":{},\"valid_bits\":{},\"big_endian\":{},\"rate\":{},"
				"\"frames\":{},\"payload_offset\":{},\"payload_bytes\":{},\"file_size\":{},\"chunks\":[{}]}}",
				c.type_name(), l.channels, l.bits, l.valid_bits, l.big_endian ? "true" : "false", l.sample_rate,
				l.frames, l.payload_offset, l.payload_bytes, c.file_size(), chunks) << std::endl;
			return 0;
		}
		if (argc >= 4 && !std::strcmp(argv[1], "invert")) {
			AudioContainer c(argv[2]);
			const PcmLayout& l = c.pcm();
			std::vector<unsigned char> buf(l.payload_bytes);
			c.read_payload(0, l.payload_bytes, buf.data());
			for (auto& b : buf) b ^= 0xFF;
			const int fd = c.create_output(argv[3]);
			AudioContainer::write_payload(fd, l, 0, l.payload_bytes, buf.data());
			AudioContainer::close_output(fd);
			return 0;
		}
		std::cerr << "usage: container_tool info <file> | invert <in> <out>" << std::endl;
		return 2;
	} catch (const std::exception& e) {
		std::cerr << e.what() << std::endl;
		return 1;
	}
}
