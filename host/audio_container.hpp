// audio_container.hpp -- chunk-order-preserving RIFF/WAVE (+RF64) and IFF/AIFF(-C) access.
//
// Re-creates, for the hot path's host, what the reference gets from c_lib's AudioFile /
// AudioFormat (ProcessFile.cp:34-35,43,49,105-116): where the sample payload sits, how
// it is laid out (channels, bits, endianness, rate), and an output file that carries
// EVERY byte of the input outside that payload unchanged (the reference copies every
// chunk verbatim, then overwrites the samples: ProcessFile.cp:107-117).
//
// The parser never interprets non-audio chunks (LIST, bext, ID3, MARK, ...): it only
// walks ids and sizes, so unknown chunks, odd sizes with their pad byte and chunks that
// FOLLOW the sample chunk survive byte for byte.
#pragma once
#include <cstdint>
#include <filesystem>
#include <string>
#include <vector>

namespace lowcut {

enum class ContainerType { Wave, Rf64, Aiff, Aifc };

struct ChunkInfo {
	char id[5];
	uint64_t header_offset; // of the 8-byte chunk header in the file
	uint64_t data_offset;
	uint64_t size;          // as declared (RF64: resolved through ds64)
};

struct PcmLayout {
	uint64_t payload_offset = 0; // first sample byte in the file
	uint64_t payload_bytes = 0;  // whole frames only
	uint64_t frames = 0;
	int channels = 0;
	int bits = 0;                // container width: 16, 24 or 32
	int valid_bits = 0;          // as declared (e.g. 20 in a 24-bit container)
	bool big_endian = false;
	double sample_rate = 0.0;
};

// A file descriptor that is closed when its owner goes away -- including when the owner's
// constructor throws half-way (a destructor does not run for a partially built object, the
// destructors of its finished members do).
struct UniqueFd {
	int fd = -1;
	UniqueFd() = default;
	explicit UniqueFd(int f) : fd(f) {}
	UniqueFd(const UniqueFd&) = delete;
	UniqueFd& operator=(const UniqueFd&) = delete;
	~UniqueFd();
	operator int() const { return fd; }
};

class AudioContainer {
public:
	// Parses the chunk structure; throws FormatError on anything that is not
	// integer PCM of 16/24/32 bits in a WAVE/RF64/AIFF/AIFF-C container.
	explicit AudioContainer(const std::filesystem::path& path);
	AudioContainer(const AudioContainer&) = delete;
	AudioContainer& operator=(const AudioContainer&) = delete;

	ContainerType type() const { return type_; }
	const char* type_name() const;
	const PcmLayout& pcm() const { return pcm_; }
	const std::vector<ChunkInfo>& chunks() const { return chunks_; }
	uint64_t file_size() const { return size_; }

	// Hint the kernel to start reading the file into the page cache (best effort).
	void advise_willneed() const;

	// Read payload bytes [offset, offset+n) (relative to the payload) into dst.
	void read_payload(uint64_t offset, uint64_t n, void* dst) const;

	// Creates `out` as a byte-for-byte copy of this file outside the payload and
	// leaves the payload region to be filled with write_payload().  Returns an fd.
	int create_output(const std::filesystem::path& out) const;
	static void write_payload(int fd, const PcmLayout& pcm, uint64_t offset, uint64_t n, const void* src);
	static void close_output(int fd);

private:
	void parse_riff();
	void parse_iff();
	std::filesystem::path path_;
	UniqueFd fd_;
	uint64_t size_ = 0;
	ContainerType type_ = ContainerType::Wave;
	std::vector<ChunkInfo> chunks_;
	PcmLayout pcm_;
};

// 80-bit IEEE 754 extended (AIFF COMM sampleRate) -> double.
double extended80_to_double(const unsigned char* p);

} // namespace lowcut
