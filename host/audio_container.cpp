#include "audio_container.hpp"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <format>

#include "errors.hpp"

namespace lowcut {
namespace {

void pread_all(int fd, void* dst, uint64_t n, uint64_t off, const std::string& what)
{
	auto* p = static_cast<unsigned char*>(dst);
	while (n) {
		const ssize_t r = ::pread(fd, p, n > (1u << 30) ? (1u << 30) : n, (off_t) off);
		if (r < 0 && errno == EINTR) continue;
		if (r <= 0) throw FormatError("short read in " + what);
		p += r;
		off += (uint64_t) r;
		n -= (uint64_t) r;
	}
}

void pwrite_all(int fd, const void* src, uint64_t n, uint64_t off)
{
	auto* p = static_cast<const unsigned char*>(src);
	while (n) {
		const ssize_t r = ::pwrite(fd, p, n > (1u << 30) ? (1u << 30) : n, (off_t) off);
		if (r < 0 && errno == EINTR) continue;
		if (r <= 0) throw std::runtime_error(std::string("write failed: ") + std::strerror(errno));
		p += r;
		off += (uint64_t) r;
		n -= (uint64_t) r;
	}
}

uint32_t le32(const unsigned char* p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t) p[3] << 24; }
uint16_t le16(const unsigned char* p) { return (uint16_t) (p[0] | p[1] << 8); }
uint64_t le64(const unsigned char* p) { return le32(p) | (uint64_t) le32(p + 4) << 32; }
uint32_t be32(const unsigned char* p) { return (uint32_t) p[0] << 24 | p[1] << 16 | p[2] << 8 | p[3]; }
uint16_t be16(const unsigned char* p) { return (uint16_t) (p[0] << 8 | p[1]); }

int container_bits(int valid_bits, int bytes_per_sample, const std::string& what)
{
	if (bytes_per_sample < 2 || bytes_per_sample > 4 || valid_bits < 9 || valid_bits > 8 * bytes_per_sample)
		throw FormatError(std::format("{}: only 16-, 24- and 32-bit integer PCM is supported ({} valid bits in "
		                              "{} bytes)", what, valid_bits, bytes_per_sample));
	return 8 * bytes_per_sample;
}

} // namespace

double extended80_to_double(const unsigned char* p)
{
	const int sign = p[0] >> 7;
	const int exp = ((p[0] & 0x7f) << 8) | p[1];
	uint64_t mant = 0;
	for (int i = 0; i < 8; ++i) mant = mant << 8 | p[2 + i];
	if (exp == 0 && mant == 0) return 0.0;
	const double v = std::ldexp((double) mant, exp - 16383 - 63);
	return sign ? -v : v;
}

UniqueFd::~UniqueFd()
{
	if (fd >= 0) ::close(fd);
}

AudioContainer::AudioContainer(const std::filesystem::path& path) : path_(path)
{
	fd_.fd = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
	if (fd_ < 0) throw FileNotFound(path.string());
	struct stat st{};
	if (::fstat(fd_, &st) != 0) throw FormatError("cannot stat " + path.string());
	size_ = (uint64_t) st.st_size;
	if (size_ < 12) throw FormatError(path.string() + ": not a WAVE or AIFF file (too short)");
	unsigned char hdr[12];
	pread_all(fd_, hdr, 12, 0, path.string());
	if ((!std::memcmp(hdr, "RIFF", 4) || !std::memcmp(hdr, "RF64", 4)) && !std::memcmp(hdr + 8, "WAVE", 4)) {
		type_ = hdr[1] == 'F' ? ContainerType::Rf64 : ContainerType::Wave;
		parse_riff();
	} else if (!std::memcmp(hdr, "FORM", 4) && (!std::memcmp(hdr + 8, "AIFF", 4) || !std::memcmp(hdr + 8, "AIFC", 4))) {
		type_ = hdr[11] == 'C' ? ContainerType::Aifc : ContainerType::Aiff;
		parse_iff();
	} else {
		throw FormatError(path.string() + ": not a WAVE or AIFF file");
	}
	if (pcm_.channels < 1) throw FormatError(path.string() + ": no format chunk before the samples");
	if (pcm_.payload_offset == 0) throw FormatError(path.string() + ": no sample chunk");
	if (pcm_.payload_offset > size_) throw FormatError(path.string() + ": the sample chunk starts beyond the end of the file");
	const uint64_t fb = (uint64_t) pcm_.channels * (pcm_.bits / 8);
	// a truncated file keeps what is there; partial trailing frames are not samples.  Sizes come
	// from the file (an RF64 ds64 size can be anything up to 2^64): compare by subtraction, a sum
	// could wrap and slip past the clamp
	pcm_.payload_bytes = std::min(pcm_.payload_bytes, size_ - pcm_.payload_offset);
	pcm_.frames = pcm_.payload_bytes / fb;
	pcm_.payload_bytes = pcm_.frames * fb;
}

const char* AudioContainer::type_name() const
{
	switch (type_) {
	case ContainerType::Wave: return "WAVE";
	case ContainerType::Rf64: return "RF64";
	case ContainerType::Aiff: return "AIFF";
	default: return "AIFF-C";
	}
}

void AudioContainer::parse_riff()
{
	uint64_t pos = 12, ds64_data = 0;
	bool have_ds64 = false;
	while (pos + 8 <= size_) {
		unsigned char h[8];
		pread_all(fd_, h, 8, pos, path_.string());
		ChunkInfo c{};
		std::memcpy(c.id, h, 4);
		c.header_offset = pos;
		c.data_offset = pos + 8;
		c.size = le32(h + 4);
		if (!std::memcmp(c.id, "ds64", 4) && c.size >= 24 && c.data_offset + 24 <= size_) {
			unsigned char d[24];
			pread_all(fd_, d, 24, c.data_offset, path_.string());
			ds64_data = le64(d + 8);
			have_ds64 = true;
		}
		if (!std::memcmp(c.id, "data", 4) && c.size == 0xFFFFFFFFu && have_ds64) c.size = ds64_data;
		// a chunk cannot reach beyond the end of the file (truncated file, hostile ds64 size): it
		// ends where the file ends, and the walk stops there instead of wrapping around
		const uint64_t room = size_ - c.data_offset; // pos + 8 <= size_
		const bool truncated = c.size > room;
		if (truncated) c.size = room;
		if (!std::memcmp(c.id, "fmt ", 4)) {
			unsigned char f[40] = {};
			const uint64_t n = std::min<uint64_t>(c.size, 40);
			if (c.size < 16 || c.data_offset + n > size_) throw FormatError(path_.string() + ": bad fmt chunk");
			pread_all(fd_, f, n, c.data_offset, path_.string());
			int tag = le16(f);
			const int channels = le16(f + 2), block = le16(f + 12), vbits = le16(f + 14);
			if (tag == 0xFFFE && c.size >= 26) tag = le16(f + 24); // WAVE_FORMAT_EXTENSIBLE: SubFormat GUID
			if (tag != 1)
				throw FormatError(std::format("{}: WAVE format tag {} is not integer PCM", path_.string(), tag));
			if (channels < 1 || block % channels) throw FormatError(path_.string() + ": bad fmt chunk");
			pcm_.channels = channels;
			pcm_.valid_bits = vbits;
			pcm_.bits = container_bits(vbits, block / channels, path_.string());
			pcm_.big_endian = false;
			pcm_.sample_rate = (double) le32(f + 4);
		}
		if (!std::memcmp(c.id, "data", 4) && pcm_.payload_offset == 0) {
			pcm_.payload_offset = c.data_offset;
			pcm_.payload_bytes = c.size;
		}
		chunks_.push_back(c);
		if (truncated) break;
		pos = c.data_offset + c.size + (c.size & 1); // <= size_ + 1: cannot wrap
	}
}

void AudioContainer::parse_iff()
{
	uint64_t pos = 12, comm_frames = 0;
	uint64_t ssnd_data = 0, ssnd_size = 0;
	while (pos + 8 <= size_) {
		unsigned char h[8];
		pread_all(fd_, h, 8, pos, path_.string());
		ChunkInfo c{};
		std::memcpy(c.id, h, 4);
		c.header_offset = pos;
		c.data_offset = pos + 8;
		c.size = be32(h + 4);
		const uint64_t room = size_ - c.data_offset; // pos + 8 <= size_
		const bool truncated = c.size > room;
		if (truncated) c.size = room;
		if (!std::memcmp(c.id, "COMM", 4)) {
			unsigned char f[22] = {};
			const uint64_t n = std::min<uint64_t>(c.size, 22);
			if (c.size < 18 || c.data_offset + n > size_) throw FormatError(path_.string() + ": bad COMM chunk");
			pread_all(fd_, f, n, c.data_offset, path_.string());
			pcm_.channels = (int16_t) be16(f);
			comm_frames = be32(f + 2);
			pcm_.valid_bits = (int16_t) be16(f + 6);
			pcm_.sample_rate = extended80_to_double(f + 8);
			pcm_.big_endian = true;
			if (type_ == ContainerType::Aifc) {
				if (c.size < 22) throw FormatError(path_.string() + ": bad AIFF-C COMM chunk");
				if (!std::memcmp(f + 18, "sowt", 4)) pcm_.big_endian = false;
				else if (std::memcmp(f + 18, "NONE", 4) && std::memcmp(f + 18, "twos", 4))
					throw FormatError(std::format("{}: AIFF-C compression '{}' is not integer PCM", path_.string(),
					                              std::string((const char*) f + 18, 4)));
			}
			if (pcm_.channels < 1) throw FormatError(path_.string() + ": bad COMM chunk");
			pcm_.bits = container_bits(pcm_.valid_bits, (pcm_.valid_bits + 7) / 8, path_.string());
		}
		if (!std::memcmp(c.id, "SSND", 4) && ssnd_data == 0 && c.size >= 8) {
			unsigned char f[8];
			pread_all(fd_, f, 8, c.data_offset, path_.string());
			const uint64_t offset = be32(f); // + blockSize at f+4: alignment hint only
			if (offset > c.size - 8) throw FormatError(path_.string() + ": SSND offset lies beyond its chunk");
			ssnd_data = c.data_offset + 8 + offset;
			ssnd_size = c.size - 8 - offset;
		}
		chunks_.push_back(c);
		if (truncated) break;
		pos = c.data_offset + c.size + (c.size & 1); // <= size_ + 1: cannot wrap
	}
	if (ssnd_data && pcm_.channels > 0) { // COMM may come after SSND
		pcm_.payload_offset = ssnd_data;
		const uint64_t want = comm_frames * (uint64_t) pcm_.channels * (pcm_.bits / 8);
		pcm_.payload_bytes = std::min(want, ssnd_size);
	}
}

void AudioContainer::advise_willneed() const
{
#if defined(POSIX_FADV_WILLNEED)
	(void) ::posix_fadvise(fd_, 0, (off_t) size_, POSIX_FADV_WILLNEED);
#endif
}

void AudioContainer::read_payload(uint64_t offset, uint64_t n, void* dst) const
{
	if (offset > pcm_.payload_bytes || n > pcm_.payload_bytes - offset) throw FormatError("payload read out of range");
	pread_all(fd_, dst, n, pcm_.payload_offset + offset, path_.string());
}

int AudioContainer::create_output(const std::filesystem::path& out) const
{
	const int fd = ::open(out.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
	if (fd < 0) throw std::runtime_error("cannot create " + out.string() + ": " + std::strerror(errno));
	try {
		if (::ftruncate(fd, (off_t) size_) != 0) throw std::runtime_error("cannot size " + out.string());
#if defined(__linux__)
		// Have the file system back the whole file NOW, in one call, while the GPU is still filtering:
		// on tmpfs / the page cache, writing into pages that already exist is a plain copy, whereas
		// every first touch of a fresh page is a fault -- measured 2x slower per thread and much worse
		// when 32 lanes do it at once.  Best effort: file systems without fallocate just skip it.
		if (!std::getenv("LOWCUT_NO_FALLOCATE")) (void) ::fallocate(fd, 0, 0, (off_t) size_);
#endif
		std::vector<unsigned char> buf(1u << 22);
		auto copy = [&](uint64_t a, uint64_t b) {
			while (a < b) {
				const uint64_t n = std::min<uint64_t>(buf.size(), b - a);
				pread_all(fd_, buf.data(), n, a, path_.string());
				pwrite_all(fd, buf.data(), n, a);
				a += n;
			}
		};
		copy(0, pcm_.payload_offset);                          // headers and chunks before the samples
		copy(pcm_.payload_offset + pcm_.payload_bytes, size_); // partial frame, pad byte, trailing chunks
	} catch (...) {
		::close(fd);
		throw;
	}
	return fd;
}

void AudioContainer::write_payload(int fd, const PcmLayout& pcm, uint64_t offset, uint64_t n, const void* src)
{
	if (offset > pcm.payload_bytes || n > pcm.payload_bytes - offset) throw FormatError("payload write out of range");
	pwrite_all(fd, src, n, pcm.payload_offset + offset);
}

void AudioContainer::close_output(int fd)
{
	if (fd >= 0 && ::close(fd) != 0) throw std::runtime_error(std::string("close failed: ") + std::strerror(errno));
}

} // namespace lowcut
